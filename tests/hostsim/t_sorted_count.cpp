// TEST INFRASTRUCTURE ONLY: exhaustive check of b2c_sorted_count (csrc/b2c_beam_fast.h, the branch-free binary search of
// the merge-free sorted step) against a linear scan, for every table size 1..128 and every answer 0..n.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "cuda_shim.h"
#include "b2c_beam.h"
#include "b2c_prepare.h"      // B2cFrameRec, B2C_RUN
struct B2cBeamArgs {
    B2cParams P; B2cLayout L; int n_utts; const int* order; u32* next; const u64* frame_off; const int* T; const B2cFrameRec* tok_rec;
    const u32* tok_ids; const double* tok_lp; u8* gws; const B2cLmState* start_states; int* out_nbeams; int* out_status; double* out_scores;
    int* out_ntok; int* out_nwords; u32* out_toks; int* out_frames; B2cLmState* out_states; int chunk_t0, chunk_t1, chunk_last, pad_chunk;
    u8* state; u64 state_stride; const u32* gate; int gate_n; int gate_bounds[5]; u64* phase_clk; u32* m_stats;
};
#include "b2c_beam_fast.h"
int main() {
    double logit[128];
    int bad = 0;
    for (int n = 1; n <= 128; ++n) {
        for (int i = 0; i < n; ++i) logit[i] = -0.01 * (i / 2);          // non-increasing, with ties
        for (int t = 0; t <= n; ++t) {
            const double s = t == n ? -100.0 : logit[t] + 0.0;
            for (int ge = 0; ge < 2; ++ge) {
                int want = 0;
                for (int i = 0; i < n; ++i) { const double s2 = (logit[i] + 0.0) + 0.0; want += ge ? (s2 >= s) : (s2 > s); }
                const u32 got = b2c_sorted_count<128>(logit, static_cast<u32>(n), 0.0, s, ge != 0);
                if (static_cast<int>(got) != want) { if (bad < 10) std::printf("n=%d t=%d ge=%d want %d got %u\n", n, t, ge, want, got); ++bad; }
            }
        }
    }
    std::printf("bad %d\n", bad);
    return bad ? 1 : 0;
}
