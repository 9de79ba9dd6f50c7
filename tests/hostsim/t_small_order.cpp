// TEST INFRASTRUCTURE ONLY: b2c_pyset_small_order_any (csrc/b2c_prepare.h: CPython set order of up to three selected tokens
// plus the arg-max, kept in registers by the wide-alphabet streaming kernel) against the general set emulation
// (b2c_pyset_add / b2c_pyset_copy_or, which tests/test_oracle.py checks against real CPython sets): every subset of
// size 0..3 of a token range with collisions in the low bits, every arg-max of another range.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "cuda_shim.h"
#include "b2c_beam.h"
#include "b2c_prepare.h"
static int check(const std::vector<u32>& keys, u32 amax, u16* b0, u16* b1) {
    B2cPySet set;
    b2c_pyset_init(set, b0, b1);
    for (u32 k : keys) b2c_pyset_add(set, k);
    b2c_pyset_copy_or(set, amax);
    std::vector<u32> want;
    const u16* tab = set.buf[set.cur];
    for (u32 s = 0; s <= set.mask; ++s) if (tab[s] != 0xFFFFu) want.push_back(tab[s]);
    u32 out[4] = {0, 0, 0, 0};
    const u32 n = b2c_pyset_small_order_any(keys.size() > 0 ? keys[0] : 0, keys.size() > 1 ? keys[1] : 0, keys.size() > 2 ? keys[2] : 0,
                                            static_cast<u32>(keys.size()), amax, out);
    if (n != want.size()) return 1;
    for (u32 i = 0; i < n; ++i) if (out[i] != want[i]) return 1;
    return 0;
}
int main() {
    static u16 b0[65536], b1[65536];
    // token ids that collide in the low three bits and differ in the perturbation bits
    std::vector<u32> pool;
    for (u32 hi : {0u, 1u, 2u, 5u, 31u, 32u, 33u, 127u}) for (u32 lo : {0u, 1u, 3u, 7u}) pool.push_back(hi * 8 + lo);
    for (u32 v : {1016u, 1023u, 4095u, 30000u, 65000u}) pool.push_back(v);
    std::vector<u32> sorted = pool;
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    long bad = 0, n = 0;
    const size_t P = sorted.size();
    for (u32 amax : sorted) {
        bad += check({}, amax, b0, b1); ++n;
        for (size_t a = 0; a < P; ++a) {
            bad += check({sorted[a]}, amax, b0, b1); ++n;
            for (size_t b = a + 1; b < P; ++b) {
                bad += check({sorted[a], sorted[b]}, amax, b0, b1); ++n;
                for (size_t c = b + 1; c < P; ++c) { bad += check({sorted[a], sorted[b], sorted[c]}, amax, b0, b1); ++n; }
            }
        }
    }
    std::printf("cases %ld bad %ld\n", n, bad);
    return bad ? 1 : 0;
}
