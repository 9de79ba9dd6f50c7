// b200ctc -- prepare kernel body: input normalisation + per-frame token selection.
//
// One CTA per utterance.  Restates reference decoder.py:756-765 (probabilities vs logits,
// log-softmax, clipping) and decoder.py:444-445 (tokens >= token_min_logp united with the
// argmax, in CPython set iteration order) and writes one compact (token id, log-prob) list
// per frame, so that the beam kernel never touches the [T,V] matrix again.
//
// Streaming and HBM bound: every logit is read from HBM once (the second and third pass of
// a row hit L1/L2), algorithmic bytes per frame = V * sizeof(dtype).
//
// Numerical definition shared with the oracle (oracle/ctc_oracle.cpp normalise_rows):
//   logits branch  d = fl_T(x - max), S = sum exp((double)d) in "warp order" (32 lane-strided
//                  partial sums, then xor butterfly 16,8,4,2,1), lp = fl_T((double)d - log S),
//                  clipped in float64 to [log(1e-15), 0];
//   probs branch   lp = fl_T(log((double)clip(x, fl_T(1e-15), 1)));
//   the branch decision is math.isclose(x.sum(axis=1).mean(), 1) evaluated bit-exactly like
//   numpy does (pairwise summation in the input dtype).
#pragma once
#include "b2c_cta.h"

// ---------------------------------------------------------------------------------------
// numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, PW_BLOCKSIZE = 128)
// ---------------------------------------------------------------------------------------
template <class T>
B2C_HD T b2c_np_leaf_sum(const T* a, long n) {
    if (n < 8) {
        T res = static_cast<T>(-0.0);
        for (long i = 0; i < n; ++i) res = res + a[i];
        return res;
    }
    T r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    long i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 = r0 + a[i + 0]; r1 = r1 + a[i + 1]; r2 = r2 + a[i + 2]; r3 = r3 + a[i + 3];
        r4 = r4 + a[i + 4]; r5 = r5 + a[i + 5]; r6 = r6 + a[i + 6]; r7 = r7 + a[i + 7];
    }
    T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res = res + a[i];
    return res;
}

// full pairwise sum; `leaf` returns the sum of the leaf block [off, off+n) (n <= 128)
struct B2cPwFrame { long off, n; int st; };
template <class T, class LeafFn>
B2C_HD T b2c_np_pairwise_generic(long n, LeafFn leaf) {
    B2cPwFrame stk[48];
    T left[48];
    int sp = 0;
    T ret = static_cast<T>(0);
    stk[0].off = 0; stk[0].n = n; stk[0].st = 0;
    sp = 1;
    while (sp > 0) {
        B2cPwFrame& f = stk[sp - 1];
        if (f.n <= 128) { ret = leaf(f.off, f.n); --sp; continue; }
        long n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.st == 0) {
            f.st = 1;
            stk[sp].off = f.off; stk[sp].n = n2; stk[sp].st = 0;
            ++sp;
        } else if (f.st == 1) {
            left[sp - 1] = ret;
            f.st = 2;
            stk[sp].off = f.off + n2; stk[sp].n = f.n - n2; stk[sp].st = 0;
            ++sp;
        } else {
            ret = left[sp - 1] + ret;
            --sp;
        }
    }
    return ret;
}
template <class T>
struct B2cLeafDirect {
    const T* a;
    B2C_HD T operator()(long off, long n) const { return b2c_np_leaf_sum(a + off, n); }
};
template <class T>
B2C_HD T b2c_np_pairwise(const T* a, long n) {
    B2cLeafDirect<T> lf;
    lf.a = a;
    return b2c_np_pairwise_generic<T>(n, lf);
}

// math.isclose(x, 1.0) with the default rel_tol=1e-9, abs_tol=0
B2C_HD bool b2c_isclose_one(double x) {
    if (x == 1.0) return true;
    if (!(x - x == 0.0)) return false;  // inf or nan
    double diff = fabs(1.0 - x);
    return (diff <= fabs(1e-9 * 1.0)) || (diff <= fabs(1e-9 * x));
}

// ---------------------------------------------------------------------------------------
// per-element log-probability
// ---------------------------------------------------------------------------------------
template <class T>
B2C_HD double b2c_lp_logit(T x, T m, double ls) {
    const T d = x - m;
    double lp = static_cast<double>(static_cast<T>(static_cast<double>(d) - ls));
    if (lp < B2C_LOG_MIN_CLIP) lp = B2C_LOG_MIN_CLIP;
    if (lp > 0.0) lp = 0.0;
    return lp;
}
template <class T>
B2C_HD double b2c_lp_prob(T x) {
    const T lo = static_cast<T>(1e-15);
    T c = x;
    if (c < lo) c = lo;
    if (c > static_cast<T>(1)) c = static_cast<T>(1);
    return static_cast<double>(static_cast<T>(log(static_cast<double>(c))));
}
template <class T>
B2C_HD double b2c_lp(T x, bool is_prob, T m, double ls) { return is_prob ? b2c_lp_prob<T>(x) : b2c_lp_logit<T>(x, m, ls); }

// ---------------------------------------------------------------------------------------
// CPython 3.12 set of small non-negative ints (Objects/setobject.c), tables of u16.
// Two ping-pong buffers of `cap` entries each; 0xFFFF marks an empty slot.
// ---------------------------------------------------------------------------------------
struct B2cPySet {
    u16* buf[2];
    int cur;
    u32 mask, fill;
};
B2C_HD void b2c_pyset_clear(u16* t, u32 size) { for (u32 i = 0; i < size; ++i) t[i] = 0xFFFFu; }
B2C_HD void b2c_pyset_init(B2cPySet& s, u16* b0, u16* b1) {
    s.buf[0] = b0; s.buf[1] = b1; s.cur = 0; s.mask = 7; s.fill = 0;
    b2c_pyset_clear(b0, 8);
}
B2C_HD void b2c_pyset_insert_clean(u16* t, u32 mask, u32 key) {
    u32 perturb = key;
    u32 i = key & mask;
    while (true) {
        if (t[i] == 0xFFFFu) { t[i] = static_cast<u16>(key); return; }
        if (i + 9 <= mask) {
            for (u32 j = 1; j <= 9; ++j)
                if (t[i + j] == 0xFFFFu) { t[i + j] = static_cast<u16>(key); return; }
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & mask;
    }
}
B2C_HD void b2c_pyset_resize(B2cPySet& s, u32 minused) {
    u32 newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    u16* old = s.buf[s.cur];
    u16* nt = s.buf[s.cur ^ 1];
    b2c_pyset_clear(nt, newsize);
    for (u32 i = 0; i <= s.mask; ++i)
        if (old[i] != 0xFFFFu) b2c_pyset_insert_clean(nt, newsize - 1, old[i]);
    s.cur ^= 1;
    s.mask = newsize - 1;
}
B2C_HD void b2c_pyset_add(B2cPySet& s, u32 key) {
    u16* t = s.buf[s.cur];
    u32 perturb = key;
    u32 i = key & s.mask;
    while (true) {
        const u32 probes = (i + 9 <= s.mask) ? 9u : 0u;
        for (u32 j = 0; j <= probes; ++j) {
            const u16 e = t[i + j];
            if (e == 0xFFFFu) {
                t[i + j] = static_cast<u16>(key);
                ++s.fill;
                if (s.fill * 5 >= s.mask * 3) b2c_pyset_resize(s, s.fill > 50000 ? s.fill * 2 : s.fill * 4);
                return;
            }
            if (e == key) return;
        }
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & s.mask;
    }
}
// result = copy(s) | {extra}: set_copy (set_merge into an empty set) then set_merge of a 1-element set.
// On return s holds the result; iterate slots 0..mask ascending for the iteration order.
B2C_HD void b2c_pyset_copy_or(B2cPySet& s, u32 extra) {
    // --- copy: new empty table with mask 7 receives all of s
    const u32 used = s.fill;
    u32 nmask = 7;
    if (used > 0) {
        if ((0 + used) * 5 >= nmask * 3) {
            u32 newsize = 8;
            while (newsize <= used * 2) newsize <<= 1;
            nmask = newsize - 1;
        }
        if (nmask != s.mask) {
            // different size: set_insert_clean in slot order of the source
            u16* old = s.buf[s.cur];
            u16* nt = s.buf[s.cur ^ 1];
            b2c_pyset_clear(nt, nmask + 1);
            for (u32 i = 0; i <= s.mask; ++i)
                if (old[i] != 0xFFFFu) b2c_pyset_insert_clean(nt, nmask, old[i]);
            s.cur ^= 1;
            s.mask = nmask;
        }
        // same size: slots are copied verbatim -> nothing to do
    } else {
        b2c_pyset_clear(s.buf[s.cur], 8);
        s.mask = 7;
    }
    // --- merge {extra}
    if ((s.fill + 1) * 5 >= s.mask * 3) b2c_pyset_resize(s, (s.fill + 1) * 2);
    if (s.fill == 0) {
        // empty target, same mask (7) as the one-element source: slot copy
        s.buf[s.cur][extra & 7] = static_cast<u16>(extra);
        s.fill = 1;
        return;
    }
    b2c_pyset_add(s, extra);
}

// ---------------------------------------------------------------------------------------
// kernel bodies.  Three launches:
//   rowsum  -- numpy-order sum of every logit row (coalesced: 8 lanes per row, 4 rows per warp)
//   decide  -- per utterance: numpy-order mean of the row sums, math.isclose(mean, 1)
//   tokens  -- one warp per RUN of 8 consecutive frames of one utterance, persistent grid, no
//              block barrier: log-softmax, clip, token selection, CPython set order, compact lists
// Token lists are compact inside a run; the run starting at frame t0 of utterance u owns the
// entry range beginning at (frame_off[u] + t0) * V.
// ---------------------------------------------------------------------------------------
#define B2C_RUN 8                   // frames per run
#define B2C_PREP_WARPS 8
#define B2C_PREP_SMEM_SET 128       // entries per smem set buffer (enough for 32 selected tokens)
#define B2C_PREP_LEAF_CAP 1024
#define B2C_ROWSUM_MAX_LEAF 64      // rows up to 8192 elements take the coalesced warp-per-row sum

// token list of one frame: offset inside its run, length, and the FIRST token of the list inline (a single-token
// frame -- most frames of ASR posteriors -- is fully described by its 16-byte record: the beam kernel's in-place
// runs never touch the token arrays)
struct B2cFrameRec { u32 off; u16 cnt; u16 id0; double lp0; };

struct B2cPrepArgs {
    const void* logits;      // packed [total_frames, V]
    const u64* frame_off;    // [B]
    const int* T;            // [B]
    const u64* run_off;      // [B+1] exclusive prefix of ceil(T/8)
    int n_utts;
    u64 total_frames;
    int V;
    double token_min_logp;
    B2cFrameRec* tok_rec;    // [total_frames]
    u32* tok_ids;            // [total_frames * V]  (32-bit: the beam kernel stages them with 4-byte cp.async)
    double* tok_lp;
    void* rowsum;            // [total_frames] scratch, input dtype
    u16* set_scratch;        // [total warps][2][set_cap] spill space for large token sets
    u32 set_cap;             // power of two >= 8 * (V + 1)
    // leaves (<= 128 elements) of numpy's pairwise recursion over one row, in visiting order; n_leaf == 0:
    // not tabulated (V <= 128 needs no table, very long rows use the per-thread path)
    int n_leaf;
    u32 leaf_off[B2C_ROWSUM_MAX_LEAF];
    u32 leaf_n[B2C_ROWSUM_MAX_LEAF];
    int* is_prob;            // [B]
    u32* max_k;              // [B] largest per-frame token count (zeroed before the launch)
    u32* sum_k;              // [B] total number of selected tokens (zeroed before the launch)
};

// ---- rowsum ---------------------------------------------------------------------------------
template <class T>
struct B2cLeafShared {
    const double* sums;
    mutable int next;
    B2C_HD T operator()(long, long) const { return static_cast<T>(sums[next++]); }
};

template <class T>
B2C_HD void b2c_rowsum_block(const B2cPrepArgs& A, int block_idx, int n_blocks, double* leaf_sums) {
    const T* x = static_cast<const T*>(A.logits);
    T* rs = static_cast<T*>(A.rowsum);
    const int V = A.V;
#if defined(__CUDA_ARCH__)
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, j = lane & 7, g = lane >> 3;
    const u64 warp = static_cast<u64>(block_idx) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const u64 n_warps = static_cast<u64>(n_blocks) * (blockDim.x >> 5);
    if (V >= 8 && V <= 128) {
        // numpy's 8 accumulators are the 8 lanes of a group; xor-butterfly 1,2,4 reproduces
        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); the tail (n % 8) is added sequentially
        const int body = V - (V % 8);
        for (u64 r0 = warp * 4; r0 < A.total_frames; r0 += n_warps * 4) {
            const u64 r = r0 + g;
            const bool ok = r < A.total_frames;
            const T* a = x + (ok ? r : 0) * static_cast<u64>(V);
            T acc = a[j];
            for (int i = 8; i < body; i += 8) acc = acc + a[i + j];
            acc = acc + __shfl_xor_sync(full, acc, 1);
            acc = acc + __shfl_xor_sync(full, acc, 2);
            acc = acc + __shfl_xor_sync(full, acc, 4);
            for (int i = body; i < V; ++i) acc = acc + a[i];
            if (ok && j == 0) rs[r] = acc;
        }
        return;
    }
    if (A.n_leaf > 0) {
        // long rows (V > 128): one warp per row, four leaves at a time (8 lanes = numpy's 8 accumulators of a
        // leaf, 32-byte sectors fully used), then the recursion over the leaf sums by lane 0
        const int wib = static_cast<int>(threadIdx.x >> 5);
        for (u64 r = warp; r < A.total_frames; r += n_warps) {
            const T* a = x + r * static_cast<u64>(V);
            for (int l0 = 0; l0 < A.n_leaf; l0 += 4) {
                const int lf = l0 + g;
                const bool ok = lf < A.n_leaf;
                const u32 off = ok ? A.leaf_off[lf] : 0u;
                const int n = ok ? static_cast<int>(A.leaf_n[lf]) : 8;
                T res;
                if (n < 8) {
                    res = static_cast<T>(-0.0);
                    for (int i = 0; i < n; ++i) res = res + a[off + i];
                } else {
                    const int body = n - (n % 8);
                    T acc = a[off + j];
                    for (int i = 8; i < body; i += 8) acc = acc + a[off + i + j];
                    acc = acc + __shfl_xor_sync(full, acc, 1);
                    acc = acc + __shfl_xor_sync(full, acc, 2);
                    acc = acc + __shfl_xor_sync(full, acc, 4);
                    for (int i = body; i < n; ++i) acc = acc + a[off + i];
                    res = acc;
                }
                if (ok && j == 0) leaf_sums[wib * B2C_ROWSUM_MAX_LEAF + lf] = static_cast<double>(res);
            }
            __syncwarp();
            if (lane == 0) {
                B2cLeafShared<T> lfn;
                lfn.sums = leaf_sums + wib * B2C_ROWSUM_MAX_LEAF;
                lfn.next = 0;
                rs[r] = b2c_np_pairwise_generic<T>(V, lfn);
            }
            __syncwarp();
        }
        return;
    }
    // generic: one thread per row, numpy's recursion evaluated sequentially
    for (u64 r = warp * 32 + lane; r < A.total_frames; r += n_warps * 32) rs[r] = b2c_np_pairwise<T>(x + r * static_cast<u64>(V), V);
#else
    if (block_idx == 0)
        for (u64 r = 0; r < A.total_frames; ++r) rs[r] = b2c_np_pairwise<T>(x + r * static_cast<u64>(V), V);
    (void)n_blocks;
    (void)leaf_sums;
#endif
}

// ---- decide ---------------------------------------------------------------------------------
struct B2cDecideShared { double leaf_sum[B2C_PREP_LEAF_CAP]; };

B2C_HD int b2c_count_leaves(long n) {
    long st[48];
    int sp = 0, nl = 0;
    st[sp++] = n;
    while (sp > 0) {
        long q = st[--sp];
        if (q <= 128) { ++nl; continue; }
        long q2 = q / 2;
        q2 -= q2 % 8;
        st[sp++] = q - q2;
        st[sp++] = q2;
    }
    return nl;
}

template <class T>
B2C_HD void b2c_decide_block(const B2cPrepArgs& A, int u, B2cDecideShared* sh) {
    const int Tn = A.T[u];
    const T* rs = static_cast<const T*>(A.rowsum) + A.frame_off[u];
    const int n_leaf = b2c_count_leaves(Tn);
    const bool par_leaves = n_leaf <= B2C_PREP_LEAF_CAP && Tn > 128;
    if (par_leaves) {
        B2C_FOR(lf, n_leaf) {
            long off = 0, n = Tn;
            int idx = lf;
            while (n > 128) {
                long n2 = n / 2;
                n2 -= n2 % 8;
                const int nl = b2c_count_leaves(n2);
                if (idx < nl) { n = n2; } else { idx -= nl; off += n2; n = n - n2; }
            }
            sh->leaf_sum[lf] = static_cast<double>(b2c_np_leaf_sum<T>(rs + off, n));
        }
        B2C_SYNC();
    }
    B2C_LEADER {
        int isp = 0;
        if (Tn > 0) {
            T tot;
            if (par_leaves) {
                B2cLeafShared<T> lf;
                lf.sums = sh->leaf_sum;
                lf.next = 0;
                tot = b2c_np_pairwise_generic<T>(Tn, lf);
            } else {
                tot = b2c_np_pairwise<T>(rs, Tn);
            }
            const T mean = tot / static_cast<T>(Tn);
            isp = b2c_isclose_one(static_cast<double>(mean)) ? 1 : 0;
        }
        A.is_prob[u] = isp;
    }
}

// ---- tokens ---------------------------------------------------------------------------------
struct B2cPrepShared { u16 sets[B2C_PREP_WARPS][2][B2C_PREP_SMEM_SET]; };

// row statistics: max, log-sum-exp, argmax of the clipped log-probs, and the selected set
// fed in ascending order into `set`.  Warp-cooperative on the device, a plain loop in hostsim.
// kMaskOnly (V <= 32): the selected set is returned as a bit mask in nsel_out instead of being inserted
template <class T, bool kMaskOnly>
B2C_HD void b2c_prep_row(const T* row, int V, bool is_prob, double thr, int lane, B2cPySet& set, T& m_out,
                         double& ls_out, int& amax_out, u32& nsel_out) {
#if defined(__CUDA_ARCH__)
    const unsigned full = 0xFFFFFFFFu;
    T m = static_cast<T>(0);
    double ls = 0.0;
    if (!is_prob) {
        bool have = false;
        for (int v = lane; v < V; v += 32) {
            const T x = row[v];
            if (!have || x > m) { m = x; have = true; }
        }
        for (int off = 16; off >= 1; off >>= 1) {
            const T o = __shfl_xor_sync(full, m, off);
            const int oh = __shfl_xor_sync(full, have ? 1 : 0, off);
            if (oh && (!have || o > m)) { m = o; have = true; }
        }
        if (!(static_cast<double>(m) - static_cast<double>(m) == 0.0)) m = static_cast<T>(0);
        double part = 0.0;
        for (int v = lane; v < V; v += 32) part += exp(static_cast<double>(static_cast<T>(row[v] - m)));
        for (int off = 16; off >= 1; off >>= 1) part = part + __shfl_xor_sync(full, part, off);
        ls = log(part);
    }
    double best = 0.0;
    int besti = -1;
    u32 nsel = 0;
    if (!kMaskOnly && lane == 0) b2c_pyset_init(set, set.buf[0], set.buf[1]);
    for (int base = 0; base < V; base += 32) {
        const int v = base + lane;
        bool sel = false;
        if (v < V) {
            const double lp = b2c_lp<T>(row[v], is_prob, m, ls);
            if (besti < 0 || lp > best) { best = lp; besti = v; }
            sel = lp >= thr;
        }
        unsigned mask = __ballot_sync(full, sel);
        if (kMaskOnly) { nsel = mask; continue; }
        nsel += __popc(mask);
        if (lane == 0) {
            while (mask) {
                const int bit = __ffs(mask) - 1;
                mask &= mask - 1;
                b2c_pyset_add(set, static_cast<u32>(base + bit));
            }
        }
        __syncwarp();
    }
    for (int off = 16; off >= 1; off >>= 1) {
        const double ob = __shfl_xor_sync(full, best, off);
        const int oi = __shfl_xor_sync(full, besti, off);
        if (oi >= 0 && (besti < 0 || ob > best || (ob == best && oi < besti))) { best = ob; besti = oi; }
    }
    m_out = m; ls_out = ls; amax_out = besti; nsel_out = nsel;
#else
    (void)lane;
    T m = static_cast<T>(0);
    double ls = 0.0;
    if (!is_prob) {
        m = row[0];
        for (int v = 1; v < V; ++v) if (row[v] > m) m = row[v];
        if (!(static_cast<double>(m) - static_cast<double>(m) == 0.0)) m = static_cast<T>(0);
        double part[32];
        for (int l = 0; l < 32; ++l) {
            double s = 0.0;
            for (int v = l; v < V; v += 32) s += exp(static_cast<double>(static_cast<T>(row[v] - m)));
            part[l] = s;
        }
        for (int off = 16; off >= 1; off >>= 1) {
            double nxt[32];
            for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ off];
            for (int l = 0; l < 32; ++l) part[l] = nxt[l];
        }
        ls = log(part[0]);
    }
    double best = 0.0;
    int besti = -1;
    u32 nsel = 0;
    if (!kMaskOnly) b2c_pyset_init(set, set.buf[0], set.buf[1]);
    for (int v = 0; v < V; ++v) {
        const double lp = b2c_lp<T>(row[v], is_prob, m, ls);
        if (besti < 0 || lp > best) { best = lp; besti = v; }
        if (lp >= thr) {
            if (kMaskOnly) nsel |= 1u << v;
            else { ++nsel; b2c_pyset_add(set, static_cast<u32>(v)); }
        }
    }
    m_out = m; ls_out = ls; amax_out = besti; nsel_out = nsel;
#endif
}

// CPython iteration order of set(ascending ints of `mask`) | {amax} for at most 3 selected tokens < 32 with
// amax among them (or none selected): the table keeps its initial 8 slots (no resize below 5 entries, the
// linear-probe window i+9 <= mask is never open at mask 7), one byte per slot in a 64-bit register.
// Returns the number of tokens written to out[0..2].
B2C_HD u32 b2c_pyset_small_order(u32 mask, u32 amax, u32* out) {
    if (mask == 0) { out[0] = amax; return 1; }
    u64 tab = ~0ull;
    u32 mm = mask;
    while (mm) {
#if defined(__CUDA_ARCH__)
        const u32 key = static_cast<u32>(__ffs(static_cast<int>(mm)) - 1);
#else
        u32 key = 0;
        while (!((mm >> key) & 1u)) ++key;
#endif
        mm &= mm - 1;
        u32 perturb = key, i = key & 7u;
        while (((tab >> (8 * i)) & 0xFFull) != 0xFFull) {
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & 7u;
        }
        tab = (tab & ~(0xFFull << (8 * i))) | (static_cast<u64>(key) << (8 * i));
    }
    u32 n = 0;
    for (u32 i = 0; i < 8; ++i) {
        const u32 e = static_cast<u32>((tab >> (8 * i)) & 0xFFull);
        if (e != 0xFFu) out[n++] = e;
    }
    return n;
}

// V <= 32: one run (<= 8 frames) per warp.  Row statistics are warp-wide per frame and leave the frame's
// selected set as a bit mask in lane f; the ordering of the (typically 1-3) tokens, the offsets inside the
// run and the compact writes are then done by 8 lanes in parallel, one frame each.  Frames with more than
// 3 selected tokens go through the general set emulation afterwards, at offsets that are already known.
template <class T>
B2C_HD void b2c_tokens_run_v32(const B2cPrepArgs& A, u64 run, int lane, u16* set0, u16* set1) {
    int lo = 0, hi = A.n_utts - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (A.run_off[mid] <= run) lo = mid; else hi = mid - 1;
    }
    const int u = lo;
    const int Tn = A.T[u];
    const int t0 = static_cast<int>(run - A.run_off[u]) * B2C_RUN;
    const int t1 = t0 + B2C_RUN < Tn ? t0 + B2C_RUN : Tn;
    const int nf = t1 - t0;
    const int V = A.V;
    const u64 f0 = A.frame_off[u];
    const bool is_prob = A.is_prob[u] != 0;
    const T* x = static_cast<const T*>(A.logits) + f0 * static_cast<u64>(V);
    const u64 base = (f0 + static_cast<u64>(t0)) * static_cast<u64>(V);
    u32* ids = A.tok_ids + base;
    double* lps = A.tok_lp + base;
    B2cPySet set;
    set.buf[0] = set0;
    set.buf[1] = set1;
#if defined(__CUDA_ARCH__)
    const unsigned full = 0xFFFFFFFFu;
    u32 my_mask = 0, my_amax = 0;
    T my_m = static_cast<T>(0);
    double my_ls = 0.0;
    if (sizeof(T) == 4 && !is_prob) {
        // float32 logits (the common case), same arithmetic as b2c_prep_row with far fewer instructions:
        // row max and argmax through REDUX on order-preserving integer keys, log(sum) once per RUN (lane f
        // takes frame f), the 8 rows of the run held in registers
        const bool has = lane < V;
        float d[B2C_RUN];
        double my_S = 1.0;
#pragma unroll
        for (int f = 0; f < B2C_RUN; ++f)      // all rows of the run in flight before the first reduction
            d[f] = (has && f < nf) ? static_cast<float>(x[static_cast<u64>(t0 + f) * V + lane]) : 0.0f;
#pragma unroll
        for (int f = 0; f < B2C_RUN; ++f) {
            if (f < nf) {
                const float xv = d[f];
                // max over the row: int order == float order on these keys (NaN sorts above +inf -> non-finite -> 0)
                const u32 xb = __float_as_uint(xv);
                const u32 key = has ? ((xb & 0x80000000u) ? ~xb : (xb | 0x80000000u)) : 0u;
                const u32 km = __reduce_max_sync(full, key);
                float m = __uint_as_float((km & 0x80000000u) ? (km & 0x7FFFFFFFu) : ~km);
                if (!(static_cast<double>(m) - static_cast<double>(m) == 0.0)) m = 0.0f;
                d[f] = xv - m;
                double part = has ? exp(static_cast<double>(d[f])) : 0.0;
                for (int off = 16; off >= 1; off >>= 1) part = part + __shfl_xor_sync(full, part, off);
                if (lane == f) { my_S = part; my_m = static_cast<T>(m); }
            }
        }
        my_ls = log(my_S);
#pragma unroll
        for (int f = 0; f < B2C_RUN; ++f) {
            if (f < nf) {
                const double ls = __shfl_sync(full, my_ls, f);
                double lp = static_cast<double>(static_cast<float>(static_cast<double>(d[f]) - ls));
                if (lp < B2C_LOG_MIN_CLIP) lp = B2C_LOG_MIN_CLIP;
                if (lp > 0.0) lp = 0.0;
                const u32 mask = __ballot_sync(full, has && lp >= A.token_min_logp);
                // argmax like the sequential scan: the first element wins if it is NaN, otherwise the largest
                // non-NaN value, lowest index among equals
                const float lpf = static_cast<float>(lp);
                const u32 lb = __float_as_uint(lpf);
                const u32 lkey = (has && lpf == lpf) ? ((lb & 0x80000000u) ? ~lb : (lb | 0x80000000u)) : 0u;
                const u32 lmax = __reduce_max_sync(full, lkey);
                const u32 eq = __ballot_sync(full, has && lkey == lmax);
                const u32 first_nan = __ballot_sync(full, lane == 0 && !(lpf == lpf));
                const u32 amax = (first_nan || eq == 0) ? 0u : static_cast<u32>(__ffs(static_cast<int>(eq)) - 1);
                if (lane == f) { my_mask = mask; my_amax = amax; }
            }
        }
    } else {
        for (int f = 0; f < nf; ++f) {
            T m;
            double ls;
            int amax;
            u32 mask;
            b2c_prep_row<T, true>(x + static_cast<u64>(t0 + f) * V, V, is_prob, A.token_min_logp, lane, set, m, ls, amax, mask);
            amax = __shfl_sync(full, amax, 0);          // lane 0's view, like the general path
            if (lane == f) { my_mask = mask; my_amax = static_cast<u32>(amax); my_m = m; my_ls = ls; }
        }
    }
    const bool mine = lane < nf;
    const u32 nsel = __popc(my_mask);
    const bool small = nsel == 0 || (nsel <= 3 && ((my_mask >> my_amax) & 1u));
    const u32 cnt = mine ? static_cast<u32>(__popc(my_mask | (1u << my_amax))) : 0u;
    u32 incl = cnt;
    for (int off = 1; off < B2C_RUN; off <<= 1) {
        const u32 o = __shfl_up_sync(full, incl, off);
        if (lane >= off) incl += o;
    }
    const u32 my_off = incl - cnt;
    if (mine) {
        B2cFrameRec rec;
        rec.off = my_off;
        rec.cnt = static_cast<u16>(cnt);
        rec.id0 = 0;
        rec.lp0 = 0.0;
        if (small) {
            const T* row = x + static_cast<u64>(t0 + lane) * V;
            u32 toks[3];
            const u32 n = b2c_pyset_small_order(my_mask, my_amax, toks);
            for (u32 q = 0; q < n; ++q) {
                const double lp = b2c_lp<T>(row[toks[q]], is_prob, my_m, my_ls);
                ids[my_off + q] = toks[q];
                lps[my_off + q] = lp;
                if (q == 0) { rec.id0 = static_cast<u16>(toks[0]); rec.lp0 = lp; }
            }
        }
        A.tok_rec[f0 + static_cast<u64>(t0 + lane)] = rec;     // frames of the general set emulation: first token patched below
    }
    u32 big = __ballot_sync(full, mine && !small);
    while (big) {
        const int f = __ffs(static_cast<int>(big)) - 1;
        big &= big - 1;
        const T* row = x + static_cast<u64>(t0 + f) * V;
        T m;
        double ls;
        int amax;
        u32 ns;
        b2c_prep_row<T, false>(row, V, is_prob, A.token_min_logp, lane, set, m, ls, amax, ns);
        u32 off = __shfl_sync(full, my_off, f);
        if (lane == 0) {
            b2c_pyset_copy_or(set, static_cast<u32>(amax));
            const u16* tab = set.buf[set.cur];
            bool first = true;
            for (u32 s = 0; s <= set.mask; ++s) {
                const u16 tok = tab[s];
                if (tok == 0xFFFFu) continue;
                const double lp = b2c_lp<T>(row[tok], is_prob, m, ls);
                ids[off] = tok;
                lps[off] = lp;
                if (first) {
                    B2cFrameRec* r = A.tok_rec + f0 + static_cast<u64>(t0 + f);
                    r->id0 = tok;
                    r->lp0 = lp;
                    first = false;
                }
                ++off;
            }
        }
        __syncwarp();
    }
    u32 mx = cnt;
    for (int off = 1; off < B2C_RUN; off <<= 1) {
        const u32 o = __shfl_xor_sync(full, mx, off);
        mx = o > mx ? o : mx;
    }
    const u32 total = __shfl_sync(full, incl, B2C_RUN - 1);
    if (lane == 0 && mx > 0) {
        b2c_atomic_max_u32(&A.max_k[u], mx);
        b2c_atomic_add_u32(&A.sum_k[u], total);
    }
#else
    (void)lane;
    u32 off = 0, mx = 0;
    for (int f = 0; f < nf; ++f) {
        const T* row = x + static_cast<u64>(t0 + f) * V;
        T m;
        double ls;
        int amax;
        u32 mask;
        b2c_prep_row<T, true>(row, V, is_prob, A.token_min_logp, 0, set, m, ls, amax, mask);
        u32 nsel = 0;
        for (u32 b = 0; b < 32; ++b) nsel += (mask >> b) & 1u;
        const bool small = nsel == 0 || (nsel <= 3 && ((mask >> amax) & 1u));
        B2cFrameRec rec;
        rec.off = off;
        rec.id0 = 0;
        rec.lp0 = 0.0;
        if (small) {
            u32 toks[3];
            const u32 n = b2c_pyset_small_order(mask, static_cast<u32>(amax), toks);
            for (u32 q = 0; q < n; ++q) {
                ids[off] = toks[q];
                lps[off] = b2c_lp<T>(row[toks[q]], is_prob, m, ls);
                if (q == 0) { rec.id0 = static_cast<u16>(toks[0]); rec.lp0 = lps[off]; }
                ++off;
            }
            rec.cnt = static_cast<u16>(n);
        } else {
            u32 ns;
            b2c_prep_row<T, false>(row, V, is_prob, A.token_min_logp, 0, set, m, ls, amax, ns);
            b2c_pyset_copy_or(set, static_cast<u32>(amax));
            rec.cnt = static_cast<u16>(set.fill);
            const u16* tab = set.buf[set.cur];
            bool first = true;
            for (u32 s2 = 0; s2 <= set.mask; ++s2) {
                const u16 tok = tab[s2];
                if (tok == 0xFFFFu) continue;
                ids[off] = tok;
                lps[off] = b2c_lp<T>(row[tok], is_prob, m, ls);
                if (first) { rec.id0 = tok; rec.lp0 = lps[off]; first = false; }
                ++off;
            }
        }
        A.tok_rec[f0 + static_cast<u64>(t0 + f)] = rec;
        if (rec.cnt > mx) mx = rec.cnt;
    }
    if (mx > 0) {
        b2c_atomic_max_u32(&A.max_k[u], mx);
        b2c_atomic_add_u32(&A.sum_k[u], off);
    }
#endif
}

// one run (<= 8 frames) handled by one warp
template <class T>
B2C_HD void b2c_tokens_run(const B2cPrepArgs& A, u64 run, int lane, u16* set0, u16* set1) {
    // locate the utterance of this run (binary search over the prefix of runs per utterance)
    int lo = 0, hi = A.n_utts - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (A.run_off[mid] <= run) lo = mid; else hi = mid - 1;
    }
    const int u = lo;
    const int Tn = A.T[u];
    const int t0 = static_cast<int>(run - A.run_off[u]) * B2C_RUN;
    const int t1 = t0 + B2C_RUN < Tn ? t0 + B2C_RUN : Tn;
    const int V = A.V;
    const u64 f0 = A.frame_off[u];
    const bool is_prob = A.is_prob[u] != 0;
    const T* x = static_cast<const T*>(A.logits) + f0 * static_cast<u64>(V);
    const u64 base = (f0 + static_cast<u64>(t0)) * static_cast<u64>(V);
    u32* ids = A.tok_ids + base;
    double* lps = A.tok_lp + base;
    u32 off = 0, mx = 0;
    for (int t = t0; t < t1; ++t) {
        const T* row = x + static_cast<u64>(t) * V;
        B2cPySet set;
        set.buf[0] = set0;
        set.buf[1] = set1;
        T m;
        double ls;
        int amax;
        u32 nsel;
        b2c_prep_row<T, false>(row, V, is_prob, A.token_min_logp, lane, set, m, ls, amax, nsel);
        if (lane == 0) {
            b2c_pyset_copy_or(set, static_cast<u32>(amax));
            const u32 cnt = set.fill;
            B2cFrameRec rec;
            rec.off = off;
            rec.cnt = static_cast<u16>(cnt);
            rec.id0 = 0;
            rec.lp0 = 0.0;
            const u16* tab = set.buf[set.cur];
            bool first = true;
            for (u32 s = 0; s <= set.mask; ++s) {
                const u16 tok = tab[s];
                if (tok == 0xFFFFu) continue;
                ids[off] = tok;
                lps[off] = b2c_lp<T>(row[tok], is_prob, m, ls);
                if (first) { rec.id0 = tok; rec.lp0 = lps[off]; first = false; }
                ++off;
            }
            A.tok_rec[f0 + static_cast<u64>(t)] = rec;
            if (cnt > mx) mx = cnt;
        }
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
    }
    if (lane == 0 && mx > 0) {
        b2c_atomic_max_u32(&A.max_k[u], mx);
        b2c_atomic_add_u32(&A.sum_k[u], off);
    }
}

template <class T>
B2C_HD void b2c_tokens_block(const B2cPrepArgs& A, int block_idx, int n_blocks, B2cPrepShared* sh) {
    const u64 total_runs = A.run_off[A.n_utts];
    const bool small = A.V <= 32;
#if defined(__CUDA_ARCH__)
    const int w = static_cast<int>(threadIdx.x >> 5), lane = static_cast<int>(threadIdx.x & 31);
    const u64 gw = static_cast<u64>(block_idx) * B2C_PREP_WARPS + w;
    const u64 n_warps = static_cast<u64>(n_blocks) * B2C_PREP_WARPS;
    u16* b0 = small ? sh->sets[w][0] : A.set_scratch + (gw * 2 + 0) * A.set_cap;
    u16* b1 = small ? sh->sets[w][1] : A.set_scratch + (gw * 2 + 1) * A.set_cap;
    if (small) { for (u64 run = gw; run < total_runs; run += n_warps) b2c_tokens_run_v32<T>(A, run, lane, b0, b1); }
    else { for (u64 run = gw; run < total_runs; run += n_warps) b2c_tokens_run<T>(A, run, lane, b0, b1); }
#else
    for (int w = 0; w < B2C_PREP_WARPS; ++w) {
        const u64 gw = static_cast<u64>(block_idx) * B2C_PREP_WARPS + w;
        const u64 n_warps = static_cast<u64>(n_blocks) * B2C_PREP_WARPS;
        u16* b0 = small ? sh->sets[w][0] : A.set_scratch + (gw * 2 + 0) * A.set_cap;
        u16* b1 = small ? sh->sets[w][1] : A.set_scratch + (gw * 2 + 1) * A.set_cap;
        for (u64 run = gw; run < total_runs; run += n_warps) {
            if (small) b2c_tokens_run_v32<T>(A, run, 0, b0, b1);
            else b2c_tokens_run<T>(A, run, 0, b0, b1);
        }
    }
#endif
}
