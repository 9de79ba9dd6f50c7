// b200ctc -- the per-utterance prefix beam search (one CTA per utterance, all T frames).
//
// Restates, string-free, the body of BeamSearchDecoderCTC._partial_decode_logits
// (reference decoder.py:443-554) and _finalize_beams/_decode_logits (:558-667):
//   expand (4 branches, :452-534) -> merge by (text, partial_word, last_char) with
//   log-sum-exp in iteration order (:211-224) -> LM / hotword fusion (:346-424) ->
//   score threshold (:545-546) -> stable top-N (:165-167) -> history prune (:227-258).
//
// Order dependence of the reference that is reproduced exactly:
//   * candidates are enumerated token-major in the CPython set order of the frame's tokens
//     (computed by the prepare kernel), beams in rank order inside a token;
//   * a merged group sits at the position of its FIRST member, takes the metadata (frames,
//     text/next_word split) of its LAST member, and folds scores left to right;
//   * ties in lm_score keep enumeration order (heapq.nlargest is stable);
//   * the single force_next_break flag of BPE alphabets (:442,474-482).
//
// Phases are separated by block barriers; see b2c_cta.h for the execution-model macros.
#pragma once
#include "b2c_cta.h"
#include "b2c_lm.h"

struct B2cBeamTab {
    double* logit;      // logit_score
    double* lm_hw;      // LM + hotword score of the beam's text (reference cached lm_hw_score)
    double* pscore;     // score of the unfinished word (0 when empty)
    u64* text_hash;
    u64* part_hash;
    u64* hist_hash;     // hash of the last hist_n finished words (history-prune key)
    u32* text_node;     // -> B2cText arena
    u32* chain;         // -> B2cChain arena (backtrack)
    int* pf_s;          // partial_frames
    int* pf_e;
    u16* last_tok;      // canonical token id or B2C_NO_TOK
    u16* part_len;      // python len(partial_word)
};

struct B2cScalars {
    u64 max_key;
    u32 n_beams, n_surv, n_new, chain_used, text_used, status, force_break, n_sel;
};

#define B2C_NBUCKET 256      // score buckets of the O(m) ranking (monotone in the score)

struct B2cCandTier {     // per-frame candidate working set (shared memory tier or HBM tier)
    u32 cap;             // candidates
    u32 ht_cap;          // hash slots (power of two >= 2*cap)
    u64* ckey;           // merge key; after fusion: order-preserving lm_score key of group leaders, 0 otherwise
    double* cfold;       // merged logit_score of group leaders
    u32* cslot;          // candidate -> group slot
    u32* cnext;          // bucket list link (leaders that survive the score threshold)
    u32* clast;          // leader -> last member of its group (metadata donor)
    u32* ht_idx;         // slot -> representative candidate
    u32* ht_min;         // slot -> first member (dict position)
    u32* ht_max;         // slot -> last member
    u32* ht_cnt;
};

struct B2cWork {
    B2cScalars* sc;
    B2cBeamTab cur, nxt;
    B2cCandTier tier_s, tier_g;
    // selection (capacity beam_width)
    u32* ord;            // rank -> candidate index
    u64* phk;            // rank -> history-prune key
    u32* pslot;          // rank -> slot in the prune table
    u32* newidx;         // new beam -> rank
    u32* pt_idx;         // history-prune table: slot -> representative rank
    u32* pt_min;         //                      slot -> best rank with that key
    u32 pt_cap;          // power of two >= 2 * beam_width
    u32* bcnt;           // [B2C_NBUCKET] survivors per score bucket, then exclusive prefix
    u32* bhead;          // [B2C_NBUCKET] list heads
    // per-frame token side arrays for BPE force_next_break (capacity V, HBM)
    u32* tk_ffirst;
    u8* tk_fall;
    // arenas (HBM)
    B2cChain* chain;
    u32 chain_cap;
    B2cText* text;
    u32 text_cap;
#if defined(B2C_PHASE_CLOCKS)
    u64 clk[16];
    long long clk_last;
#endif
};

struct B2cExp {
    int type;            // 0 blank/repeat, 1 BPE word start, 2 space, 3 continuation
    u32 canon;
    u64 text_hash;
    u64 part_hash;
    u32 part_len;
    u64 word_hash;       // finished word (types 1,2) -- valid when word_len > 0
    u32 word_len;
    int pf_s, pf_e;
};

B2C_HD void b2c_swap_tabs(B2cBeamTab& a, B2cBeamTab& b) {
    B2cBeamTab t = a;
    a = b;
    b = t;
}

// one (token, beam) pair of the reference's double loop (decoder.py:447-534)
B2C_HD void b2c_expand(const B2cParams& P, const B2cBeamTab& cur, int b, u32 tok, bool forced, int t, B2cExp& e) {
    const B2cTok ti = P.toks[tok];
    e.canon = ti.canon;
    const u32 plen = cur.part_len[b];
    const u64 ph = cur.part_hash[b];
    e.word_len = 0;
    e.word_hash = 0;
    if ((ti.flags & B2C_TF_BLANK) || cur.last_tok[b] == ti.canon) {                  // (i)
        e.type = 0;
        e.text_hash = cur.text_hash[b];
        e.part_hash = ph;
        e.part_len = plen;
        e.pf_s = cur.pf_s[b];
        e.pf_e = (ti.flags & B2C_TF_BLANK) ? cur.pf_e[b] : t + 1;
    } else if (P.is_bpe && ((ti.flags & B2C_TF_BPE_LEAD) || forced)) {                // (ii)
        e.type = 1;
        e.text_hash = plen ? b2c_text_append(cur.text_hash[b], ph) : cur.text_hash[b];
        e.word_hash = ph;
        e.word_len = plen;
        e.part_hash = ti.clean_hash;
        e.part_len = ti.clean_nchars;
        e.pf_s = t;
        e.pf_e = t + 1;
    } else if (!P.is_bpe && (ti.flags & B2C_TF_SPACE)) {                              // (iii)
        e.type = 2;
        e.text_hash = plen ? b2c_text_append(cur.text_hash[b], ph) : cur.text_hash[b];
        e.word_hash = ph;
        e.word_len = plen;
        e.part_hash = 0;
        e.part_len = 0;
        e.pf_s = -1;
        e.pf_e = -1;
    } else {                                                                          // (iv)
        e.type = 3;
        e.text_hash = cur.text_hash[b];
        e.part_hash = b2c_hash_append(ph, ti.raw_hash, ti.raw_pow);
        e.part_len = plen + ti.raw_nchars;
        e.pf_s = cur.pf_s[b] < 0 ? t : cur.pf_s[b];
        e.pf_e = t + 1;
    }
}

// text-level quantities of "text + word" (reference _get_lm_beams cache miss, decoder.py:388-395)
struct B2cTextNew {
    double raw_lm, lm_hw;
    u32 hw_count;
    B2cLmState st;
};
B2C_HD void b2c_text_extend(const B2cParams& P, const B2cText& parent, u64 word_hash, u32 word_len, bool is_eos,
                            B2cTextNew& out) {
    out.hw_count = parent.hw_count + b2c_hot_is_word(P, word_hash, word_len);
    if (P.lm.order > 0) {
        double sc = b2c_lm_score_word(P, parent.st, word_hash, word_len, is_eos, out.st);
        out.raw_lm = parent.raw_lm + sc;
        out.lm_hw = out.raw_lm + P.hot_weight * static_cast<double>(out.hw_count);
    } else {
        out.raw_lm = 0.0;
        out.st.length = 0;
        out.lm_hw = P.hot_weight * static_cast<double>(out.hw_count);
    }
}

B2C_HD double b2c_combine_score(const B2cParams& P, double logit, double lm_hw, double pscore, u32 part_len) {
    double s;
    if (P.lm.order > 0) {
        double l = lm_hw;                                  // decoder.py:396-420
        if (part_len > 0) l += pscore;
        s = logit + l;
    } else {
        s = logit + lm_hw + pscore;                        // decoder.py:363-367
    }
    return s + 0.0;                                        // -0.0 -> +0.0 so that key order == float order
}

B2C_HD B2cCandTier b2c_pick_tier(const B2cWork& W, u32 M) {
    B2cCandTier c = W.tier_s;
    if (M > W.tier_s.cap) c = W.tier_g;
    return c;
}

B2C_HD u32 b2c_ht_size(u32 M) {
    u32 h = 16;
    while (h < 2 * M) h <<= 1;
    return h;
}

B2C_HD void b2c_fence_block() {
#if defined(__CUDA_ARCH__)
    __threadfence_block();
#endif
}

// group candidates with equal keys.  The caller has stored C.ckey[i] and issued a block fence;
// a thread that loses the slot race reads the winner's key, which the winner published
// (store, fence) before its CAS.
B2C_HD void b2c_group_insert(const B2cCandTier& C, u32 hmask, u32 i, u64 key) {
    u32 slot = static_cast<u32>(b2c_mix64(key)) & hmask;
    while (true) {
        const u32 rep = b2c_atomic_cas_u32(&C.ht_idx[slot], B2C_NONE_U32, i);
        if (rep == B2C_NONE_U32) break;
        b2c_fence_block();
        if (C.ckey[rep] == key) break;
        slot = (slot + 1) & hmask;
    }
    C.cslot[i] = slot;
    b2c_atomic_min_u32(&C.ht_min[slot], i);
    b2c_atomic_max_u32(&C.ht_max[slot], i);
    b2c_atomic_add_u32(&C.ht_cnt[slot], 1u);
}

// clear the grouping table (first H slots), the score buckets and the history-prune table for the
// next use; called in a phase where none of them is read any more
B2C_HD void b2c_clear_tables(const B2cWork& W, const B2cCandTier& C, u32 H) {
    B2C_FOR(s, H) {
        C.ht_idx[s] = B2C_NONE_U32;
        C.ht_min[s] = B2C_NONE_U32;
        C.ht_max[s] = 0;
        C.ht_cnt[s] = 0;
    }
    B2C_FOR(s, B2C_NBUCKET) { W.bcnt[s] = 0; W.bhead[s] = B2C_NONE_U32; }
    B2C_FOR(s, W.pt_cap) { W.pt_idx[s] = B2C_NONE_U32; W.pt_min[s] = B2C_NONE_U32; }
}

// i -> (i / n, i % n) without an integer division (float reciprocal + exact correction)
B2C_HD void b2c_divmod(u32 i, u32 n, float rcp, u32& q, u32& r) {
    if (i >= (1u << 22)) { q = i / n; r = i - q * n; return; }
    q = static_cast<u32>(static_cast<float>(i) * rcp);
    r = i - q * n;
    if (static_cast<int>(r) < 0) { --q; r += n; }
    else if (r >= n) { ++q; r -= n; }
}

// score bucket, monotone non-increasing in the score: a larger score never gets a larger bucket
B2C_HD u32 b2c_bucket(double max_score, double score, double scale) {
    const double d = (max_score - score) * scale;
    u32 b = d >= static_cast<double>(B2C_NBUCKET - 1) ? static_cast<u32>(B2C_NBUCKET - 1) : static_cast<u32>(d);
    return b;
}
B2C_HD double b2c_bucket_scale(double prune_logp) {
    double range = -prune_logp;
    if (!(range >= 1.0)) range = 1.0;      // also catches NaN
    if (range > 32.0) range = 32.0;
    return static_cast<double>(B2C_NBUCKET) / range;
}

// exclusive prefix over the bucket counts (one warp), total -> *n_total
B2C_HD void b2c_bucket_scan(u32* bcnt, u32* n_total) {
#if defined(__CUDA_ARCH__)
    if (threadIdx.x < 32) {
        const u32 lane = threadIdx.x;
        const u32 per = B2C_NBUCKET / 32;
        u32 v[B2C_NBUCKET / 32];
        u32 sum = 0;
#pragma unroll
        for (u32 q = 0; q < per; ++q) { v[q] = bcnt[lane * per + q]; sum += v[q]; }
        u32 incl = sum;
        for (int off = 1; off < 32; off <<= 1) {
            const u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= static_cast<u32>(off)) incl += o;
        }
        u32 run = incl - sum;
#pragma unroll
        for (u32 q = 0; q < per; ++q) { bcnt[lane * per + q] = run; run += v[q]; }
        if (lane == 31) *n_total = incl;
    }
#else
    u32 run = 0;
    for (u32 b = 0; b < B2C_NBUCKET; ++b) { const u32 c = bcnt[b]; bcnt[b] = run; run += c; }
    *n_total = run;
#endif
}

// compaction of the ranks that survive the history prune: newidx[pos] = rank, ascending
B2C_HD void b2c_compact_kept(const u32* pslot, const u32* pt_min, u32* newidx, u32* n_new, u32 nsel) {
#if defined(__CUDA_ARCH__)
    if (threadIdx.x < 32) {
        const u32 lane = threadIdx.x;
        const u32 per = (nsel + 31) >> 5;
        const u32 beg = lane * per;
        const u32 end = beg + per < nsel ? beg + per : nsel;
        u32 cnt = 0;
        for (u32 r = beg; r < end; ++r) cnt += (pt_min[pslot[r]] == r) ? 1u : 0u;
        u32 incl = cnt;
        for (int off = 1; off < 32; off <<= 1) {
            const u32 v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (lane >= static_cast<u32>(off)) incl += v;
        }
        u32 pos = incl - cnt;
        for (u32 r = beg; r < end; ++r)
            if (pt_min[pslot[r]] == r) newidx[pos++] = r;
        if (lane == 31) *n_new = incl;
    }
#else
    u32 pos = 0;
    for (u32 r = 0; r < nsel; ++r)
        if (pt_min[pslot[r]] == r) newidx[pos++] = r;
    *n_new = pos;
#endif
}

// opt-in phase timing (-DB2C_PHASE_CLOCKS, profiling builds only): thread 0 accumulates the cycles
// between consecutive marks into W.clk[]
#if defined(B2C_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
#define B2C_MARK(idx) do { if (threadIdx.x == 0) { const long long _c = clock64(); W.clk[idx] += static_cast<u64>(_c - W.clk_last); W.clk_last = _c; } } while (0)
#else
#define B2C_MARK(idx) ((void)0)
#endif

// -----------------------------------------------------------------------------------------
// one frame.  Barriers: fuse(keys+group) | fold+score | threshold+bucket | scan | rank(+history
// keys) | compaction | commit(+clear for the next frame)  -> 7 (6 without history pruning).
// kFast: the candidate tier is the shared-memory one; all table views are value copies so that the
// compiler keeps them in registers and can prove the shared-memory address space of every access.
// On entry the grouping table (first ht_size(n*K) slots), the buckets and the prune table are clear.
// -----------------------------------------------------------------------------------------
template <bool kFast>
B2C_HD void b2c_frame_step(const B2cParams& P, B2cWork& W, int t, const u16* tk_id, const double* tk_lp, int K, int K_next) {
    B2cScalars* sc = W.sc;
    const u32 n = sc->n_beams;
    const u32 M = n * static_cast<u32>(K);
    const float rcp_n = 1.0f / static_cast<float>(n);
    const B2cCandTier C = kFast ? W.tier_s : b2c_pick_tier(W, M);
    if (M > C.cap) {  // cannot happen: the host sizes the tiers from beam_width and the token counts
        B2C_LEADER { sc->status = B2C_ERR_CAND_FULL; }
        B2C_SYNC();
        return;
    }
    B2C_MARK(0);
    const u32 hmask = b2c_ht_size(M) - 1;
    const B2cBeamTab cur = W.cur;
    const B2cBeamTab nx = W.nxt;
    u32* const ord = W.ord;
    u64* const phk = W.phk;
    u32* const pslot = W.pslot;
    u32* const newidx = W.newidx;
    u32* const bcnt = W.bcnt;
    u32* const bhead = W.bhead;
    u32* const pt_idx = W.pt_idx;
    u32* const pt_min = W.pt_min;
    const u32 ptmask = W.pt_cap - 1;

    // ---- phase 0 (BPE only): who consumes force_next_break -------------------------------
    if (P.is_bpe) {
        B2C_FOR(k, K) {
            const B2cTok ti = P.toks[tk_id[k]];
            u32 first = B2C_NONE_U32;
            if (!(ti.flags & B2C_TF_BLANK)) {
                for (u32 b = 0; b < n; ++b)
                    if (cur.last_tok[b] != ti.canon) { first = b; break; }
            }
            W.tk_ffirst[k] = first;
        }
        B2C_SYNC();
        B2C_LEADER {
            u32 F = sc->force_break;
            for (int k = 0; k < K; ++k) {
                const u16 fl = P.toks[tk_id[k]].flags;
                const u32 first = W.tk_ffirst[k];
                u8 all = 0;
                u32 one = B2C_NONE_U32;
                if (first != B2C_NONE_U32) {
                    const u32 trail = (fl & B2C_TF_BPE_TRAIL) ? 1u : 0u;
                    if (fl & B2C_TF_BPE_LEAD) { all = 1; F = trail; }
                    else if (F) { one = first; all = static_cast<u8>(trail); F = trail; }
                }
                W.tk_ffirst[k] = one;
                W.tk_fall[k] = all;
            }
            sc->force_break = F;
        }
        B2C_SYNC();
    }

    // ---- phase 1: merge keys + grouping (publish key, fence, claim slot) ------------------
    B2C_FOR(i, M) {
        u32 k, b;
        b2c_divmod(static_cast<u32>(i), n, rcp_n, k, b);
        const bool forced = P.is_bpe && (W.tk_fall[k] || W.tk_ffirst[k] == b);
        B2cExp e;
        b2c_expand(P, cur, b, tk_id[k], forced, t, e);
        const u64 key = b2c_beam_key(e.text_hash, e.part_hash, e.part_len, e.canon);
        C.ckey[i] = key;
        b2c_fence_block();
        b2c_group_insert(C, hmask, static_cast<u32>(i), key);
    }
    B2C_LEADER { sc->max_key = 0; }
    B2C_SYNC();
    B2C_MARK(1);

    // ---- phase 2: fold scores of each group, LM / hotword fusion, running max ------------
    B2C_FOR(i, M) {
        const u32 slot = C.cslot[i];
        if (C.ht_min[slot] != static_cast<u32>(i)) { C.ckey[i] = 0; continue; }
        const u32 last = C.ht_max[slot], cnt = C.ht_cnt[slot];
        // members of a group normally share the token; tokens with identical label strings
        // (string compare in the reference) may merge across tokens, so decode every index
        u32 k0, b0, kl, bl;
        b2c_divmod(static_cast<u32>(i), n, rcp_n, k0, b0);
        b2c_divmod(last, n, rcp_n, kl, bl);
        double s = cur.logit[b0] + tk_lp[k0];
        if (cnt == 2) {
            s = b2c_sum_log_scores(s, cur.logit[bl] + tk_lp[kl]);
        } else if (cnt > 2) {
            for (u32 j = static_cast<u32>(i) + 1; j <= last; ++j) {
                if (C.cslot[j] != slot) continue;
                u32 kj, bj;
                b2c_divmod(j, n, rcp_n, kj, bj);
                s = b2c_sum_log_scores(s, cur.logit[bj] + tk_lp[kj]);
            }
        }
        C.cfold[i] = s;
        C.clast[i] = last;
        const bool forced = P.is_bpe && (W.tk_fall[kl] || W.tk_ffirst[kl] == bl);
        B2cExp e;
        b2c_expand(P, cur, bl, tk_id[kl], forced, t, e);
        double lm_hw = cur.lm_hw[bl];
        if (e.word_len > 0) {
            B2cTextNew tn;
            b2c_text_extend(P, W.text[cur.text_node[bl]], e.word_hash, e.word_len, false, tn);
            lm_hw = tn.lm_hw;
        }
        double ps = 0.0;
        if (e.type == 0) ps = cur.pscore[bl];
        else if (e.part_len > 0) ps = b2c_partial_score(P, e.part_hash, e.part_len);
        const u64 key = b2c_f64_key(b2c_combine_score(P, s, lm_hw, ps, e.part_len));
        C.ckey[i] = key;
        b2c_atomic_max_u64(&sc->max_key, key);
    }
    B2C_SYNC();
    B2C_MARK(2);

    // ---- phase 3: score threshold (decoder.py:545-546) + monotone score buckets ------------
    const double max_score = b2c_key_f64(sc->max_key);
    const double thr = max_score + P.prune_logp;
    const double bscale = P.bucket_scale;
    B2C_FOR(i, M) {
        const u64 key = C.ckey[i];
        if (key == 0) continue;
        const double sco = b2c_key_f64(key);
        if (sco >= thr) {
            const u32 b = b2c_bucket(max_score, sco, bscale);
            b2c_atomic_add_u32(&bcnt[b], 1u);
#if defined(__CUDA_ARCH__)
            C.cnext[i] = atomicExch(&bhead[b], static_cast<u32>(i));
#else
            C.cnext[i] = bhead[b];
            bhead[b] = static_cast<u32>(i);
#endif
        } else {
            C.ckey[i] = 0;
        }
    }
    B2C_SYNC();
    B2C_MARK(3);
    b2c_bucket_scan(bcnt, &sc->n_surv);
    B2C_SYNC();
    B2C_MARK(4);

    // ---- phase 4: stable top-N (decoder.py:548): rank = bucket prefix + exact order inside the
    //      bucket; the history-prune key of every selected candidate goes into the prune table ----
    const u32 m = sc->n_surv;
    const u32 nsel = m < static_cast<u32>(P.beam_width) ? m : static_cast<u32>(P.beam_width);
    B2C_FOR(i, M) {
        const u64 key = C.ckey[i];
        if (key == 0) continue;
        const u32 b = b2c_bucket(max_score, b2c_key_f64(key), bscale);
        u32 rank = bcnt[b];
        for (u32 j = bhead[b]; j != B2C_NONE_U32; j = C.cnext[j]) {
            if (j == static_cast<u32>(i)) continue;
            const u64 kj = C.ckey[j];
            rank += (kj > key || (kj == key && j < static_cast<u32>(i))) ? 1u : 0u;
        }
        if (rank >= nsel) continue;
        ord[rank] = static_cast<u32>(i);
        if (P.prune_history) {
            const u32 last = C.clast[i];
            u32 k, bl;
            b2c_divmod(last, n, rcp_n, k, bl);
            const bool forced = P.is_bpe && (W.tk_fall[k] || W.tk_ffirst[k] == bl);
            B2cExp e;
            b2c_expand(P, cur, bl, tk_id[k], forced, t, e);
            u64 hh = cur.hist_hash[bl];
            if (e.word_len > 0) {
                const B2cText& par = W.text[cur.text_node[bl]];
                const u32 keep = (par.n_win + 1 < static_cast<u32>(P.hist_n)) ? par.n_win : static_cast<u32>(P.hist_n) - 1;
                hh = B2C_HIST_SEED;
                for (int w = static_cast<int>(keep) - 1; w >= 0; --w) hh = b2c_hist_fold(hh, par.win[w]);
                hh = b2c_hist_fold(hh, e.word_hash);
            }
            const u64 hk = b2c_beam_key(hh, e.part_hash, e.part_len, e.canon);
            phk[rank] = hk;
            b2c_fence_block();
            u32 slot = static_cast<u32>(b2c_mix64(hk)) & ptmask;
            while (true) {
                const u32 rep = b2c_atomic_cas_u32(&pt_idx[slot], B2C_NONE_U32, rank);
                if (rep == B2C_NONE_U32) break;
                b2c_fence_block();
                if (phk[rep] == hk) break;
                slot = (slot + 1) & ptmask;
            }
            pslot[rank] = slot;
            b2c_atomic_min_u32(&pt_min[slot], rank);
        }
    }
    B2C_SYNC();
    B2C_MARK(5);

    // ---- phase 5: history prune (decoder.py:550-552): keep the best rank of every key -------
    u32 n_new = nsel;
    if (P.prune_history) {
        b2c_compact_kept(pslot, pt_min, newidx, &sc->n_new, nsel);
        B2C_SYNC();
        B2C_MARK(6);
        n_new = sc->n_new;
    }

    // ---- phase 6: commit the surviving beams; clear the tables for the next frame ------------
    B2C_FOR(j, n_new) {
        const u32 r = P.prune_history ? newidx[j] : static_cast<u32>(j);
        const u32 i = ord[r];
        const u32 last = C.clast[i];
        u32 k, bl;
        b2c_divmod(last, n, rcp_n, k, bl);
        const bool forced = P.is_bpe && (W.tk_fall[k] || W.tk_ffirst[k] == bl);
        B2cExp e;
        b2c_expand(P, cur, bl, tk_id[k], forced, t, e);
        nx.logit[j] = C.cfold[i];
        nx.text_hash[j] = e.text_hash;
        nx.part_hash[j] = e.part_hash;
        nx.part_len[j] = static_cast<u16>(e.part_len);
        nx.last_tok[j] = static_cast<u16>(e.canon);
        nx.pf_s[j] = e.pf_s;
        nx.pf_e[j] = e.pf_e;
        // backtrack chain
        u32 chain = cur.chain[bl];
        if (e.type != 0) {
            const u32 id = b2c_atomic_add_u32(&sc->chain_used, 1u);
            if (id < W.chain_cap) {
                B2cChain c;
                c.parent = chain;
                c.tok = static_cast<u16>(tk_id[k]);
                c.kind = e.type == 3 ? B2C_CK_CONT : (e.type == 2 ? B2C_CK_SPACE : B2C_CK_BPE);
                c.has_word = e.word_len > 0 ? 1 : 0;
                c.ws = cur.pf_s[bl];
                c.we = cur.pf_e[bl];
                W.chain[id] = c;
                chain = id;
            } else {
                b2c_atomic_or_u32(&sc->status, B2C_ERR_CHAIN_FULL);
            }
        }
        nx.chain[j] = chain;
        // text level
        u32 tnode = cur.text_node[bl];
        double lm_hw = cur.lm_hw[bl];
        u64 hh = cur.hist_hash[bl];
        if (e.word_len > 0) {
            const B2cText& par = W.text[tnode];
            B2cTextNew tn;
            b2c_text_extend(P, par, e.word_hash, e.word_len, false, tn);
            lm_hw = tn.lm_hw;
            const u32 id = b2c_atomic_add_u32(&sc->text_used, 1u);
            if (id < W.text_cap) {
                B2cText nt;
                const u32 keep = (par.n_win + 1 < static_cast<u32>(P.hist_n)) ? par.n_win : static_cast<u32>(P.hist_n) - 1;
                nt.win[0] = e.word_hash;
                for (u32 w = 0; w < keep; ++w) nt.win[w + 1] = par.win[w];
                for (u32 w = keep + 1; w < B2C_MAX_HIST; ++w) nt.win[w] = 0;
                nt.n_win = keep + 1;
                hh = B2C_HIST_SEED;
                for (int w = static_cast<int>(nt.n_win) - 1; w >= 0; --w) hh = b2c_hist_fold(hh, nt.win[w]);
                nt.hist_hash = hh;
                nt.raw_lm = tn.raw_lm;
                nt.st = tn.st;
                nt.hw_count = tn.hw_count;
                W.text[id] = nt;
                tnode = id;
            } else {
                b2c_atomic_or_u32(&sc->status, B2C_ERR_TEXT_FULL);
            }
        }
        nx.text_node[j] = tnode;
        nx.lm_hw[j] = lm_hw;
        nx.hist_hash[j] = hh;
        double ps = 0.0;
        if (e.type == 0) ps = cur.pscore[bl];
        else if (e.part_len > 0) ps = b2c_partial_score(P, e.part_hash, e.part_len);
        nx.pscore[j] = ps;
    }
    {
        const u32 M_next = n_new * static_cast<u32>(K_next);
        const B2cCandTier Cn = kFast ? W.tier_s : b2c_pick_tier(W, M_next);
        u32 Hn = b2c_ht_size(M_next);
        if (Hn > Cn.ht_cap) Hn = Cn.ht_cap;
        b2c_clear_tables(W, Cn, Hn);
    }
    B2C_LEADER { sc->n_beams = n_new; }
    B2C_SYNC();
    B2C_MARK(7);
    b2c_swap_tabs(W.cur, W.nxt);
}

// -----------------------------------------------------------------------------------------
// start of an utterance: EMPTY_START_BEAM (decoder.py:130,628) and the root text node
// -----------------------------------------------------------------------------------------
B2C_HD void b2c_utt_begin(const B2cParams& P, B2cWork& W, const B2cLmState* start_state, int K_first) {
    {
        const u32 M0 = static_cast<u32>(K_first > 0 ? K_first : 1);
        const B2cCandTier C0 = b2c_pick_tier(W, M0);
        b2c_clear_tables(W, C0, b2c_ht_size(M0));
    }
    B2C_LEADER {
        B2cScalars* sc = W.sc;
        sc->n_beams = 1;
        sc->chain_used = 0;
        sc->text_used = 1;
        sc->status = B2C_OK;
        sc->force_break = 0;
        B2cText root;
        for (int w = 0; w < B2C_MAX_HIST; ++w) { root.win[w] = 0; root.st.words[w] = 0; root.st.backoff[w] = 0.0f; }
        root.n_win = 0;
        root.hist_hash = B2C_HIST_SEED;
        root.raw_lm = 0.0;
        root.hw_count = 0;
        root.st.length = 0;
        if (P.lm.order > 0) {
            if (start_state) {
                root.st = *start_state;
            } else if (P.score_boundary) {   // BeginSentenceWrite (language_model.py:311-312)
                root.st.length = 1;
                root.st.words[0] = P.lm.bos_id;
                root.st.backoff[0] = P.lm.uni[P.lm.bos_id].backoff;
            }
        }
        W.text[0] = root;
        const B2cBeamTab& c = W.cur;
        c.logit[0] = 0.0;
        c.lm_hw[0] = P.lm.order > 0 ? 0.0 : P.hot_weight * 0;
        c.pscore[0] = 0.0;
        c.text_hash[0] = B2C_TEXT_SEED;
        c.part_hash[0] = 0;
        c.hist_hash[0] = B2C_HIST_SEED;
        c.text_node[0] = 0;
        c.chain[0] = B2C_NONE_U32;
        c.pf_s[0] = -1;
        c.pf_e[0] = -1;
        c.last_tok[0] = B2C_NO_TOK;
        c.part_len[0] = 0;
    }
    B2C_SYNC();
}

// -----------------------------------------------------------------------------------------
// _finalize_beams(force_next_word=True, is_end=True) + output (decoder.py:558-667)
// -----------------------------------------------------------------------------------------
struct B2cOut {              // per-utterance output views (HBM)
    int* n_beams;            // [1]
    int* status;             // [1]
    double* scores;          // [out_beams][2]  logit_score, lm_score
    int* n_tok;              // [out_beams]
    int* n_words;            // [out_beams]
    u32* toks;               // [out_beams][stride]  token | kind << 16, last emission first
    int* frames;             // [out_beams][stride][2] word frames, last word first
    B2cLmState* states;      // [out_beams] LM state after the last word (last_lm_state)
    u32 stride;              // T + 1
};

B2C_HD void b2c_finalize(const B2cParams& P, B2cWork& W, const B2cOut& O) {
    // on entry the grouping table is clear for ht_size(n_beams) slots (last commit / utt_begin)
    B2cScalars* sc = W.sc;
    const u32 n = sc->n_beams;
    const B2cCandTier C = b2c_pick_tier(W, n);
    const u32 hmask = b2c_ht_size(n) - 1;
    const B2cBeamTab cur = W.cur;
    B2C_FOR(b, n) {
        const u64 th = cur.part_len[b] ? b2c_text_append(cur.text_hash[b], cur.part_hash[b]) : cur.text_hash[b];
        const u64 key = b2c_beam_key(th, 0, 0, B2C_NO_TOK);
        C.ckey[b] = key;
        b2c_fence_block();
        b2c_group_insert(C, hmask, static_cast<u32>(b), key);
    }
    B2C_LEADER { sc->max_key = 0; sc->n_surv = 0; }
    B2C_SYNC();
    B2C_FOR(b, n) {
        const u32 slot = C.cslot[b];
        if (C.ht_min[slot] != static_cast<u32>(b)) { C.ckey[b] = 0; continue; }
        const u32 last = C.ht_max[slot];
        double s = cur.logit[b];
        for (u32 j = static_cast<u32>(b) + 1; j <= last; ++j)
            if (C.cslot[j] == slot) s = b2c_sum_log_scores(s, cur.logit[j]);
        C.cfold[b] = s;
        C.clast[b] = last;
        // the LAST duplicate decides the (text, next_word) split that gets scored with is_eos
        // (decoder.py:387-395: an empty next_word is scored as a word -> <unk>)
        double lm_hw;
        if (P.lm.order > 0 || cur.part_len[last] > 0) {
            B2cTextNew tn;
            b2c_text_extend(P, W.text[cur.text_node[last]], cur.part_hash[last], cur.part_len[last], true, tn);
            lm_hw = tn.lm_hw;
        } else {
            lm_hw = cur.lm_hw[last];
        }
        const u64 key = b2c_f64_key(b2c_combine_score(P, s, lm_hw, 0.0, 0));
        C.ckey[b] = key;
        b2c_atomic_max_u64(&sc->max_key, key);
    }
    B2C_SYNC();
    const double thr = b2c_key_f64(sc->max_key) + P.prune_logp;
    B2C_FOR(b, n) {
        const u64 key = C.ckey[b];
        if (key == 0) continue;
        if (b2c_key_f64(key) >= thr) b2c_atomic_add_u32(&sc->n_surv, 1u);
        else C.ckey[b] = 0;
    }
    B2C_SYNC();
    // once per utterance: plain counting rank over the <= beam_width survivors
    const u32 m = sc->n_surv;
    const u32 nsel = m < static_cast<u32>(P.beam_width) ? m : static_cast<u32>(P.beam_width);
    B2C_FOR(b, n) {
        const u64 key = C.ckey[b];
        if (key == 0) continue;
        u32 rank = 0;
        for (u32 j = 0; j < n; ++j) {
            const u64 kj = C.ckey[j];
            rank += (kj > key || (kj == key && kj != 0 && j < static_cast<u32>(b))) ? 1u : 0u;
        }
        if (rank < nsel) W.ord[rank] = static_cast<u32>(b);
    }
    B2C_SYNC();
    const u32 n_out = nsel < static_cast<u32>(P.out_beams) ? nsel : static_cast<u32>(P.out_beams);
    B2C_LEADER { *O.n_beams = static_cast<int>(n_out); *O.status = static_cast<int>(sc->status); }
    // ---- backtrack: one thread per output beam walks its chain ---------------------------
    B2C_FOR(r, n_out) {
        const u32 b = W.ord[r];
        const u32 last = C.clast[b];
        O.scores[2 * r] = C.cfold[b];
        O.scores[2 * r + 1] = b2c_key_f64(C.ckey[b]);
        B2cLmState st;
        st.length = 0;
        for (int w = 0; w < B2C_MAX_HIST; ++w) { st.words[w] = 0; st.backoff[w] = 0.0f; }
        if (P.lm.order > 0) {
            B2cTextNew tn;
            b2c_text_extend(P, W.text[cur.text_node[last]], cur.part_hash[last], cur.part_len[last], true, tn);
            st = tn.st;
        }
        O.states[r] = st;
        u32* toks = O.toks + static_cast<u64>(r) * O.stride;
        int* frames = O.frames + static_cast<u64>(r) * O.stride * 2;
        u32 nt = 0, nw = 0;
        if (cur.part_len[last] > 0) {
            frames[0] = cur.pf_s[last];
            frames[1] = cur.pf_e[last];
            nw = 1;
        }
        u32 node = cur.chain[last];
        while (node != B2C_NONE_U32 && nt < O.stride) {
            const B2cChain c = W.chain[node];
            toks[nt++] = static_cast<u32>(c.tok) | (static_cast<u32>(c.kind) << 16);
            if (c.kind != B2C_CK_CONT && c.has_word && nw < O.stride) {
                frames[2 * nw] = c.ws;
                frames[2 * nw + 1] = c.we;
                ++nw;
            }
            node = c.parent;
        }
        O.n_tok[r] = static_cast<int>(nt);
        O.n_words[r] = static_cast<int>(nw);
    }
    B2C_SYNC();
}
