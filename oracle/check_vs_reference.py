"""TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).

Differential check of the C++ oracle (oracle/ctc_oracle.cpp) against the UNMODIFIED reference
imported from /root/reference with the stand-ins in oracle/refshim.  Run:
    python oracle/check_vs_reference.py [n_random]
Prints one line per case family and exits non-zero on any transcript / frame mismatch that is
not a documented near-tie.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import logging  # noqa: E402

logging.disable(logging.CRITICAL)

from pyctcdecode import build_ctcdecoder  # noqa: E402  (the reference)

from oracle import oracle as orc  # noqa: E402
from tests import synth  # noqa: E402

SCORE_TOL = 2e-4  # numpy evaluates log-softmax in float32 with its own SIMD exp/log; see DESIGN.md


def compare(ref_beams, orc_beams, tag, stats):
    ok = True
    if len(ref_beams) != len(orc_beams):
        ok = False
    else:
        for rb, ob in zip(ref_beams, orc_beams):
            if rb.text != ob[0] or [tuple(f) for f in rb.text_frames] != [(w, tuple(fr)) for w, fr in ob[1]]:
                ok = False
                break
            if abs(rb.logit_score - ob[2]) > SCORE_TOL + 1e-6 * abs(rb.logit_score) or \
               abs(rb.lm_score - ob[3]) > SCORE_TOL + 1e-6 * abs(rb.lm_score):
                ok = False
                break
    stats["n"] += 1
    if not ok:
        # near tie?  compare as sets of (text) with score tolerance
        rset = {b.text: b.lm_score for b in ref_beams}
        oset = {b[0]: b[3] for b in orc_beams}
        common = set(rset) & set(oset)
        worst = max([abs(rset[k] - oset[k]) for k in common], default=0.0)
        top_same = bool(ref_beams) and bool(orc_beams) and ref_beams[0].text == orc_beams[0][0]
        stats["mismatch"].append((tag, len(ref_beams), len(orc_beams), top_same, worst))
    return ok


def main():
    n_random = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    stats = {"n": 0, "mismatch": []}
    ref_tests = "/root/reference/pyctcdecode/tests"
    arpa = os.path.join(ref_tests, "sample_data", "bugs_bunny_kenlm.arpa")

    # ---- the reference's own fixtures -------------------------------------------------
    sys.path.insert(0, os.path.dirname(ref_tests))
    SAMPLE_LABELS = [" ", "b", "g", "n", "s", "u", "y", ""]
    vocab = {c: n for n, c in enumerate(SAMPLE_LABELS)}

    def onehot(chars):
        m = np.zeros((len(chars), len(SAMPLE_LABELS)))
        for i, c in enumerate(chars):
            m[i][vocab[c]] = 1
        return m

    bugs, bunny = onehot("bugs"), onehot(["b", "u", "n", "", "n", "y"])
    blank, space = onehot([""]), onehot([" "])
    test_probs = np.vstack([np.vstack([bugs, blank, blank]) * 0.49 + bunny * 0.51, space, bunny])
    test_logits = np.log(np.clip(test_probs, 1e-15, 1))
    with open(os.path.join(ref_tests, "sample_data", "libri_logits.json")) as fh:
        libri = np.array(json.load(fh))

    fixtures = [("test_logits", SAMPLE_LABELS, test_logits), ("test_probs", SAMPLE_LABELS, test_probs),
                ("libri", synth.LIBRI_LABELS, libri),
                ("libri_f32", synth.LIBRI_LABELS, libri.astype(np.float32))]
    lm_variants = [dict(), dict(kenlm_model_path=arpa), dict(kenlm_model_path=arpa, unigrams=["bugs", "bunny"]),
                   dict(kenlm_model_path=arpa, unigrams=["bunny"], alpha=1.0),
                   dict(kenlm_model_path=arpa, alpha=1.0, unk_score_offset=0.0, lm_score_boundary=False)]
    dec_variants = [dict(), dict(prune_history=True), dict(hotwords=["bugs"], hotword_weight=20.0),
                    dict(beam_prune_logp=-20.0, token_min_logp=-8.0), dict(beam_width=3),
                    dict(hotwords=["bunny bugs", "i have"], hotword_weight=5.0, prune_history=True)]
    for name, labels, logits in fixtures:
        for li, lmkw in enumerate(lm_variants):
            if "libri" in name and lmkw:
                continue
            ref = build_ctcdecoder(labels, **lmkw)
            mine = orc.OracleDecoder(labels, **lmkw)
            for di, dkw in enumerate(dec_variants):
                compare(ref.decode_beams(logits, **dkw), mine.decode_beams(logits, **dkw), "%s/lm%d/d%d" % (name, li, di), stats)
    print("fixtures: %d cases, %d mismatches" % (stats["n"], len(stats["mismatch"])))

    # ---- random synthetic -----------------------------------------------------------
    wl_nolm = synth.CharWorkload("B", n_words=2000, lm_order=0)
    wl_lm = synth.CharWorkload("B", n_words=2000, lm_order=3)
    wl_a = synth.CharWorkload("A", n_words=2000, lm_order=2)
    bpe = synth.BpeWorkload(n_words=2000, lm_order=3)
    rng = np.random.default_rng(123)
    fams = [("B/nolm", wl_nolm, {}), ("B/3gram", wl_lm, dict(kenlm_model_path=wl_lm.arpa, unigrams=wl_lm.words, alpha=0.5, beta=1.0)),
            ("A/2gram", wl_a, dict(kenlm_model_path=wl_a.arpa, unigrams=wl_a.words)),
            ("C/bpe-nolm", bpe, {}), ("C/bpe-3gram", bpe, dict(kenlm_model_path=bpe.arpa, unigrams=bpe.words))]
    for fam, wl, lmkw in fams:
        ref = build_ctcdecoder(wl.labels, **lmkw)
        mine = orc.OracleDecoder(wl.labels, **lmkw)
        n0 = stats["n"]
        m0 = len(stats["mismatch"])
        for i in range(n_random):
            T = int(rng.integers(0, 160))
            regime = ["peaky", "diffuse", "flat"][i % 3] if wl.V <= 64 else ["peaky", "diffuse"][i % 2]
            x = wl.utterance(1000 + i, T, regime) if T > 0 else np.zeros((0, wl.V), np.float32)
            if i % 5 == 4:
                x = x.astype(np.float64)
            if i % 7 == 6 and T > 0:
                e = np.exp(x - x.max(1, keepdims=True))
                x = (e / e.sum(1, keepdims=True)).astype(x.dtype)
            dkw = dict(beam_width=[100, 8, 25][i % 3], prune_history=bool(i % 2))
            if i % 4 == 3:
                dkw.update(hotwords=[wl.words[3], wl.words[10] + " " + wl.words[11]], hotword_weight=6.0)
            compare(ref.decode_beams(x, **dkw), mine.decode_beams(x, **dkw), "%s/%d" % (fam, i), stats)
        print("%s: %d cases, %d mismatches" % (fam, stats["n"] - n0, len(stats["mismatch"]) - m0))

    hard = [m for m in stats["mismatch"] if not m[3] or m[4] > 1e-3]
    for m in stats["mismatch"]:
        print("MISMATCH", m)
    print("total %d cases, %d mismatches (%d hard)" % (stats["n"], len(stats["mismatch"]), len(hard)))
    return 1 if hard else 0


if __name__ == "__main__":
    sys.exit(main())
