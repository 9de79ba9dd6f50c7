"""CPU tests: the oracle (oracle/ctc_oracle.cpp) against the golden vectors produced by the
unmodified reference, plus the bit-exact host emulations it relies on."""
import math
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import goldens, synth


def _cases():
    g = goldens.load()
    return [c["name"] for c in g["meta"]["cases"]]


@pytest.mark.parametrize("name", _cases())
def test_oracle_matches_reference_golden(golden, name):
    case = next(c for c in golden["meta"]["cases"] if c["name"] == name)
    dec = orc.OracleDecoder(case["labels"], **goldens.lm_kwargs(golden, case))
    x = golden["arrays"][case["array"]]
    got = dec.decode_beams(x, **case["decode"])
    assert goldens.beams_match(case["beams"], got) == ""
    kw = {k: v for k, v in case["decode"].items() if k != "prune_history"}
    assert dec.decode(x, **kw) == case["decode_text"]


def test_exact_scores_pinned_by_reference_test():
    # reference tests/test_decoder.py:328-374
    g = goldens.load()
    case = next(c for c in g["meta"]["cases"] if c["name"] == "lm_unigrams")
    dec = orc.OracleDecoder(case["labels"], **goldens.lm_kwargs(g, case))
    beams = dec.decode_beams(g["arrays"]["test_logits"])
    assert len(beams) == 1
    text, frames, logit, lm = beams[0]
    assert text == "bugs bunny"
    assert frames == [("bugs", (0, 4)), ("bunny", (7, 13))]
    assert abs(logit - (-2.853399551509947)) < 1e-12
    assert abs(lm - 0.14660044849005294) < 1e-12


def test_cpython_set_order_emulation():
    rng = np.random.default_rng(0)
    for V in (8, 29, 32, 64, 300, 1024, 5000):
        for trial in range(300):
            k = int(rng.integers(0, min(V, 1 + trial % 60 if trial % 3 else V)))
            sel = np.sort(rng.choice(V, size=k, replace=False)) if k else np.array([], dtype=np.int64)
            amax = int(rng.integers(0, V)) if (k == 0 or trial % 2) else int(sel[rng.integers(0, k)])
            expected = [int(v) for v in (set(sel) | {np.int64(amax)})]
            assert orc.token_order([int(v) for v in sel], amax) == expected, (V, list(sel), amax)


def test_numpy_pairwise_sum_emulation():
    rng = np.random.default_rng(1)
    L = orc.lib()
    for n in (0, 1, 5, 7, 8, 9, 29, 32, 127, 128, 129, 255, 1000, 1024, 4099):
        a32 = rng.random(n).astype(np.float32)
        a64 = rng.random(n)
        if n:
            assert L.orc_pairwise_sum_f32(a32.ctypes.data, n) == a32.sum()
            assert L.orc_pairwise_sum_f64(a64.ctypes.data, n) == a64.sum()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_probability_detection_matches_numpy(dtype):
    rng = np.random.default_rng(2)
    agree = 0
    for trial in range(200):
        T, V = int(rng.integers(1, 60)), int(rng.choice([5, 29, 32, 100]))
        x = rng.standard_normal((T, V)).astype(dtype)
        if trial % 2:
            e = np.exp(x - x.max(1, keepdims=True))
            x = (e / e.sum(1, keepdims=True)).astype(dtype)
        ref = math.isclose(x.sum(axis=1).mean(), 1)
        assert orc.looks_like_probs(x) == ref
        agree += ref
    assert agree > 0  # the probability branch is exercised


def test_normalise_close_to_numpy_log_softmax():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((50, 32)) * 4).astype(np.float32)
    lp = orc.normalise(x)
    ref = x - x.max(1, keepdims=True)
    ref = ref - np.log(np.exp(ref).sum(1, keepdims=True))
    ref = np.clip(ref, np.log(1e-15), 0)
    assert np.max(np.abs(lp - ref)) < 5e-6
    assert lp.max() <= 0 and lp.min() >= np.log(1e-15)
    # -inf entries (masked logits) clip to log(1e-15)
    x[3, 5] = -np.inf
    assert orc.normalise(x)[3, 5] == np.log(1e-15)


ARPA3 = """\\data\\
ngram 1=6
ngram 2=5
ngram 3=2

\\1-grams:
-2.0\t<unk>
-99\t<s>\t-0.5
-1.0\t</s>
-0.7\ta\t-0.3
-0.9\tb\t-0.4
-1.2\tc

\\2-grams:
-0.2\t<s> a\t-0.25
-0.6\ta b\t-0.35
-0.8\tb a
-0.4\tb </s>
-0.5\ta </s>

\\3-grams:
-0.1\t<s> a b
-0.15\ta b </s>

\\end\\
"""


def test_ngram_backoff_known_answers(tmp_path):
    """Hand-computed KenLM semantics for order 3 with non-zero backoffs (parity unpinned by the
    reference's own tests, whose toy model has only zero backoffs)."""
    p = tmp_path / "t.arpa"
    p.write_text(ARPA3)
    m = orc.OracleNgram(str(p))
    assert m.order == 3
    f = np.float32
    st = m.start_state(bos=True)
    assert st.get() == ([1], [f(-0.5)])
    s, st_a = m.base_score(st, "a")          # "<s> a" found
    assert s == f(-0.2)
    assert st_a.get()[0] == [3, 1]          # a (ext: "a b" exists), "<s> a" (ext: "<s> a b")
    s, st_ab = m.base_score(st_a, "b")       # "<s> a b" found (trigram)
    assert s == f(-0.1)
    assert st_ab.get()[0] == [4, 3]          # "a b" has an extension ("a b </s>"), length 2 kept
    s, _ = m.base_score(st_ab, "</s>")       # "a b </s>"
    assert s == f(-0.15)
    s, st_c = m.base_score(st_a, "c")        # unigram c + bo("a") + bo("<s> a")
    assert s == f(f(f(-1.2) + f(-0.3)) + f(-0.25))
    assert st_c.get()[0] == []               # "c" has no extension
    s, st_u = m.base_score(st_ab, "zzz")     # OOV -> <unk> + bo("b") + bo("a b")
    assert s == f(f(f(-2.0) + f(-0.4)) + f(-0.35))
    assert "zzz" not in m and "<unk>" not in m and "a" in m
    s, st_ba = m.base_score(st_ab, "a")      # "b a" found, then backoff of "a b" (skipped trigram ctx)
    assert s == f(f(-0.8) + f(-0.35))
    assert st_ba.get()[0] == [3]             # "b a" has no extension; "a" has
    null = m.start_state(bos=False)
    s, _ = m.base_score(null, "b")
    assert s == f(-0.9)


@pytest.mark.parametrize("name", goldens.unstable_case_names())
def test_oracle_on_reference_unstable_goldens(name):
    """Cases the unmodified reference decides by rounding noise (integer-valued LibriSpeech logits, wide prune
    settings: 22 distinct reference outcomes under a 1e-13 input perturbation).  The oracle must lie inside that
    family: same beam set, every beam's scores among those the family attaches to it, sorted
    (oracle/gen_golden_unstable.py; VERDICT r1 'the builder's own reference gate is red')."""
    from oracle import oracle as orc

    def run(labels, x, **kw):
        return orc.OracleDecoder(labels).decode_beams(x, **kw)

    assert goldens.run_unstable_case(run, name) == ""
