"""Loader for tests/golden (vectors produced by oracle/gen_golden.py from the unmodified reference)."""
import json
import os
import tempfile

import numpy as np

from tests import synth

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load():
    if "g" not in _cache:
        with open(os.path.join(HERE, "golden", "cases.json"), encoding="utf-8") as fh:
            meta = json.load(fh)
        arrays = dict(np.load(os.path.join(HERE, "golden", "arrays.npz")))
        d = os.path.join(tempfile.gettempdir(), "b200ctc_synth")
        os.makedirs(d, exist_ok=True)
        toy = os.path.join(d, "bugs_bunny_kenlm.arpa")
        with open(toy, "w", encoding="utf-8") as fh:
            fh.write(meta["toy_arpa"])
        _cache["g"] = {"meta": meta, "arrays": arrays, "toy_arpa": toy, "workloads": {}}
    return _cache["g"]


def lm_kwargs(g, case):
    """build_ctcdecoder keyword arguments (kenlm_model_path, unigrams, alpha, ...) for a golden case."""
    lm = case["lm"]
    if lm is None:
        return {}
    kw = {k: v for k, v in lm.items() if k not in ("arpa_kind", "workload")}
    if lm["arpa_kind"] == "toy":
        kw["kenlm_model_path"] = g["toy_arpa"]
    else:
        key = json.dumps(lm["workload"], sort_keys=True)
        if key not in g["workloads"]:
            g["workloads"][key] = synth.make_workload(lm["workload"])
        wl = g["workloads"][key]
        kw["kenlm_model_path"] = wl.arpa
        kw["unigrams"] = wl.words
    return kw


def beams_match(expected, got, tol=2e-4, exact_order=True):
    """got: list of (text, [(word,(s,e))...], logit, lm). Returns '' or a description of the first difference."""
    if len(expected) != len(got):
        return "beam count %d != %d" % (len(got), len(expected))
    for i, (e, g) in enumerate(zip(expected, got)):
        if e["text"] != g[0]:
            return "beam %d text %r != %r" % (i, g[0], e["text"])
        ef = [(w, (s, t)) for w, s, t in e["frames"]]
        gf = [(w, (int(f[0]), int(f[1]))) for w, f in g[1]]
        if ef != gf:
            return "beam %d frames %r != %r" % (i, gf, ef)
        if abs(e["logit_score"] - g[2]) > tol + 1e-6 * abs(e["logit_score"]):
            return "beam %d logit %r != %r" % (i, g[2], e["logit_score"])
        if abs(e["lm_score"] - g[3]) > tol + 1e-6 * abs(e["lm_score"]):
            return "beam %d lm %r != %r" % (i, g[3], e["lm_score"])
    return ""


def build_product_decoder(pkg, labels, **kw):
    """build_ctcdecoder, except for unigrams == [] where the reference itself divides by zero in
    verify_alphabet_coverage and its test assembles the pieces by hand (tests/test_decoder.py:266)."""
    if kw.get("unigrams") == []:
        lm = pkg.LanguageModel(pkg.NgramModel(kw["kenlm_model_path"]), [], alpha=kw.get("alpha", 0.5),
                               beta=kw.get("beta", 1.5), unk_score_offset=kw.get("unk_score_offset", -10.0),
                               score_boundary=kw.get("lm_score_boundary", True))
        return pkg.BeamSearchDecoderCTC(pkg.Alphabet.build_alphabet(labels), lm)
    return pkg.build_ctcdecoder(labels, **kw)
