"""Run tests.synth.special_step_cases through the CUDA path and the oracle; print every mismatching case.
    B200CTC_FORCE_V5=1 B200CTC_V5_VARIANT=0 python tools/special_probe.py [first] [last]
(also usable under `compute-sanitizer --tool racecheck` on a GPU box)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import pyctcdecode_b200 as pkg  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests import synth  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
last = int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9
reps = int(os.environ.get("PROBE_REPS", "1"))
wl = synth.make_workload(dict(kind="char", vocab="B", n_words=400, lm_order=0))
dec = pkg.build_ctcdecoder(wl.labels)
ora = orc.OracleDecoder(wl.labels)
bad = 0
for n, (x, kw) in enumerate(synth.special_step_cases(wl)):
    if n < first or n > last:
        continue
    want = ora.decode_beams(x, **kw)
    for rep in range(reps):
        got = [(b.text, b.text_frames, b.logit_score, b.lm_score) for b in dec.decode_beams(x, **kw)]
        tm = dec.last_timings()
        diff = [j for j, (w, g) in enumerate(zip(want, got)) if w[0] != g[0] or w[1] != g[1] or abs(w[3] - g[3]) > 1e-9 * max(1.0, abs(w[3]))]
        if len(want) != len(got) or diff:
            bad += 1
            j = diff[0] if diff else min(len(want), len(got))
            print("MISMATCH case %d rep %d T=%d %r: %d vs %d beams, first diff at %d: want %r got %r | variant %d cap %d oversize %d inplace %d sorted %d"
                  % (n, rep, x.shape[0], kw, len(want), len(got), j, want[j][0::3] if j < len(want) else None, got[j][0::3] if j < len(got) else None,
                     tm["kernel_variant"], tm["cap_candidates"], tm["oversize_frames"], tm["inplace_frames"], tm["sorted_frames"]), flush=True)
print("done: %d mismatching runs" % bad)
