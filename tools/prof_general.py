"""Profiling helper (GPU box): phase clocks of the wide-beam / diffuse regimes and a host profile of decode_beams_batch.
Run with the phase-clock build in place of the product library for the first part:
    python tools/prof_general.py clocks   (stderr carries one phase-clock line per call)
    python tools/prof_general.py one beam500   (two calls of one configuration, e.g. under ncu -k regex:b2c_beam_kernel -s 1 -c 1)
    python tools/prof_general.py beams    (cProfile of decode_beams_batch at the C2 shape)
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import pyctcdecode_b200 as pkg


def main():
    what = sys.argv[1]
    spec = bench.WORKLOADS["c2"]
    wl, kw, hot = bench.workload_objects(spec)
    dec = pkg.build_ctcdecoder(wl.labels, device=0, **kw)
    cfgs = (("beam500", "peaky", 500, 256), ("beam2000", "peaky", 2000, 256), ("diffuse", "diffuse", 100, 256), ("beam200", "peaky", 200, 256))
    if what in ("clocks", "one"):
        for name, regime, beam, B in [c for c in cfgs if what == "clocks" or c[0] == sys.argv[2]]:
            xs = wl.batch(1, B, spec["T"], regime)
            dev = torch.from_numpy(np.stack(xs)).cuda()
            for it in range(2):
                sys.stderr.write("== %s call %d\n" % (name, it))
                sys.stderr.flush()
                t0 = time.perf_counter()
                dec.decode_batch(None, dev, beam_width=beam)
                torch.cuda.synchronize()
                tm = dec.last_timings()
                sys.stderr.write("   wall %.2f ms  %s\n" % (1e3 * (time.perf_counter() - t0), {k: tm[k] for k in sorted(tm)}))
    else:
        xs = wl.batch(1, 256, spec["T"], "peaky")
        dev = torch.from_numpy(np.stack(xs)).cuda()
        for _ in range(2):
            dec.decode_beams_batch(None, dev, beam_width=100)
        pass
        pr = cProfile.Profile()
        pr.enable()
        t0 = time.perf_counter()
        out = dec.decode_beams_batch(None, dev, beam_width=100)
        t1 = time.perf_counter()
        pr.disable()
        print("decode_beams_batch wall %.1f ms; beams %d; words in beam 0: %d; timings %s" % (
            1e3 * (t1 - t0), sum(len(b) for b in out), len(out[0][0].text_frames), dec.last_timings()))
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


main()
