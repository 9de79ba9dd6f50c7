"""bench.py -- throughput of the CTC beam-search hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--regime peaky|diffuse]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # CPU arm: the oracle port on the host cores

A "step" is one decode_batch() pass over one batch of synthetic logits.  `value` is frames/s
with the batch already resident in HBM (device pointers handed to the C ABI); `e2e` is the same
call with PINNED HOST buffers, host->device copies and result copies inside the timed region.
Between timed steps L2 is flushed by writing a 512 MiB buffer (inputs of the headline config are
smaller than L2).  Multi-GPU: one rank per GPU, utterances sharded with no data-path
collective (the only collective is the broadcast of the LM blob before timing) -> "weak".
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "logit frames/sec decoded (batch, beam=100)"

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the headline metric is quoted on
    "c2": dict(kind="char", vocab="B", n_words=20000, lm_order=0, T=1000, batch=256, beam=100, lm={}, hot=0,
               name="Wav2Vec2-base char vocab V=32, T=1000, beam=100, no LM, batch=256 per GPU"),
    "c3": dict(kind="char", vocab="B", n_words=20000, lm_order=3, T=1000, batch=1024, beam=100,
               lm=dict(alpha=0.5, beta=1.0), hot=0,
               name="Wav2Vec2-base V=32, T=1000, beam=100 + synthetic 3-gram (alpha=0.5,beta=1.0), batch=1024 per GPU"),
    "c4": dict(kind="bpe", n_words=50000, lm_order=4, T=500, batch=512, beam=100, lm=dict(alpha=0.5, beta=1.0), hot=16,
               name="Conformer-CTC BPE V=1024, T=500, beam=100 + synthetic 4-gram + 16 hotwords, batch=512 per GPU"),
}


def make_workload(spec):
    from tests import synth
    if spec["kind"] == "char":
        return synth.CharWorkload(spec["vocab"], n_words=spec["n_words"], lm_order=spec["lm_order"])
    return synth.BpeWorkload(n_words=spec["n_words"], lm_order=spec["lm_order"])


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md's clocks line), read through NVML
    in this process every 20 ms.  (A polling `nvidia-smi -lms` child process was used first: its driver queries
    delayed the decode calls of the timed loop by up to 2 ms per step.)  Falls back to nvidia-smi without pynvml."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = None
        self.nvml = None
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    pass
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self.nvml = pynvml
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in self.BITS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def _read(self):
        for line in self.proc.stdout:
            p = [q.strip() for q in line.split(",")]
            if len(p) < 7:
                continue
            try:
                self.sm.append(float(p[0]))
                self.mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            how = "nvml, 20 ms"
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            how = "nvidia-smi -lms 100"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"]}
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": how}


def host_cores():
    """CPU threads this process can really use: os.cpu_count() capped by the scheduler affinity and by the
    cgroup CPU quota (the GPU boxes report 128 logical CPUs under a 16-CPU quota; 128 runnable threads on
    16 CPUs' worth of time only thrash -- measured 56 k frames/s against 83 k with 16-32 threads)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as fh:
                parts = fh.read().split()
            if path.endswith("cpu.max"):
                if parts[0] != "max":
                    n = min(n, max(1, -(-int(parts[0]) // int(parts[1]))))
            else:
                quota = int(parts[0])
                if quota > 0:
                    with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as fh:
                        n = min(n, max(1, -(-quota // int(fh.read().split()[0]))))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


def cpu_arm(wl, spec, kw, xs, beam, hot, target_cpu_seconds=20.0):
    """The reference's CPU path restated (oracle/ctc_oracle.cpp, a C++ port -- the reference itself is
    pure Python) on all host cores, over a bounded sample of the same workload."""
    from oracle import oracle as orc
    cores = host_cores()
    ora = orc.OracleDecoder(wl.labels, **kw)
    t0 = time.perf_counter()
    ora.decode_batch(xs[:2], n_threads=2, beam_width=beam, hotwords=hot)
    per_utt = max((time.perf_counter() - t0), 1e-3)  # two utterances on two threads ~ one utterance-time
    n = int(max(min(len(xs), target_cpu_seconds / per_utt), min(len(xs), cores)))
    sample = xs[:n]
    t0 = time.perf_counter()
    texts = ora.decode_batch(sample, n_threads=cores, beam_width=beam, hotwords=hot)
    dt = time.perf_counter() - t0
    frames = sum(x.shape[0] for x in sample)
    return {"value": frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "%d utterances x T=%d of the same workload, oracle C++ port of the reference's Python loop, "
                      "%d threads, %.1f s" % (n, spec["T"], cores, dt)}, texts, n


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: everything else a library writes to fd 1 (NCCL's version banner,
    nvcc output of a rebuild, ...) is sent to stderr for the lifetime of the process."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(obj) + "\n").encode())


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--regime", default="peaky", choices=["peaky", "diffuse"])
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: the workload's)")
    ap.add_argument("--beam", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--logits-dtype", default="f32", choices=["f32", "f16"],
                    help="dtype of the model output handed to decode_batch (f16: 2-byte elements over PCIe, widened on the device)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = dict(WORKLOADS[args.workload])
    B = args.batch or spec["batch"]
    beam = args.beam or spec["beam"]
    T = spec["T"]

    # ------------------------------------------------------------------------------------
    # reference arm: CPU only, rank 0 only
    # ------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = make_workload(spec)
        kw = dict(spec["lm"])
        if wl.arpa:
            kw.update(kenlm_model_path=wl.arpa, unigrams=wl.words)
        hot = wl.hotwords(spec["hot"]) if spec["hot"] else None
        n_gen = min(B, 8 * host_cores() + 8)
        xs = wl.batch(1, n_gen, T, args.regime)
        vals, info = [], None
        # every step is a bounded sample; the whole run (warm-up + steps) is sized to about two minutes
        per_step = min(8.0, 120.0 / max(1, args.warmup + args.steps))
        for i in range(args.warmup + args.steps):
            info, _, _ = cpu_arm(wl, spec, kw, xs, beam, hot, target_cpu_seconds=per_step)
            if i >= args.warmup:
                vals.append(info["value"])
        v = statistics.mean(vals)
        info["value"] = v
        _emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": spec["name"], "regime": args.regime, "beam_width": beam},
                          "cpu_baseline": info,
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------------------------
    # B200 arm
    # ------------------------------------------------------------------------------------
    import torch

    import __graft_entry__ as graft
    if not os.path.exists(os.path.join(ROOT, "pyctcdecode_b200", "libb200ctc.so")) or rank == 0:
        graft.build()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    else:
        dist = None
        torch.cuda.set_device(local_rank)
    import pyctcdecode_b200 as pkg
    from pyctcdecode_b200 import sharding
    if os.environ.get("B200CTC_PROFILING_LIB"):   # opt-in phase-clock build (profiles/README.md), never the default
        from pyctcdecode_b200 import _lib
        _lib.use_library(os.environ["B200CTC_PROFILING_LIB"])

    wl = make_workload(spec)
    kw = dict(spec["lm"])
    if wl.arpa:
        kw.update(kenlm_model_path=wl.arpa, unigrams=wl.words)
    hot = wl.hotwords(spec["hot"]) if spec["hot"] else None
    if world > 1 and wl.arpa:
        # rank 0 parses the ARPA file; every other rank receives the flattened LM over NCCL
        dec = sharding.build_ctcdecoder_broadcast(wl.labels, device=local_rank, **kw)
    else:
        dec = pkg.build_ctcdecoder(wl.labels, device=local_rank, **kw)

    xs = wl.batch(1 + rank * 100_000, B, T, args.regime)   # every rank its own utterances (weak scaling)
    if args.logits_dtype == "f16":                        # the model emitted half precision: both arms see those values
        xs16 = [x.astype(np.float16) for x in xs]
        xs = [x.astype(np.float32) for x in xs16]
        host = torch.from_numpy(np.stack(xs16)).pin_memory()
    else:
        host = torch.from_numpy(np.stack(xs)).pin_memory()
    dev = host.cuda(non_blocking=False)
    frames_per_step = B * T
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def step_dev():
        return dec.decode_batch(None, dev, beam_width=beam, hotwords=hot)

    def step_host():
        return dec.decode_batch(None, host, beam_width=beam, hotwords=hot)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        total, tms = 0.0, []
        for _ in range(steps):
            flush.fill_(1)                       # L2 flush, outside the timed bracket
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            total += time.perf_counter() - t0
            tms.append(dec.last_timings())
        barrier()
        if dist is not None:
            t = torch.tensor([total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, tms, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total, tms, texts = timed(step_dev, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_total, e2e_tms, texts_e2e = timed(step_host, args.steps, 1)
    assert texts == texts_e2e

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = world * frames_per_step * args.steps / total
    ms_beam = statistics.mean(t["ms_beam"] for t in tms)
    ms_prep = statistics.mean(t["ms_prepare"] for t in tms)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    esz = 2 if args.logits_dtype == "f16" else 4      # algorithmic bytes per logit: the dtype the rows arrive in
    alg_bytes = frames_per_step * wl.V * esz
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            traffic = json.load(fh).get(args.workload, {}).get("beam_kernel_dram_bytes")
    except OSError:
        pass
    ach = alg_bytes / (ms_beam * 1e-3) / 1e9
    ach_prep = alg_bytes / (ms_prep * 1e-3) / 1e9 if ms_prep > 0 else None
    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["name"], "regime": args.regime, "beam_width": beam, "batch_per_gpu": B, "T": T, "V": wl.V,
                   "logits_dtype": args.logits_dtype, "l2": "flushed between timed steps (512 MiB write)", "parallelism": "utterance-sharded x%d" % world},
        "roofline": {"bound": "hbm", "kernel": "b2c_beam_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_beam,
                     "note": "the beam kernel is a T-step serial chain per utterance (latency bound); the streaming stage is reported under roofline_prepare"},
        "roofline_prepare": {"bound": "hbm", "kernel": "b2c_prepare_kernel", "achieved": ach_prep, "peak": peak, "unit": "GB/s",
                             "frac": (ach_prep / peak) if ach_prep else None, "kernel_ms": ms_prep,
                             # what really bounds the streaming stage at V <= 32: instruction issue (float64 exp of every logit,
                             # DESIGN.md section 4).  352 warp instructions per 32-element row is the ncu count of the C2 launch
                             # (profiles/ncu_r01_c2_final_summary.json: 90 M per 256 000 rows); peak = SMs x 4 schedulers x clock
                             "issue": ({"warp_instructions_per_row_ncu": 352, "sm_mhz": clocks.get("sm_mhz"),
                                        "frac_of_issue_peak": (352.0 * frames_per_step / (ms_prep * 1e-3)) /
                                                              (148 * 4 * clocks["sm_mhz"] * 1e6)}
                                       if wl.V <= 32 and ms_prep > 0 and clocks and clocks.get("sm_mhz") else None)},
        "e2e": {"value": world * frames_per_step * args.steps / e2e_total, "unit": "frames/s",
                "h2d_bytes_per_step": int(e2e_tms[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(e2e_tms[-1]["d2h_bytes"]),
                "ms_per_step": 1e3 * e2e_total / args.steps},
        "gpu_launches": int(sum(t["launches"] for t in tms)),
        "device_ms_per_step": statistics.mean(t["ms_total"] for t in tms),
        "beam_kernel_config": {"cap_candidates": tms[-1]["cap_candidates"], "cta_threads": tms[-1]["cta_threads"],
                               "resident_ctas": tms[-1]["cta_slots"], "oversize_frames_per_step": tms[-1]["oversize_frames"],
                               "kernel_variant": tms[-1]["kernel_variant"],
                               "inplace_single_token_frames_per_step": tms[-1]["inplace_frames"],
                               "sorted_no_merge_frames_per_step": tms[-1]["sorted_frames"],
                               "frames_over_128_256_512_1024_2048_4096_total": tms[-1]["cand_hist"]},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        info, cpu_texts, n = cpu_arm(wl, spec, kw, xs, beam, hot)
        out["cpu_baseline"] = info
        out["transcripts_identical_to_oracle"] = "%d/%d" % (sum(a == b for a, b in zip(cpu_texts, texts[:n])), n)
        # "WER parity" of the metric: both systems against the word sequence the synthetic alignment spells
        from tests import synth
        err = {"b200": [0, 0], "cpu_port": [0, 0]}
        for i in range(n):
            truth = wl.truth(1 + i, T)
            for key, hyp in (("b200", texts[i]), ("cpu_port", cpu_texts[i])):
                e, m = synth.word_errors(truth, hyp)
                err[key][0] += e
                err[key][1] += m
        out["wer"] = {k: (v[0] / v[1] if v[1] else None) for k, v in err.items()}
        out["wer"].update(utterances=n, against="ground-truth word sequence of the synthetic alignment (noisy logits, so not 0)")
    _emit(out)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
