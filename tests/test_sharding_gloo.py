"""world_size-2 `gloo` test of the multi-GPU plumbing on CPU: the LM blob broadcast gives
bit-identical tables on every rank, utterance sharding is a partition, and the sharded
decode equals the single-process decode.  The kernels themselves are the tests/hostsim
simulation build here (no GPU in this container); the plumbing under test is the product's."""
import hashlib
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import ctypes as C, hashlib, json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from pyctcdecode_b200 import _lib, sharding
import pyctcdecode_b200 as pkg
from tests import synth
_lib.use_library(os.path.join(%(root)r, "tests", "hostsim", "libb200ctc_hostsim.so"))
os.environ["B200CTC_DEVICE"] = "0"   # the simulation build exposes a single fake device
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
wl = synth.CharWorkload("B", n_words=300, lm_order=3)
kw = dict(kenlm_model_path=wl.arpa if rank == 0 else wl.arpa + ".does-not-exist-on-this-rank", unigrams=wl.words, alpha=0.5, beta=1.0)
dec = sharding.build_ctcdecoder_broadcast(wl.labels, device=None, **kw)
addr, size = dec._language_model.ngram_model.blob()
digest = hashlib.sha1(bytes((C.c_uint8 * size).from_address(addr))).hexdigest()
Ts = [50, 0, 120, 7, 33, 90, 64, 1, 15]
xs = [wl.utterance(300 + i, T, "peaky") if T else np.zeros((0, wl.V), np.float32) for i, T in enumerate(Ts)]
texts = sharding.decode_batch_sharded(dec, xs, beam_width=20)
shards = sharding.shard_utterances(Ts, world)
out = [None] * world
dist.all_gather_object(out, (digest, texts, shards))
if rank == 0:
    single = pkg.build_ctcdecoder(wl.labels, kenlm_model_path=wl.arpa, unigrams=wl.words, alpha=0.5, beta=1.0)
    print(json.dumps({"digests": [o[0] for o in out], "texts": [o[1] for o in out], "shards": out[0][2],
                      "single": single.decode_batch(None, xs, beam_width=20)}))
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_utterances_is_a_balanced_partition():
    from pyctcdecode_b200 import sharding
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        lengths = rng.integers(0, 2000, size=97).tolist()
        shards = sharding.shard_utterances(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(97))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)


def test_two_rank_gloo_broadcast_and_sharded_decode(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "hostsim")])
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    import json
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert len(set(out["digests"])) == 1          # every rank holds the same LM bytes
    assert out["texts"][0] == out["texts"][1] == out["single"]
    assert sorted(i for s in out["shards"] for i in s) == list(range(9))
