#!/usr/bin/env python
"""Where the wall time of the pipelined host API goes (two submissions in flight): time inside the submit call (host work:
metadata tables, tensor maps, launches) against time waiting in xv_collect (the GPU).
usage: python tools/host_cost.py [raw]      raw: xv_submit_host_raw (front end + network) instead of xv_submit_host_utts"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xvector_b200 import _native, synthetic   # noqa: E402
import bench                                  # noqa: E402

raw_mode = len(sys.argv) > 1 and sys.argv[1] == "raw"
topo = bench.TOPOLOGIES["ModelWithoutDropoutTdnn"]
params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], weight_set="B")
eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], 512, 23, device=0)
eng.set_params(params)
B, T = 256, 400
lens = np.full(B, T, np.int32)
fh = [torch.from_numpy(synthetic.mfcc_batch(2 + i, lens)).pin_memory() for i in range(2)]
eh = [torch.empty((B, 512)).pin_memory() for _ in range(2)]
if raw_mode:
    rng = np.random.default_rng(3)
    tracks = [synthetic.synthetic_vad(rng, T) for _ in range(B)]
    vad = torch.from_numpy(np.concatenate(tracks)).pin_memory()
    keep = np.array([int(np.count_nonzero(t)) for t in tracks], np.int32)


def submit(i):
    if raw_mode:
        return eng.submit_host_raw(fh[i & 1], vad, lens, keep, keep, eh[i & 1], None)
    return eng.submit_host_utts(fh[i & 1], lens, out_host=eh[i & 1])


def run(n):
    ts, tc = 0.0, 0.0
    prev = None
    t0 = time.perf_counter()
    for i in range(n):
        a = time.perf_counter()
        t = submit(i)
        b = time.perf_counter()
        if prev is not None:
            eng.collect(prev)
        c = time.perf_counter()
        ts += b - a
        tc += c - b
        prev = t
    eng.collect(prev)
    tot = time.perf_counter() - t0
    return tot / n * 1e3, ts / n * 1e3, tc / n * 1e3


run(10)
for _ in range(4):
    print("%s: e2e ms/step %.4f  inside submit %.4f  waiting in collect %.4f" % ((("raw" if raw_mode else "utts"),) + run(100)))
