#!/usr/bin/env python
"""A/B timing of library options inside ONE process (same box, same thermal state), interleaved:
    python tools/ab_bench.py [--topology T] [--rounds R] [--steps K] name=v[,name=v...] name=v[,...] ...
Each positional argument is one configuration (comma-separated xv_set_option pairs; 'base' = defaults).
Prints the median over rounds of the device-resident ms/step and of every kernel's duration."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xvector_b200 import _native, synthetic   # noqa: E402
import bench                                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--topology", default="ModelWithoutDropoutTdnn")
ap.add_argument("--rounds", type=int, default=5)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--frames", type=int, default=400)
ap.add_argument("configs", nargs="+")
args = ap.parse_args()
topo = bench.TOPOLOGIES[args.topology]
params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], weight_set="B",
                               activation=topo.get("act", "relu"))
eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], 512, 23, device=0,
                         activation=topo.get("act", "relu"))
eng.set_params(params)
lens = np.full(args.batch, args.frames, np.int32)
feats = torch.from_numpy(synthetic.mfcc_batch(2, lens)).cuda()
emb = torch.empty((args.batch, 512), dtype=torch.float32, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
DEFAULTS = dict(resident=0, prefetch=0, fc=1, fc_max_splits=36, pdl=1, fuse_tail=1, fuse_first=1)


def apply(cfg):
    opts = dict(DEFAULTS)
    if cfg != "base":
        for kv in cfg.split(","):
            k, v = kv.split("=")
            opts[k] = int(v)
    for k, v in opts.items():
        eng.set_option(k, v)


def measure(steps):
    eng.set_option("profile", 0)
    for _ in range(3):
        eng.forward(feats, lens, emb_dev=emb, stream=stream)
    evs = []
    for _ in range(steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); eng.forward(feats, lens, emb_dev=emb, stream=stream); e.record(stream)
        evs.append((s, e))
    torch.cuda.synchronize()
    step = float(np.median([s.elapsed_time(e) for s, e in evs]))
    eng.set_option("profile", 1)
    ks = []
    for _ in range(max(steps // 3, 5)):
        flush.zero_()
        eng.forward(feats, lens, emb_dev=emb, stream=stream)
        ks.append(eng.last_kernel_ms())
    eng.set_option("profile", 0)
    return step, np.median(np.asarray(ks), axis=0)


feats_host = [torch.from_numpy(synthetic.mfcc_batch(2, lens)).pin_memory(), torch.from_numpy(synthetic.mfcc_batch(3, lens)).pin_memory()]
emb_host = [torch.empty((args.batch, 512)).pin_memory(), torch.empty((args.batch, 512)).pin_memory()]


def measure_e2e(steps):
    """ms/step through xv_submit_host_utts / xv_collect with pinned host buffers, two in flight (bench.py's e2e)."""
    import time
    def run(n):
        prev = None
        t0 = time.perf_counter()
        for i in range(n):
            t = eng.submit_host_utts(feats_host[i & 1], lens, out_host=emb_host[i & 1])
            if prev is not None:
                eng.collect(prev)
            prev = t
        eng.collect(prev)
        return (time.perf_counter() - t0) * 1e3 / n
    run(5)
    return run(steps)


res = {c: [] for c in args.configs}
e2e = {c: [] for c in args.configs}
for r in range(args.rounds):
    for c in args.configs:
        apply(c)
        res[c].append(measure(args.steps))
        e2e[c].append(measure_e2e(max(args.steps, 40)))
for c in args.configs:
    steps = np.array([x[0] for x in res[c]])
    ks = np.median(np.stack([x[1] for x in res[c]]), axis=0)
    print("%-40s step ms median %.4f (min %.4f max %.4f) | kernels us: %s" %
          (c, np.median(steps), steps.min(), steps.max(), " ".join("%.1f" % (k * 1e3) for k in ks)), flush=True)
    print("%-40s e2e ms/step median %.4f (min %.4f max %.4f)" % ("", np.median(e2e[c]), min(e2e[c]), max(e2e[c])), flush=True)
