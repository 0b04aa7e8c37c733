#!/usr/bin/env python
"""Per-launch device times of one forward over (a) configs[1]'s uniform batch and (b) the configs[2]-like ragged batch of
bench.py (same frame budget), L2 flushed before every forward: where a ragged call spends its extra time."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xvector_oracle as orc                # noqa: E402  (topology table only)
from xvector_b200 import _native, synthetic             # noqa: E402

t = orc.TOPOLOGIES["ModelWithoutDropoutTdnn"]
eng = _native.XvecEngine(t["kernel_sizes"], t["dilations"], t["layer_sizes"], 512, 23, device=0)
eng.set_params(synthetic.make_params(t["kernel_sizes"], t["layer_sizes"], t["embedding_sizes"], weight_set="B"))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
uni = np.full(256, 400, np.int32)
rag = synthetic.lengths_uniform(3, 4096)
rag = rag[:int(np.searchsorted(np.cumsum(rag), int(uni.sum())))].astype(np.int32)
srt = np.sort(rag)
for name, lens in (("uniform 256 x 400", uni), ("ragged 200-1000", rag), ("ragged, sorted by length", srt)):
    feats = torch.from_numpy(synthetic.mfcc_batch(3, lens)).cuda()
    emb = torch.empty((len(lens), 512), device="cuda")
    for _ in range(3):
        eng.forward(feats, lens, emb_dev=emb)
    torch.cuda.synchronize()
    evs = []
    for _ in range(20):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); eng.forward(feats, lens, emb_dev=emb); e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    whole = float(np.mean([s.elapsed_time(e) for s, e in evs]))
    eng.set_option("profile", 1)
    per = []
    for _ in range(10):
        flush.zero_()
        eng.forward(feats, lens, emb_dev=emb)
        per.append(eng.last_kernel_ms())
    eng.set_option("profile", 0)
    per = np.mean(np.asarray(per), axis=0)
    print("%-26s %6d frames %4d segs: %.4f ms/forward (events around the call); per launch (profile mode, no PDL): %s  sum %.4f" % (
        name, int(lens.sum()), len(lens), whole, " ".join("%.4f" % v for v in per), float(per.sum())))
