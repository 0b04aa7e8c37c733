#!/usr/bin/env python
"""GPU diagnostic: per-layer error of the sm_100a path vs the fp64 oracle (prints, never asserts).
usage: python tools/diag_gpu.py [topology] [weight_set] [option=value ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xvector_oracle as orc          # noqa: E402
from xvector_b200 import _native, synthetic       # noqa: E402

topology = sys.argv[1] if len(sys.argv) > 1 else "ModelWithoutDropout"
ws = sys.argv[2] if len(sys.argv) > 2 else "B"
t = orc.TOPOLOGIES[topology]
params = synthetic.make_params(t["kernel_sizes"], t["layer_sizes"], t["embedding_sizes"], weight_set=ws)
eng = _native.XvecEngine(t["kernel_sizes"], t["dilations"], t["layer_sizes"], 512, 23, device=0)
eng.set_params(params)
for kv in sys.argv[3:]:
    eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
lens = np.array([200, 37, 131, 25, 300], np.int32)
feats = synthetic.mfcc_batch(11, lens)
print("diag: %s set %s %s" % (topology, ws, " ".join(sys.argv[3:])), flush=True)
emb, layers, stats = eng.forward(torch.from_numpy(feats).cuda(), lens, return_layers=True)
torch.cuda.synchronize()
off = 0
for s, n in enumerate(lens):
    ref_emb, ref_layers, ref_stats = orc.forward(feats[off:off + n], params, topology, return_layers=True)
    errs = []
    for got, want in zip(layers, ref_layers):
        g = got[off:off + n].cpu().numpy().astype(np.float64)
        errs.append(np.abs(g - want).max() / np.abs(want).max())
    gs = stats[s].cpu().numpy().astype(np.float64)
    es = np.abs(gs - ref_stats).max() / np.abs(ref_stats).max()
    m = orc.parity_metrics(emb[s].cpu().numpy(), ref_emb)
    print("  seg %d len %4d layer errs %s stats %.2e emb max_rel %.2e" %
          (s, n, " ".join("%.2e" % e for e in errs), es, m["max_rel"]), flush=True)
    off += n
try:
    eng.check_overflow()
    print("  overflow: none")
except Exception as e:
    print("  overflow:", e)
