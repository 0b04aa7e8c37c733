#!/usr/bin/env python
"""Per-tile timeline of one layer of tdnn_pair_kernel (SM clock stamps written by the kernel itself).
usage: python tools/trace_tiles.py [layer] [topology] [cluster]   -> table of cycles relative to the CTA's first stamp."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xvector_b200 import _native, synthetic   # noqa: E402
import bench                                  # noqa: E402

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
topology = sys.argv[2] if len(sys.argv) > 2 else "ModelWithoutDropoutTdnn"
cluster = int(sys.argv[3]) if len(sys.argv) > 3 else 5
topo = bench.TOPOLOGIES[topology]
params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], weight_set="B")
eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], 512, 23, device=0)
eng.set_params(params)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
lens = np.full(256, 400, np.int32)
feats = torch.from_numpy(synthetic.mfcc_batch(2, lens)).cuda()
TT = 16
trace = torch.zeros((74, 2, TT, 8), dtype=torch.int64, device="cuda")
for _ in range(3):
    eng.forward(feats, lens)
torch.cuda.synchronize()
eng.set_option("trace_ptr", trace.data_ptr())
eng.set_option("trace_layer", layer)
eng.forward(feats, lens)
torch.cuda.synchronize()
eng.set_option("trace_layer", -1)
t = trace.cpu().numpy()
names = ["ld_first", "ld_last", "mma_start", "mma_issued", "epi_sees", "epi_release", "epi_done", "epi_done_all"]
for rank in (0, 1):
    x = t[cluster, rank].astype(np.int64)
    base = x[x > 0].min() if (x > 0).any() else 0
    print("layer %d cluster %d rank %d (cycles since first stamp of this CTA)" % (layer, cluster, rank))
    print("tile " + " ".join("%11s" % n for n in names))
    for it in range(TT):
        if not (x[it] > 0).any():
            break
        print("%4d " % it + " ".join("%11d" % (v - base) if v > 0 else "%11s" % "-" for v in x[it, :8]))
x = t[cluster, 0].astype(np.int64)
n = int((x[:, 4] > 0).sum())
if n > 3:
    per = np.diff(x[1:n, 4])
    print("steady-state tile period (epi_sees deltas):", per.tolist())
    print("epilogue busy (epi_done - epi_sees):", (x[1:n, 6] - x[1:n, 4]).tolist())
    print("mma issue span (mma_issued - mma_start):", (x[1:n, 3] - x[1:n, 2]).tolist())
    print("accumulator ready after last MMA issued (epi_sees - mma_issued):", (x[1:n, 4] - x[1:n, 3]).tolist())

# SM clock during the kernel: slot 7 of the fused tail kernel holds %globaltimer (ns) taken together with slot 2 (mma_start)
if (x[:n, 7] > 10**15).all() and n > 8:
    dc = float(x[n - 1, 2] - x[1, 2]); dt = float(x[n - 1, 7] - x[1, 7])
    print("SM clock over jobs 1..%d: %.0f MHz (%d cycles in %.1f us)" % (n - 1, dc / dt * 1e3, dc, dt / 1e3))
