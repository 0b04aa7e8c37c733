#!/usr/bin/env python
"""A few device-resident extraction steps at BASELINE configs[1] (256 x 400 x 23) for profilers.
    ncu --set full --clock-control none --import-source on --launch-skip 33 -c 11 -o gpurun_out/r02_prof_tdnn \
        python tools/profile_step.py ModelWithoutDropoutTdnn
(3 warm-up steps of 11 launches are skipped; the 4th step is captured.)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from xvector_b200 import _native, synthetic  # noqa: E402

topology = sys.argv[1] if len(sys.argv) > 1 else "ModelWithoutDropoutTdnn"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
topo = bench.TOPOLOGIES[topology]
params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], weight_set="B",
                               activation=topo.get("act", "relu"), pooling=topo.get("pooling", "stats"))
eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], 512, 23, device=0,
                         activation=topo.get("act", "relu"), pooling=topo.get("pooling", "stats"))
eng.set_params(params)
for kv in sys.argv[3:]:
    eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
lens = np.full(256, 400, np.int32)
feats = torch.from_numpy(synthetic.mfcc_batch(2, lens)).cuda()
out = torch.empty((256, 512), dtype=torch.float32, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(steps):
    flush.zero_()
    eng.forward_utts(feats, lens, out)
torch.cuda.synchronize()
eng.check_overflow()
print("profile_step: %s, %d steps, %d launches per step" % (topology, steps, eng.last_launch_count))
