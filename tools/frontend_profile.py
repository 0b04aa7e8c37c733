#!/usr/bin/env python
"""The feature front end alone at configs[1] geometry (256 x 400 raw frames, synthetic VAD): the target of the ncu
captures under profiles/ (tools/gpu_frontend.sh).  Prints event-timed milliseconds per call."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xvector_b200 import _native, synthetic             # noqa: E402

B, T = 256, 400
eng = _native.XvecEngine([5, 3, 3, 1, 1], [1, 2, 3, 1, 1], [512, 512, 512, 512, 1536], 512, 23, device=0)
lens = np.full(B, T, np.int32)
rng = np.random.default_rng(7)
raw = torch.from_numpy(synthetic.mfcc_batch(7, lens)).cuda()
vad_np = np.concatenate([synthetic.synthetic_vad(rng, T) for _ in range(B)])
vad = torch.from_numpy(vad_np).cuda()
keep = vad_np.reshape(B, T).astype(bool).sum(axis=1).astype(np.int32)
out = torch.empty((int(keep.sum()), 23), device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ms = []
for i in range(iters + 3):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); eng.frontend(raw, vad, lens, keep, out_dev=out); e.record()
    torch.cuda.synchronize()
    if i >= 3:
        ms.append(s.elapsed_time(e))
eng.check_overflow()
alg = B * T * (23 * 4 + 4) + int(keep.sum()) * 23 * 4
print("front end %d x %d raw frames (%.0f %% voiced): %.4f ms/call (min %.4f), %.1f GB/s algorithmic" % (
    B, T, 100.0 * keep.sum() / (B * T), float(np.mean(ms)), float(np.min(ms)), alg / (float(np.mean(ms)) * 1e-3) / 1e9))
