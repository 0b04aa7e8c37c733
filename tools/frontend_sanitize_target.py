#!/usr/bin/env python
"""Target of tools/gpu_sanitize_frontend.sh: the front-end kernels alone over a mixed batch (empty, 1-frame, whole-utterance
tiles, split utterances, one utterance long enough for the tile-count pass; with and without VAD; mean and variance)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xvector_b200 import _native                        # noqa: E402

eng = _native.XvecEngine([5, 3, 3, 1, 1], [1, 2, 3, 1, 1], [512, 512, 512, 512, 1536], 512, 23, device=0)
rng = np.random.default_rng(0)
lens = np.array([0, 1, 31, 400, 512, 513, 777, 1300, 4500], np.int32)
x = torch.from_numpy(rng.standard_normal((int(lens.sum()), 23)).astype(np.float32) * 10).cuda()
vad_np = (rng.random(int(lens.sum())) < 0.7).astype(np.float32)
keep = np.array([int(v.sum()) for v in np.split(vad_np, np.cumsum(lens)[:-1])], np.int32)
vad = torch.from_numpy(vad_np).cuda()
for opts in (_native.XvCmvnOpts(), _native.XvCmvnOpts(300, 100, True, True), _native.XvCmvnOpts(64, 32, False, False),
             _native.XvCmvnOpts(1000, 100, True, False)):
    a = eng.frontend(x, vad, lens, keep, opts)
    b = eng.frontend(x, None, lens, None, opts)
    torch.cuda.synchronize()
    eng.check_overflow()
    assert a.shape[0] == int(keep.sum()) and b.shape[0] == int(lens.sum())
    assert bool(torch.isfinite(a).all()) and bool(torch.isfinite(b).all())
print("front end sanitizer target: ok")
