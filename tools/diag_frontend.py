#!/usr/bin/env python
"""Where the device front end and the oracle differ (GPU box diagnostic): mismatch count, size in float ulps of the
output, and the (utterance, frame, bin) of the first few."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kaldi_frontend_oracle as fe          # noqa: E402
from oracle import xvector_oracle as orc                # noqa: E402
from xvector_b200 import _native                        # noqa: E402

t = orc.TOPOLOGIES["ModelWithoutDropoutTdnn"]
eng = _native.XvecEngine(t["kernel_sizes"], t["dilations"], t["layer_sizes"], 512, 23, device=0)
rng = np.random.default_rng(7)
lens = [5, 127, 300, 301, 1000, 16, 2500]
scale = 12.0 / np.sqrt(1.0 + np.arange(23))
feats = [(rng.standard_normal((n, 23)) * scale - 35.0 * (np.arange(23) == 0)).astype(np.float32) for n in lens]
for window in (7, 300):
    opts = _native.XvCmvnOpts(window, min(100, window), True, False)
    got = eng.frontend(torch.from_numpy(np.concatenate(feats)).cuda(), None, np.array(lens, np.int32), None, opts).cpu().numpy()
    want = np.concatenate([fe.sliding_window_cmn(f, window) for f in feats])
    alt = np.concatenate([fe.sliding_window_cmn_direct(f, window) for f in feats])
    x = np.concatenate(feats)
    bad = np.argwhere(got != want)
    print("window %d: %d of %d differ from the running-sum oracle; %d from the prefix-sum oracle; oracles differ in %d" % (
        window, len(bad), got.size, int((got != alt).sum()), int((want != alt).sum())))
    starts = np.concatenate([[0], np.cumsum(lens)])
    for r, d in bad[:12]:
        u = int(np.searchsorted(starts, r, side="right") - 1)
        ulps = (float(got[r, d]) - float(want[r, d])) / float(np.spacing(np.abs(want[r, d])))
        print("  utt %d (T=%d) t=%d bin %d: x=%r got=%r want=%r  (%.1f ulp of the output)" % (
            u, lens[u], r - starts[u], d, float(x[r, d]), float(got[r, d]), float(want[r, d]), ulps))
