#!/usr/bin/env python
"""Per-kernel timing of one training step at BASELINE config 5 size (64 x 400 frames, 5000 speakers) on one GPU.
usage: bench_train.py [topology] [B] [T] [classes] [steps]"""
import json
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xvector_b200 import _native, synthetic                 # noqa: E402
from xvector_b200 import models                             # noqa: E402


def main():
    topology = sys.argv[1] if len(sys.argv) > 1 else "ModelWithoutDropoutTdnn"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    NC = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 50
    cls = getattr(models, topology)
    topo = dict(kernel_sizes=cls.kernel_sizes, dilations=cls.dilation_rates, layer_sizes=cls.layer_sizes, embedding_sizes=cls.embedding_sizes)
    P = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], num_classes=NC, weight_set="B")
    eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], 512, 23, device=0)
    tr = _native.XvecTrainer(eng, NC, 512)
    tr.set_params(P)
    for kv in os.environ.get("XVEC_TRAIN_OPTS", "").split(","):
        if "=" in kv:
            tr.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    feats = torch.from_numpy(synthetic.mfcc(5, B * T)).cuda()
    lab = torch.from_numpy(np.random.default_rng(5).integers(0, NC, B).astype(np.int32)).cuda()
    for _ in range(5):
        la = tr.forward_backward(feats, lab, B, T)
        tr.apply(1e-4)
    torch.cuda.synchronize()
    eng.check_overflow()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    e0.record()
    h0 = time.perf_counter()
    for _ in range(steps):
        la = tr.forward_backward(feats, lab, B, T)
        tr.apply(1e-4)
    host_ms = (time.perf_counter() - h0) * 1e3 / steps        # time the host needs to ENQUEUE one step
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("host enqueue time %.3f ms/step" % host_ms)
    print("train step %s B=%d T=%d classes=%d: %.3f ms/step, %.1f M frames/s, loss %.4f, launches/step %d" %
          (topology, B, T, NC, ms, B * T / ms / 1e3, float(la[0]), tr.last_launch_count))
    eng.set_option("profile", 1)
    agg = OrderedDict()
    n_rep = 5
    for _ in range(n_rep):
        tr.forward_backward(feats, lab, B, T)
        tr.apply(1e-4)
        torch.cuda.synchronize()
        times, names = eng.last_kernel_ms(), tr.last_kernel_names()
        assert len(times) == len(names), (len(times), len(names))
        for n, t_ in zip(names, times):
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1; a[1] += t_
    tot = sum(v[1] for v in agg.values()) / n_rep
    print("  sum of per-launch device times %.3f ms" % tot)
    for n, (c, t_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-28s x%-3d %8.3f ms  %5.1f %%" % (n, c // n_rep, t_ / n_rep, 100 * t_ / n_rep / tot))
    eng.set_option("profile", 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"ms_per_step": ms, "kernels": {n: [c // n_rep, t_ / n_rep] for n, (c, t_) in agg.items()}},
              open("gpurun_out/bench_train_%s.json" % topology, "w"))
    tr.close(); eng.close()


if __name__ == "__main__":
    main()
