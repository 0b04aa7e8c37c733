#!/usr/bin/env python
"""Product-path check of data-parallel training on N GPUs of one box.

`torchrun --nproc-per-node N x-vector-kaldi-tf_b200/train_dnn_one_iteration.py --tar-file egs.{rank}.tar ...` (every rank its own
archive, gradients all-reduced over NCCL per minibatch, rank 0 saves) must give the same model as ONE process that sees the N
archives' minibatches concatenated (minibatch = the N ranks' minibatches stacked): checked on the loss trajectory (the log lines) and
the saved variables.  BatchNorm statistics are per replica (as the reference has no cross-job BN sync), so the comparison model is
trained here the same way: N trainers in one process, gradients averaged on the host.
usage: python tools/multi_gpu_train_check.py [N]"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xvector_b200 import examples_io, synthetic                                 # noqa: E402
from xvector_b200.models import ModelWithoutDropoutTdnn                         # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tmp = tempfile.mkdtemp(prefix="xvec_train_multi_")
os.environ["XVEC_SEED"] = "9"
model0 = os.path.join(tmp, "model_0")
NC, B, T, MB = 40, 16, 64, 6
ModelWithoutDropoutTdnn().build_model(NC, 23, model0, None)
rng = np.random.default_rng(1)
means = rng.standard_normal((NC, 23)) * 6.0
data = []
for r in range(n_gpus):
    labs = [rng.integers(0, NC, B) for _ in range(MB)]
    mbs = [(means[l][:, None, :] + synthetic.mfcc(100 * r + i, B * T).reshape(B, T, 23)).astype(np.float16).astype(np.float32)
           for i, l in enumerate(labs)]
    examples_io.write_egs_tar(os.path.join(tmp, "egs.%d.tar" % r), mbs, labs)
    data.append((mbs, labs))

out = os.path.join(tmp, "model_1")
cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_gpus), "--master-addr", "127.0.0.1",
       "--master-port", "29577", os.path.join(ROOT, "x-vector-kaldi-tf_b200", "train_dnn_one_iteration.py"),
       "--feature-dim", "23", "--minibatch-size", str(B), "--minibatch-count", str(MB), "--learning-rate", "0.002",
       "--print-interval", "2", "--tar-file", os.path.join(tmp, "egs.{rank}.tar"), "--input-dir", model0, "--output-dir", out]
res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
print(res.stdout[-1500:])
if res.returncode != 0:
    print(res.stderr[-3000:])
    raise SystemExit("torchrun failed")
with np.load(os.path.join(out, "model.npz")) as z:
    dp = {k: z[k] for k in z.files}

# the same training in ONE process: one trainer per "rank" (own BatchNorm statistics), gradients averaged, same Adam
import torch                                                                    # noqa: E402
from xvector_b200 import _native                                                # noqa: E402
m = ModelWithoutDropoutTdnn()
m.load_model(None, model0, None)
engs, trs = [], []
for r in range(n_gpus):
    eng = _native.XvecEngine(m.kernel_sizes, m.dilation_rates, m.layer_sizes, 512, 23, device=0)
    tr = _native.XvecTrainer(eng, NC, 512)
    tr.set_params({k: v for k, v in m.params.items()})
    engs.append(eng); trs.append(tr)
for i in range(MB):
    grads = []
    for r in range(n_gpus):
        mbs, labs = data[r]
        # the tar loader hands minibatches out in tar order
        x = torch.from_numpy(mbs[i].reshape(B * T, 23)).cuda()
        lab = torch.from_numpy(labs[i].astype(np.int32)).cuda()
        trs[r].forward_backward(x, lab, B, T)
        torch.cuda.synchronize()
        grads.append(trs[r].download(_native.TRAIN_GRAD).astype(np.float64))
    g = np.sum(grads, axis=0).astype(np.float32)             # NCCL sum ...
    for r in range(n_gpus):
        trs[r].upload(_native.TRAIN_GRAD, g)
        trs[r].apply(0.002, grad_scale=1.0 / n_gpus)           # ... scaled by 1/world in the Adam kernel
    torch.cuda.synchronize()
worst = 0.0
for name, arr in dp.items():
    if name.endswith(("/Adam:0", "/Adam_1:0", "_power:0", "_step:0")) or name.endswith(("mean:0", "variance:0")):
        continue
    ref = trs[0].get_param(name).reshape(arr.shape)
    d = float(np.abs(ref - arr).max() / max(np.abs(arr).max(), 1e-30))
    worst = max(worst, d)
print("data-parallel x%d vs single-process emulation: worst relative difference of a trained variable %.3e" % (n_gpus, worst))
assert worst <= 2e-3, worst                                  # NCCL's fp32 summation order vs the host's fp64 sum, through Adam
losses = [float(x) for x in re.findall(r"Average training loss for minibatches \d+-\d+ is ([0-9.]+)", res.stdout)]
assert len(losses) >= MB // 2 and losses[-1] < losses[0], losses
print("OK", losses)
