#!/usr/bin/env python
"""Data-parallel training step (configs[4]: 64 x 400 frames per GPU, 5000 speakers) under torchrun: ms/step (device events, max
over ranks) for combinations of  clusters=<CTA pairs per persistent grid>  and  buckets=fine|two|none  (none: no all-reduce at
all = the compute-only step).  NCCL_MAX_NCHANNELS etc. come from the environment.
    torchrun --nproc-per-node N tools/bench_train_dp.py clusters=74,buckets=fine clusters=70,buckets=fine ..."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xvector_b200 import _native, synthetic   # noqa: E402
import bench                                  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
topo = bench.TOPOLOGIES["ModelWithoutDropoutTdnn"]
B, T, NC = 64, 400, 5000
params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], num_classes=NC, weight_set="B")
eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], 512, 23, device=local)
eng.set_params(params)
tr = _native.XvecTrainer(eng, NC, 512)
tr.set_params(params)
feats = torch.from_numpy(synthetic.mfcc(5 + 1000 * rank, B * T)).to(dev)
lab = torch.from_numpy(np.random.default_rng(5 + rank).integers(0, NC, B).astype(np.int32)).to(dev)
grad = torch.zeros(tr.n_grad, dtype=torch.float32, device=dev)
comm = torch.cuda.Stream(dev, priority=-1 if os.environ.get("XVEC_COMM_PRIORITY", "") == "high" else 0)
stream = torch.cuda.current_stream(dev)


def step(buckets):
    if world > 1 and buckets != "none":
        tr.forward_backward_allreduce(feats, lab, B, T, grad, stream, comm, fine=(buckets == "fine"))
    else:
        tr.forward_backward(feats, lab, B, T, grad_dev=grad)
    tr.apply(1e-4, grad_dev=grad, grad_scale=1.0 / world)


for cfg in sys.argv[1:]:
    opts = dict(kv.split("=") for kv in cfg.split(","))
    eng.set_option("clusters", int(opts.get("clusters", 74)))
    res = []
    for _ in range(2):
        for _ in range(30):
            step(opts.get("buckets", "fine"))
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            step(opts.get("buckets", "fine"))
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1) / 200], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res.append(float(ms.item()))
    if rank == 0:
        print("%-32s NCCL_MAX_NCHANNELS=%s comm priority %s  ms/step %s" % (cfg, os.environ.get("NCCL_MAX_NCHANNELS", "-"), os.environ.get("XVEC_COMM_PRIORITY", "default"), " ".join("%.4f" % r for r in res)), flush=True)
tr.close()
eng.close()
if world > 1:
    dist.destroy_process_group()
