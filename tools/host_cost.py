import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from xvector_b200 import _native, synthetic
import bench
topo = bench.TOPOLOGIES["ModelWithoutDropoutTdnn"]
params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], weight_set="B")
eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], 512, 23, device=0)
eng.set_params(params)
lens = np.full(256, 400, np.int32)
fh = [torch.from_numpy(synthetic.mfcc_batch(2, lens)).pin_memory() for _ in range(2)]
eh = [torch.empty((256, 512)).pin_memory() for _ in range(2)]
def run(n):
    ts, tc = 0.0, 0.0
    prev = None
    t0 = time.perf_counter()
    for i in range(n):
        a = time.perf_counter()
        t = eng.submit_host_utts(fh[i & 1], lens, out_host=eh[i & 1])
        b = time.perf_counter()
        if prev is not None:
            eng.collect(prev)
        c = time.perf_counter()
        ts += b - a; tc += c - b
        prev = t
    eng.collect(prev)
    tot = time.perf_counter() - t0
    return tot / n * 1e3, ts / n * 1e3, tc / n * 1e3
run(10)
for _ in range(5):
    print("e2e ms/step %.4f  submit host ms %.4f  collect wait ms %.4f" % run(100))
