#!/usr/bin/env python
"""Product-path training throughput: Model.train_one_iteration over an egs tar (float16 minibatches of 64 x 400 frames, the
reference's on-disk format) -- tar member decode, host staging, H2D, training step, Adam -- against the device-resident step.
usage: bench_train_e2e.py [n_minibatches]"""
import logging
import os
import sys
import tempfile
import time
from types import SimpleNamespace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xvector_b200 import examples_io, models, synthetic          # noqa: E402

n_mb = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B, T, NC = 64, 400, 5000
tmp = tempfile.mkdtemp(prefix="xvec_train_e2e_")
os.environ["XVEC_SEED"] = "3"
logger = logging.getLogger("bench_train_e2e")
logger.setLevel(logging.WARNING)
models.ModelWithoutDropoutTdnn().build_model(NC, 23, os.path.join(tmp, "model_0"), None)
rng = np.random.default_rng(0)
base = synthetic.mfcc(1, B * T).reshape(B, T, 23).astype(np.float16)
t0 = time.time()
examples_io.write_egs_tar(os.path.join(tmp, "egs.1.tar"), [base] * n_mb, [rng.integers(0, NC, B) for _ in range(n_mb)])
print("wrote %d minibatches (%.0f MB) in %.1f s" % (n_mb, os.path.getsize(os.path.join(tmp, "egs.1.tar")) / 1e6, time.time() - t0))
args = SimpleNamespace(learning_rate=1e-4, print_interval=100, dropout_proportion=0.0, input_dir=os.path.join(tmp, "model_0"),
                       output_dir=os.path.join(tmp, "model_1"), random_seed=0)
t0 = time.time()
st = models.ModelWithoutDropoutTdnn().train_one_iteration(examples_io.TarFileDataLoader(os.path.join(tmp, "egs.1.tar"), queue_size=16), args, logger)
wall = time.time() - t0
print("train_one_iteration: %d minibatches in %.2f s wall (loop %.2f s incl. model load/save): %.2f ms/minibatch in the loop, %.1f M frames/s"
      % (n_mb, wall, st["elapsed"], 1e3 * st["elapsed"] / n_mb, n_mb * B * T / st["elapsed"] / 1e6))
