#!/usr/bin/env python
"""Product-path check on N GPUs of one box: extract_embedding.py (ark file in) single-process vs `torchrun --nproc-per-node N` must
give byte-identical outputs -- with `ark,scp:` files (every rank writes its byte ranges of the one ark and the one scp) and with a
pipe wspecifier (rows to rank 0's peer-memory table over NVLink, rank 0 writes the stream).
usage: python tools/multi_gpu_extract_check.py [N] [n_utts]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xvector_b200 import kaldi_io, synthetic          # noqa: E402
from xvector_b200.models import ModelWithoutDropoutTdnn  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
tmp = tempfile.mkdtemp(prefix="xvec_multi_")
os.environ["XVEC_SEED"] = "5"
model_dir = os.path.join(tmp, "model_0")
ModelWithoutDropoutTdnn().build_model(10, 23, model_dir, None)
lens = synthetic.lengths_uniform(4, n_utts)
lens[7], lens[11] = 10, 0                              # one too short, one empty: skipped on every rank alike
feats = synthetic.mfcc_batch(4, lens)
ark = os.path.join(tmp, "feats.ark")
with open(ark, "wb") as f:
    off = 0
    for i, n in enumerate(lens):
        kaldi_io.write_mat(f, feats[off:off + n], key="utt%07d" % i)
        off += n
cli = os.path.join(ROOT, "x-vector-kaldi-tf_b200", "extract_embedding.py")
common = ["--use-gpu=yes", "--min-chunk-size=25", "--chunk-size=10000", "--feature-rspecifier=ark:%s" % ark, "--model-dir=%s" % model_dir]
outs = {}
for tag, launcher in (("1gpu", [sys.executable]),
                      ("%dgpu" % n_gpus, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_gpus),
                                          "--master-addr", "127.0.0.1", "--master-port", "29611"])):
    o_ark, o_scp = os.path.join(tmp, tag + ".ark"), os.path.join(tmp, tag + ".scp")
    t0 = time.time()
    r = subprocess.run(launcher + [cli] + common + ["--vector-wspecifier=ark,scp:%s,%s" % (o_ark, o_scp)], capture_output=True, text=True)
    dt = time.time() - t0
    if r.returncode != 0:
        print(r.stderr[-3000:])
        raise SystemExit("%s run failed" % tag)
    outs[tag] = open(o_ark, "rb").read()
    n_out = sum(1 for _ in kaldi_io.read_vec_flt_scp(o_scp))
    print("%s: %d vectors, %d frames, wall %.2f s (incl. process start)" % (tag, n_out, int(lens.sum()), dt), flush=True)
a, b = outs["1gpu"], outs["%dgpu" % n_gpus]
assert a == b, "arks differ between 1 and %d GPUs" % n_gpus
assert open(os.path.join(tmp, "1gpu.scp")).read().replace("1gpu.ark", "X") == \
    open(os.path.join(tmp, "%dgpu.scp" % n_gpus)).read().replace("%dgpu.ark" % n_gpus, "X"), "scp files differ"
print("byte-identical ark + scp from 1 and %d GPUs (%d bytes)" % (n_gpus, len(a)))
# pipe output: the stream is written by rank 0 from its peer-memory result table
piped = os.path.join(tmp, "piped.ark")
r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_gpus), "--master-addr",
                    "127.0.0.1", "--master-port", "29612", cli] + common + ["--vector-wspecifier=| cat > %s" % piped],
                   capture_output=True, text=True)
if r.returncode != 0:
    print(r.stderr[-3000:])
    raise SystemExit("pipe-output run failed")
assert open(piped, "rb").read() == a, "pipe output differs"
print("byte-identical ark through a pipe wspecifier on %d GPUs (peer-memory table)" % n_gpus)
