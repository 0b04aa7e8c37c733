#!/usr/bin/env python
"""Host-side throughput of the extractor's ark reader alone (no kernel runs): Model._read_batches over a synthetic
configs[2]-like ark file (200-1000 frames per utterance) into the staging buffers, the batches being dropped by a null
consumer.  XVEC_READER_THREADS=1 is the sequential stream parser, > 1 the native header index + grouped pread jobs."""
import os
import queue
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401  (pinned staging buffers when a GPU is present)
from xvector_b200 import kaldi_io, models, synthetic  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else "/tmp/xvec_reader_bench.ark"
n_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
if not os.path.exists(path):
    lens = synthetic.lengths_uniform(3, n_utts)
    with open(path, "wb") as f:
        for i, n in enumerate(lens):
            kaldi_io.write_mat(f, synthetic.mfcc(1000 + i, int(n)), key="utt%07d" % i)
m = models.Model.__new__(models.Model)
staging = models._Staging(23, 400000)
work = queue.Queue(maxsize=3)
counters = dict(total_segments=0, total_segments_len=0, num_fail=0, num_success=0)


def consume():
    while True:
        b = work.get()
        if b is None:
            break
        staging.release(b.slot)


t = threading.Thread(target=consume)
t.start()
with open(path, "rb") as f:                                      # warm-up: page cache, staging buffers, thread pool
    m._read_batches(f, staging, work, counters, 25, 10000, 400000, 0, 1, None)
counters.update(total_segments_len=0, num_success=0)
t0 = time.perf_counter()
for _ in range(5):
    with open(path, "rb") as f:
        m._read_batches(f, staging, work, counters, 25, 10000, 400000, 0, 1, None)
dt = time.perf_counter() - t0
work.put(None)
t.join()
print("reader threads %s: %.1f M frames/s = %.2f GB/s (%d utterances, %.2f s)" % (
    os.environ.get("XVEC_READER_THREADS", "default"), counters["total_segments_len"] / dt / 1e6,
    counters["total_segments_len"] * 92 / dt / 1e9, counters["num_success"], dt))
