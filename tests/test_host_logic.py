"""CPU tests of the host side: make_embedding's per-utterance semantics, the extract_embedding CLI
(tmp-file / rename protocol), sharding + gather over gloo with world_size 2, and the C-ABI library's
exported symbols.  No test here launches a kernel: the device engine is replaced by a stand-in that
answers with the CPU oracle (tests are the one place allowed to call the oracle).
"""
import io
import logging
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import xvector_oracle as orc
from xvector_b200 import extract_embedding, kaldi_io, models, sharding, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOPOLOGY = "ModelWithoutDropoutTdnn"


class OracleEngine(object):
    """Stand-in for _native.XvecEngine with the same submit/collect protocol (two slots)."""

    def __init__(self, meta, params):
        self.meta, self.params = meta, params
        self.in_flight = {}
        self.next_ticket = 0
        self.calls = 0
        self.utts_calls = 0

    def _run(self, feats, lens, out):
        topo = dict(kernel_sizes=self.meta["kernel_sizes"], dilations=self.meta["dilation_rates"],
                    layer_sizes=self.meta["layer_sizes"], embedding_sizes=self.meta["embedding_sizes"])
        off = 0
        for i, n in enumerate(lens):
            out[i] = orc.forward(feats[off:off + n], self.params, topo).astype(np.float32)
            off += n

    def submit_host(self, feats, lens, emb):
        assert len(self.in_flight) < 2, "more than two submissions in flight"
        t = self.next_ticket % 2
        assert t not in self.in_flight
        self.next_ticket += 1
        self.calls += 1
        self.in_flight[t] = (np.array(feats, copy=True), np.array(lens), emb)
        return t

    def submit_host_utts(self, feats, lens, utt_first_seg=None, dst_rows=None, out_dev=None, out_host=None):
        assert out_dev is None and dst_rows is None, "the stand-in has no peer memory"
        self.utts_calls += 1
        seg = np.empty((len(lens), 512), np.float32)
        t = self.submit_host(feats, lens, seg)
        first = np.arange(len(lens) + 1) if utt_first_seg is None else np.array(utt_first_seg)
        self.in_flight[t] = self.in_flight[t] + (first, out_host)
        return t

    def collect(self, ticket):
        item = self.in_flight.pop(ticket)
        feats, lens, emb = item[:3]
        self._run(feats, lens, emb)
        if len(item) == 5:          # utterance-level: the reference's float32 chunk average (models.py:398-421)
            first, out_host = item[3], item[4]
            utts = [(u, None, int(first[u]), [int(n) for n in lens[first[u]:first[u + 1]]]) for u in range(len(first) - 1)]
            out_host[...] = models._average_chunks(emb, utts)

    def extract_host(self, feats, lens, emb=None):
        emb = np.empty((len(lens), 512), np.float32) if emb is None else emb
        self._run(np.asarray(feats), np.asarray(lens), emb)
        return emb


@pytest.fixture()
def model_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("XVEC_SEED", "11")
    d = str(tmp_path / "model_0")
    models.ModelWithoutDropoutTdnn().build_model(7, 23, d, None)
    engines = []

    def fake(meta, params, device):
        engines.append(OracleEngine(meta, params))
        return engines[-1]

    monkeypatch.setattr(models, "_create_engine", fake)
    return d, engines


def _ark(utts):
    buf = io.BytesIO()
    for k, m in utts.items():
        kaldi_io.write_mat(buf, m, key=k)
    return buf.getvalue()


def _utts():
    lens = dict(a=130, tooshort=10, b=60, empty=0, c=101, d=25, e=299)
    return {k: (synthetic.mfcc(40 + i, n) if n else np.zeros((0, 23), np.float32)) for i, (k, n) in enumerate(lens.items())}


def test_build_model_writes_the_reference_directory_contract(model_dir):
    d, _ = model_dir
    from xvector_b200.ze_utils import is_correct_model_dir
    assert is_correct_model_dir(d)                                   # non-empty model.meta + done (ze_utils.py:561-567)
    w = models.Model().get_models_weights(d)
    assert w["frame_level_info_layer-1/w:0"].shape == (3, 512, 512)  # tdnn: taps 3 (models.py:545)
    assert w["frame_level_info_layer-4/w:0"].shape == (1, 512, 1536)
    assert w["embed_layer-0/w:0"].shape == (3072, 512) and w["output/w:0"].shape == (512, 7)
    assert np.allclose(w["frame_level_info_layer-0/b:0"], 0.1) and np.allclose(w["frame_level_info_layer-0/variance:0"], 1.0)
    assert np.abs(w["frame_level_info_layer-2/w:0"]).max() <= 0.2 + 1e-6   # truncated normal, 2 sigma of 0.1


@pytest.mark.parametrize("batch_frames", ["100", "400000"])
def test_make_embedding_semantics_match_the_reference_loop(model_dir, monkeypatch, caplog, batch_frames):
    d, engines = model_dir
    monkeypatch.setenv("XVEC_BATCH_FRAMES", batch_frames)            # tiny batches exercise the 2-deep pipeline
    utts = _utts()
    out = io.BytesIO()
    logger = logging.getLogger("test_host_logic")
    with caplog.at_level(logging.INFO, logger="test_host_logic"):
        models.Model().make_embedding(io.BytesIO(_ark(utts)), out, d, 25, 100, True, logger)
    got = dict(kaldi_io.read_vec_flt_ark(io.BytesIO(out.getvalue())))
    assert list(got) == ["a", "b", "c", "d", "e"]                    # skipped: < min_chunk_size and empty; order kept
    params = models.Model().get_models_weights(d)
    for k, v in got.items():
        want = orc.make_embedding_one(utts[k], params, TOPOLOGY, 25, 100)   # chunks of 100, tail < 25 dropped
        assert v.dtype == np.float32 and v.shape == (512,)
        assert orc.parity_metrics(v, want)["max_rel"] < 1e-5
    text = caplog.text
    assert "Processed 7 features of average size" in text and "Done 5 and failed 2" in text   # models.py:425-427
    assert "Total time for neural network computations is" in text
    assert "Elapsed time for extracting whole embeddings is" in text
    if batch_frames == "100":
        assert engines[0].calls >= 4


def test_topology_variants_keep_the_reference_variable_names(tmp_path, monkeypatch):
    monkeypatch.setenv("XVEC_SEED", "3")
    for cls, act in ((models.ModelWithoutDropoutPRelu, "prelu"), (models.ModelL2LossWithoutDropoutLRelu, "lrelu"),
                     (models.ModelL2LossWithoutDropoutReluHeInit, "relu"), (models.ModelL2LossWithoutDropoutPRelu, "prelu")):
        d = str(tmp_path / cls.__name__)
        cls().build_model(5, 23, d, None)
        m = models.Model()
        w = m.get_models_weights(d)
        assert m.meta["activation"] == act and m.meta["model_class"] == cls.__name__
        assert ("frame_level_info_layer-3/prelu/prelu:0" in w) == (act == "prelu")       # tf_block.py:40-46
        if act == "prelu":
            assert np.allclose(w["frame_level_info_layer-0/prelu/prelu:0"], 0.1)           # constant_initializer(0.1)
        if cls is models.ModelL2LossWithoutDropoutReluHeInit:                               # he_normal, models.py:1160
            assert abs(w["frame_level_info_layer-1/w:0"].std() - np.sqrt(2.0 / (5 * 512)) * 0.88) < 0.002
            assert not np.allclose(w["frame_level_info_layer-1/b:0"], 0.1)


def test_chunk_plan_edges():
    cp = models.chunk_plan
    assert cp(0, 25, 10000) is None and cp(24, 25, 10000) is None
    assert cp(25, 25, 10000) == [(0, 25)]
    assert cp(10000, 25, 10000) == [(0, 10000)]
    assert cp(10001, 25, 10000) == [(0, 10000)]                      # 1-frame tail < min_chunk_size is dropped
    assert cp(25000, 25, 10000) == [(0, 10000), (10000, 10000), (20000, 5000)]
    assert cp(777, 25, -1) == [(0, 777)]
    for rows in (25, 99, 100, 101, 250, 10001):
        assert cp(rows, 25, 100) == orc.chunk_plan(rows, 25, 100)


def test_cli_flags_and_tmp_rename_protocol(model_dir, tmp_path):
    d, _ = model_dir
    utts = _utts()
    feats = tmp_path / "feats.ark"
    feats.write_bytes(_ark(utts))
    ark, scp = str(tmp_path / "xvector.1.ark"), str(tmp_path / "xvector.1.scp")
    argv = ["--use-gpu=no", "--min-chunk-size=25", "--chunk-size=10000", "--feature-rspecifier=ark:%s" % feats,
            "--vector-wspecifier=ark,scp:%s,%s" % (ark, scp), "--model-dir=%s" % d]
    args = extract_embedding.get_args(argv)
    assert (args.use_gpu, args.min_chunk_size, args.chunk_size) == ("no", 25, 10000)
    assert extract_embedding.process_wspecifier("| copy-vector ark:- ark,scp:%s,%s" % (ark, scp)) == \
        ("| copy-vector ark:- ark,scp:%s.tmp.ark,%s.tmp.scp" % (ark, scp), ark, scp)     # extract_embedding.py:94-108
    extract_embedding.eval_dnn(args)
    assert os.path.exists(ark) and os.path.exists(scp)
    assert not os.path.exists(ark + ".tmp.ark") and not os.path.exists(scp + ".tmp")
    assert os.path.exists(scp + ".tmp.scp")           # the reference leaves it too (its os.remove is commented out, :150)
    lines = open(scp).read().splitlines()
    assert [l.split()[0] for l in lines] == ["a", "b", "c", "d", "e"]
    assert all(l.split()[1].startswith(ark + ":") and ".tmp" not in l for l in lines)
    by_scp = dict(kaldi_io.read_vec_flt_scp(scp))
    by_ark = dict(kaldi_io.read_vec_flt_ark(ark))
    assert all(np.array_equal(by_scp[k], by_ark[k]) for k in by_ark)
    mtime = os.path.getmtime(ark)
    extract_embedding.eval_dnn(args)                                  # both outputs exist: return at once (:126-128)
    assert os.path.getmtime(ark) == mtime


def test_block_cyclic_dealing_of_a_stream():
    assert [sharding.block_cyclic_rank(i, 2, block=2) for i in range(8)] == [0, 0, 1, 1, 0, 0, 1, 1]


_WORKER = r"""
import io, os, sys, logging
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch.distributed as dist
from xvector_b200 import models, kaldi_io
import test_host_logic as T
dist.init_process_group(backend="gloo")
rank = dist.get_rank()
models._create_engine = lambda meta, params, device: T.OracleEngine(meta, params)
os.environ["XVEC_BATCH_FRAMES"] = "300"
utts = T._utts()
out = io.BytesIO() if rank == 0 else None
models.Model().make_embedding(io.BytesIO(T._ark(utts)), out, %(model)r, 25, 100, True, logging.getLogger("w"))
if rank == 0:
    open(%(out)r, "wb").write(out.getvalue())
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_extraction_is_byte_identical_to_one_rank(model_dir, tmp_path, monkeypatch):
    d, _ = model_dir
    monkeypatch.setenv("XVEC_BATCH_FRAMES", "300")
    utts = _utts()
    single = io.BytesIO()
    models.Model().make_embedding(io.BytesIO(_ark(utts)), single, d, 25, 100, True, None)
    out_path = str(tmp_path / "two_rank.ark")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % dict(root=ROOT, model=d, out=out_path))
    env = dict(os.environ, XVEC_SEED="11")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(out_path, "rb").read() == single.getvalue()          # same keys, same order, same bytes


def test_gather_to_rank0_single_process():
    idx = np.array([2, 0, 1])
    emb = np.arange(3 * 4, dtype=np.float32).reshape(3, 4)
    full = sharding.gather_to_rank0(idx, emb, 3, 4)
    assert np.array_equal(full[idx], emb)


def test_c_abi_library_exports_every_declared_symbol():
    from xvector_b200 import _native
    inc = os.path.join(ROOT, "include")
    header = "".join(open(os.path.join(inc, h)).read() for h in sorted(os.listdir(inc)) if h.endswith(".h"))
    assert "xv_frontend" in header and "xv_train_create" in header
    declared = sorted(set(re.findall(r"\b(xv_[a-z_0-9]+)\s*\(", header)))
    assert declared and set(declared) == set(_native.EXPORTED_SYMBOLS)
    path = _native.build_library()                                    # nvcc cross-compiles without a GPU
    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    assert not [s for s in declared if s not in exported]
    lib = _native.load_library()                                      # loads without a GPU; no compute call here
    assert b"sm_100a" in lib.xv_version()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "x-vector-kaldi-tf_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert "oracle" not in src.replace("the oracle", ""), "%s mentions the oracle package" % name


def test_native_ark_index_and_parallel_reader_match_the_stream_parser(model_dir, tmp_path, monkeypatch):
    """An ark in a regular file goes through xv_ark_scan (native header index over the mmap'ed file) and grouped pread
    jobs on the reader pool; the x-vectors written must be byte-identical to the in-memory stream path, also when the
    file holds float64, text and zero-row entries (the scanner hands the text entry and what follows to the parser)."""
    from xvector_b200 import _native
    _native.build_library()                                           # no-op when the in-tree library is up to date
    d, _ = model_dir
    monkeypatch.setenv("XVEC_BATCH_FRAMES", "300")
    monkeypatch.setattr(models._Batch, "GROUP_BYTES", 20000)          # several pool jobs per batch
    utts = _utts()
    buf = io.BytesIO()
    for i, (k, m) in enumerate(utts.items()):
        if k == "c":                                                  # a text-form matrix in the middle of the file
            buf.write((k + "  [\n" + "\n".join(" ".join("%.9g" % v for v in row) for row in m) + " ]\n").encode())
        else:
            kaldi_io.write_mat(buf, m.astype(np.float64) if k == "b" else m, key=k)
    data = buf.getvalue()
    path = tmp_path / "feats.ark"
    path.write_bytes(data)

    (key_off, key_len, rows, cols, elem, pay), consumed = _native.ark_scan(data)
    keys = [data[o:o + n].decode() for o, n in zip(key_off.tolist(), key_len.tolist())]
    assert keys == ["a", "tooshort", "b", "empty"]                    # stops in front of the text entry "c"
    assert rows.tolist() == [130, 10, 60, 0] and cols.tolist() == [23, 23, 23, 23] and elem.tolist() == [4, 4, 8, 4]
    assert data[consumed:consumed + 2] == b"c " and np.array_equal(
        np.frombuffer(data, "<f4", 130 * 23, int(pay[0])).reshape(130, 23), utts["a"])

    want = io.BytesIO()
    models.Model().make_embedding(io.BytesIO(data), want, d, 25, 100, True, None)
    for threads in ("4", "1"):
        monkeypatch.setenv("XVEC_READER_THREADS", threads)
        got = io.BytesIO()
        with open(path, "rb") as f:
            models.Model().make_embedding(f, got, d, 25, 100, True, None)
        assert got.getvalue() == want.getvalue() and len(got.getvalue()) > 5 * 512 * 4
    path.write_bytes(data[:len(data) // 3])                           # truncated payload: the parser's error surfaces
    monkeypatch.setenv("XVEC_READER_THREADS", "4")
    with open(path, "rb") as f, pytest.raises(Exception):
        models.Model().make_embedding(f, io.BytesIO(), d, 25, 100, True, None)


def test_adam_step_counter_survives_the_float32_underflow_of_beta1_power():
    # 0.9**t stored as float32 is 0.0 from t ~ 990 on; the counter must come from the explicit entry, else
    # from beta2_power, and a power that reads 0 is a saturated counter -- never step 0
    from xvector_b200.models import adam_step_from_checkpoint, ADAM_STEP_KEY
    f32 = lambda v: np.asarray([v], dtype=np.float32)
    for t in (0, 1, 12, 985, 1000, 5000, 40000):
        tf_style = {"beta1_power:0": f32(0.9 ** t), "beta2_power:0": f32(0.999 ** t)}
        got = adam_step_from_checkpoint(tf_style)
        assert abs(got - t) <= max(1, t // 2000), (t, got)                  # float32 beta2_power: exact to ~0.05 %
        assert adam_step_from_checkpoint(dict(tf_style, **{ADAM_STEP_KEY: np.asarray([t], np.int64)})) == t
    assert adam_step_from_checkpoint({"beta1_power:0": f32(0.0), "beta2_power:0": f32(0.0)}) >= 100000
    assert adam_step_from_checkpoint({"beta1_power:0": f32(0.0)}) >= 100000
    assert adam_step_from_checkpoint({}) == 0


def test_pooled_reader_never_lets_a_dropped_tail_overwrite_the_next_utterance(tmp_path, monkeypatch):
    # utterances whose tail chunk is dropped (< min_chunk_size) keep only `used` rows in the staging
    # buffer and the next utterance starts right behind them; payloads are fetched by concurrent pread jobs, so a job must
    # fetch the kept rows only.  40 utterances of 1050 rows, chunks of 1000, tail of 50 < 100 dropped, many small jobs.
    import queue
    import threading
    monkeypatch.setenv("XVEC_READER_THREADS", "4")
    monkeypatch.setattr(models._Batch, "GROUP_BYTES", 64 << 10)
    n_utt, rows = 40, 1050
    mats = [synthetic.mfcc(300 + i, rows) + np.float32(i) for i in range(n_utt)]
    path = str(tmp_path / "feats.ark")
    with open(path, "wb") as f:
        for i, m in enumerate(mats):
            kaldi_io.write_mat(f, m, key="utt%03d" % i)
    for trial in range(5):
        model = models.Model.__new__(models.Model)
        staging = models._Staging(23, 8000)
        work = queue.Queue(maxsize=3)
        counters = dict(total_segments=0, total_segments_len=0, num_fail=0, num_success=0)
        seen = []

        def consume():
            while True:
                b = work.get()
                if b is None:
                    return
                view = staging.view(b.slot, b.n_frames).copy()
                off = 0
                for idx, key, first, lengths in b.utts:
                    used = sum(lengths)
                    seen.append((idx, view[off:off + used]))
                    off += used
                staging.release(b.slot)

        th = threading.Thread(target=consume)
        th.start()
        try:
            with open(path, "rb") as f:
                model._read_batches(f, staging, work, counters, 100, 1000, 8000, 0, 1, None)
        finally:
            work.put(None)
            th.join()
        assert len(seen) == n_utt
        for idx, got in seen:
            assert got.shape == (1000, 23)
            assert np.array_equal(got, mats[idx][:1000]), "utterance %d corrupted in trial %d" % (idx, trial)
