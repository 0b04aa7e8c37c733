"""CPU tests of the native host side of an extraction job (include/xvec_job.h, x-vector-kaldi-tf_b200/ark_job.py):
the striped ark reader against the Python restatement of make_embedding's rules, the stripe-boundary protocol on
adversarial archives, and the vector-ark / scp formatters against the writers pinned on the reference's bytes.
No kernel is launched; the device engine is the oracle stand-in of test_host_logic."""
import io
import os
import subprocess
import sys

import numpy as np
import pytest

from xvector_b200 import _native, ark_job, kaldi_io, models, synthetic

import test_host_logic as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _random_ark(path, seed, n_utt, lo=0, hi=400, double_every=0, poison=False):
    """Random archive; returns [(key, matrix)].  ``poison`` plants plausible matrix headers INSIDE payloads."""
    rng = np.random.default_rng(seed)
    utts = []
    with open(path, "wb") as f:
        for i in range(n_utt):
            n = int(rng.integers(lo, hi + 1))
            if rng.random() < 0.1:
                n = int(rng.integers(0, 30))
            m = rng.standard_normal((n, 23)).astype(np.float32)
            if poison and n >= 40:
                # "xy \0BFM \4<rows>\4<23>" as payload bytes, chained once: what a naive resynchronisation would bite on
                fake = b"xy \0BFM \4" + np.int32(3).tobytes() + b"\4" + np.int32(23).tobytes()
                raw = m.view(np.uint8).reshape(-1)
                pos = int(rng.integers(0, raw.shape[0] - len(fake) - 1))
                raw[pos:pos + len(fake)] = np.frombuffer(fake, np.uint8)
                raw[-6:] = np.frombuffer(b"Az09_-", np.uint8)       # payload tail that reads like the start of the next key
            key = "utt%05d-%s" % (i, "abcxyz"[: 1 + i % 6])
            if double_every and i % double_every == 3:
                kaldi_io.write_mat(f, m.astype(np.float64), key=key)
                m = m.astype(np.float64).astype(np.float32)
            else:
                kaldi_io.write_mat(f, m, key=key)
            utts.append((key, m))
    return utts


def _expected(utts, min_chunk, chunk):
    ok, fail = [], []
    for key, m in utts:
        plan = models.chunk_plan(m.shape[0], min_chunk, chunk)
        if plan is None:
            fail.append((key, 1 if m.shape[0] == 0 else 2, m.shape[0]))
        else:
            ok.append((key, m, [n for _, n in plan]))
    return ok, fail


def _drain(reader, base=0):
    """All batches of a started reader: per ok utterance (rows, seg lens, dst row), plus batch sizes."""
    reader.start(base)
    out, sizes = [], []
    while True:
        b = reader.next()
        if b is None:
            break
        assert b.utt_first_seg[0] == 0 and b.utt_first_seg[-1] == b.n_seg and b.seg_len.sum() == b.n_rows
        off = 0
        for u in range(b.n_utt):
            segs = b.seg_len[b.utt_first_seg[u]:b.utt_first_seg[u + 1]].tolist()
            used = sum(segs)
            out.append((b.feats[off:off + used].copy(), segs, int(b.utt_dst_row[u])))
            off += used
        sizes.append(b.n_rows)
        reader.release(b.slot)
    return out, sizes


@pytest.mark.parametrize("min_chunk,chunk,batch_frames,threads", [(25, 10000, 3000, 4), (25, 100, 700, 3), (1, -1, 400000, 1),
                                                                  (30, 64, 257, 2)])
def test_reader_matches_make_embeddings_rules(tmp_path, min_chunk, chunk, batch_frames, threads):
    path = str(tmp_path / "feats.ark")
    utts = _random_ark(path, 1, 300, double_every=7)
    ok, fail = _expected(utts, min_chunk, chunk)
    r = _native.ArkReader(path, 23, min_chunk, chunk, batch_frames, n_threads=threads, pinned=False)
    info = r.index()
    assert (info["n_entries"], info["n_ok"], info["n_fail"]) == (len(utts), len(ok), len(fail))
    assert info["stopped_at"] == -1 and info["next_marker_off"] == os.path.getsize(path)
    assert info["rows_used"] == sum(sum(s) for _, _, s in ok) and info["n_segments"] == sum(len(s) for _, _, s in ok)
    blob, off = r.keys()
    raw = blob.tobytes()
    assert [raw[off[i]:off[i + 1]].decode() for i in range(len(ok))] == [k for k, _, _ in ok]
    assert r.failures() == fail
    got, sizes = _drain(r, base=1000)
    assert len(got) == len(ok)
    for i, ((rows, segs, dst), (key, m, want_segs)) in enumerate(zip(got, ok)):
        assert segs == want_segs and dst == 1000 + i
        assert np.array_equal(rows, m[:sum(want_segs)]), key
    longest = max(sum(s) for _, _, s in ok)
    assert all(n <= max(batch_frames, longest) for n in sizes) and len(sizes) == info["n_batches"]
    r.close()


def _striped(path, world, min_chunk=25, chunk=10000, start=0, batch_frames=5000):
    """The stripe protocol of ark_job.open_striped_reader, run for all ranks in one process (no process group)."""
    size = os.path.getsize(path)
    readers, infos = [], []
    for rank in range(world):
        b, e = ark_job.stripe_bounds(start, size, rank, world)
        rd = _native.ArkReader(path, 23, min_chunk, chunk, batch_frames, byte_begin=b, byte_end=(-1 if rank == world - 1 else e),
                               begin_is_boundary=(rank == 0), n_threads=2, pinned=False)
        readers.append(rd)
        infos.append(rd.index())
    applied = [None] * world
    rounds = 0
    for _ in range(world + 1):
        rounds += 1
        table = [(i["next_marker_off"], i["next_key_off"]) for i in infos]
        changed = 0
        for rank in range(1, world):
            if table[rank - 1] != applied[rank]:
                before = (infos[rank]["next_marker_off"], infos[rank]["next_key_off"])
                infos[rank] = readers[rank].set_first(*table[rank - 1])
                applied[rank] = table[rank - 1]
                changed += (infos[rank]["next_marker_off"], infos[rank]["next_key_off"]) != before
        if not changed:
            break
    else:
        raise AssertionError("stripes did not settle")
    return readers, infos, rounds


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("poison", [False, True])
def test_stripes_partition_the_archive_exactly(tmp_path, world, poison):
    path = str(tmp_path / "feats.ark")
    utts = _random_ark(path, 2 + world, 120, lo=20, hi=300, double_every=11, poison=poison)
    ok, fail = _expected(utts, 25, 10000)
    readers, infos, _ = _striped(path, world)
    assert sum(i["n_entries"] for i in infos) == len(utts)
    keys, rows, base = [], [], 0
    for rd, info in zip(readers, infos):
        blob, off = rd.keys()
        raw = blob.tobytes()
        keys += [raw[off[i]:off[i + 1]].decode() for i in range(info["n_ok"])]
        got, _ = _drain(rd, base)
        assert [d for _, _, d in got] == list(range(base, base + info["n_ok"]))
        rows += [g[0] for g in got]
        base += info["n_ok"]
        rd.close()
    assert keys == [k for k, _, _ in ok]                       # every utterance exactly once, in file order
    assert all(np.array_equal(a, m) for a, (_, m, _) in zip(rows, ok))
    # the stripes are balanced by bytes: no stripe holds more than its share plus one utterance
    share = os.path.getsize(path) / world
    for info in infos[:-1]:
        assert info["next_marker_off"] >= info["first_marker_off"] or info["n_entries"] == 0
    assert max(i["rows_used"] for i in infos) * 92 <= share + 300 * 92 + 4096


def test_more_stripes_than_entries_and_a_start_offset(tmp_path):
    path = str(tmp_path / "feats.ark")
    with open(path, "wb") as f:
        f.write(b"\0" * 37)                                           # the stream starts behind something else
        start = f.tell()
        mats = [(k, synthetic.mfcc(7 + i, n)) for i, (k, n) in enumerate([("a", 300), ("bb", 40), ("ccc", 700)])]
        for k, m in mats:
            kaldi_io.write_mat(f, m, key=k)
    readers, infos, rounds = _striped(path, 8, start=start)
    assert sum(i["n_entries"] for i in infos) == 3 and sum(i["n_entries"] == 0 for i in infos) >= 5   # empty stripes hand their boundary on
    keys = []
    for rd, info in zip(readers, infos):
        blob, off = rd.keys()
        keys += [blob.tobytes()[off[i]:off[i + 1]].decode() for i in range(info["n_ok"])]
        rd.close()
    assert keys == ["a", "bb", "ccc"]
    rd = _native.ArkReader(path, 23, 25, -1, 1000, byte_begin=start, pinned=False)
    assert rd.index()["n_ok"] == 3
    rd.close()


def test_unknown_entries_and_wrong_width_are_reported(tmp_path):
    path = str(tmp_path / "mixed.ark")
    with open(path, "wb") as f:
        kaldi_io.write_mat(f, synthetic.mfcc(1, 50), key="a")
        marker = f.tell()
        f.write(b"t  [\n 1 2 3 ]\n")                                   # a text matrix: not this scanner's business
    rd = _native.ArkReader(path, 23, 25, -1, 1000, pinned=False)
    assert rd.index()["stopped_at"] == marker
    rd.close()
    reader, counts = ark_job.open_striped_reader(path, 0, 23, 25, -1, 1000, pinned=False, rank=0, world=1)
    assert reader is None and counts[0][4] == marker                  # -> the caller takes the general parser
    with open(path, "wb") as f:
        kaldi_io.write_mat(f, np.zeros((30, 24), np.float32), key="wide")
    rd = _native.ArkReader(path, 23, 25, -1, 1000, pinned=False)
    with pytest.raises(_native.XvecError, match="feature dim 24"):
        rd.index()
    rd.close()
    with pytest.raises(_native.XvecError, match="min_chunk_size"):
        _native.ArkReader(path, 23, 100, 50, 1000, pinned=False)
    path2 = str(tmp_path / "cut.ark")
    with open(path2, "wb") as f:
        kaldi_io.write_mat(f, synthetic.mfcc(1, 50), key="a")
        f.truncate(f.tell() - 10)
    rd = _native.ArkReader(path2, 23, 25, -1, 1000, pinned=False)
    assert rd.index()["stopped_at"] == 0                               # a truncated payload is left to the parser's error
    rd.close()


def test_formatters_emit_the_reference_writers_bytes(tmp_path):
    rng = np.random.default_rng(5)
    keys = ["utt%d_%s" % (i, "x" * (i % 5)) for i in range(9000)]
    vecs = rng.standard_normal((len(keys), 512)).astype(np.float32)
    want = io.BytesIO()
    for k, v in zip(keys[:200], vecs[:200]):
        kaldi_io.write_vec_flt(want, v, key=k)                          # pinned on the reference's writer (tests/golden)
    blob = np.frombuffer("".join(keys).encode(), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(k) for k in keys])]).astype(np.int64)
    got = _native.vec_ark_format(blob, off[:201], vecs[:200])
    assert got.tobytes() == want.getvalue()
    # a window in the middle of the key table, many entries (threaded formatting), markers + scp lines
    lo, hi = 300, 9000
    a_ark, a_scp = str(tmp_path / "a.ark"), str(tmp_path / "a.scp")
    b_ark, b_scp = str(tmp_path / "b.ark"), str(tmp_path / "b.scp")
    with kaldi_io.ArkScpWriter(a_ark, a_scp, scp_ark_name="x.ark") as w:
        w.write_vec_entries(keys[:lo], vecs[:lo])
        w.write_vec_entries(keys[lo:hi], vecs[lo:hi])
    with kaldi_io.ArkScpWriter(b_ark, b_scp, scp_ark_name="x.ark") as w:
        w.write_vec_block(blob, off[:lo + 1], vecs[:lo])
        w.write_vec_block(blob, off[lo:hi + 1], vecs[lo:hi])
    assert open(a_ark, "rb").read() == open(b_ark, "rb").read()
    assert open(a_scp, "rb").read() == open(b_scp, "rb").read()
    assert _native.vec_ark_format(blob, off[:1], vecs[:0]).shape[0] == 0


def test_make_embedding_over_a_file_equals_the_stream_path(T_model_dir, tmp_path, monkeypatch, caplog):
    import logging
    d, engines = T_model_dir
    monkeypatch.setenv("XVEC_BATCH_FRAMES", "300")
    utts = T._utts()
    data = T._ark(utts)
    want = io.BytesIO()
    models.Model().make_embedding(io.BytesIO(data), want, d, 25, 100, True, None)          # general path (in-memory stream)
    path = tmp_path / "feats.ark"
    path.write_bytes(data)
    logger = logging.getLogger("test_ark_job")
    for source in ("fileobj", "name", "ark:name"):
        got = io.BytesIO()
        with caplog.at_level(logging.INFO, logger="test_ark_job"):
            caplog.clear()
            if source == "fileobj":
                with open(path, "rb") as f:
                    models.Model().make_embedding(f, got, d, 25, 100, True, logger)
            else:
                models.Model().make_embedding(("ark:" if source == "ark:name" else "") + str(path), got, d, 25, 100, True, logger)
        assert got.getvalue() == want.getvalue(), source
        text = caplog.text
        assert "Processed 7 features of average size" in text and "Done 5 and failed 2" in text
        assert "Zero-length utterance: 'empty'" in text and "number of rows in utterance: tooshort" in text
    assert engines[-1].utts_calls >= 2                                 # ... and it was the native job path that ran
    # ark,scp pair
    a, s = str(tmp_path / "x.ark"), str(tmp_path / "x.scp")
    with kaldi_io.open_vector_writer("ark,scp:%s,%s" % (a, s)) as w, open(path, "rb") as f:
        models.Model().make_embedding(f, w, d, 25, 100, True, None)
    assert open(a, "rb").read() == want.getvalue()
    assert [k for k, _ in kaldi_io.read_vec_flt_scp(s)] == ["a", "b", "c", "d", "e"]
    # XVEC_NATIVE_READER=0 keeps the general path
    monkeypatch.setenv("XVEC_NATIVE_READER", "0")
    n_before = engines[-1].utts_calls if engines else 0
    got = io.BytesIO()
    with open(path, "rb") as f:
        models.Model().make_embedding(f, got, d, 25, 100, True, None)
    assert got.getvalue() == want.getvalue() and engines[-1].utts_calls == 0


@pytest.fixture()
def T_model_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("XVEC_SEED", "11")
    d = str(tmp_path / "model_0")
    models.ModelWithoutDropoutTdnn().build_model(7, 23, d, None)
    engines = []

    def fake(meta, params, device):
        engines.append(T.OracleEngine(meta, params))
        return engines[-1]

    monkeypatch.setattr(models, "_create_engine", fake)
    return d, engines


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
engines = []
def fake(meta, params, device):
    engines.append(T.OracleEngine(meta, params))
    return engines[-1]
models._create_engine = fake
os.environ["XVEC_BATCH_FRAMES"] = "300"
out = io.BytesIO() if rank == 0 else None
with open(%(feats)r, "rb") as f:
    models.Model().make_embedding(f, out, %(model)r, 25, 100, True, logging.getLogger("w"))
assert engines[-1].utts_calls >= 1, "rank %%d did not take the native job path" %% rank
if rank == 0:
    open(%(out)r, "wb").write(out.getvalue())
dist.barrier()
dist.destroy_process_group()
"""


def test_three_rank_gloo_job_over_a_striped_file_is_byte_identical(T_model_dir, tmp_path, monkeypatch):
    d, _ = T_model_dir
    monkeypatch.setenv("XVEC_BATCH_FRAMES", "300")
    feats = str(tmp_path / "feats.ark")
    utts = _random_ark(feats, 9, 40, lo=20, hi=260)
    single = io.BytesIO()
    models.Model().make_embedding(io.BytesIO(open(feats, "rb").read()), single, d, 25, 100, True, None)
    out_path = str(tmp_path / "striped.ark")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % dict(root=ROOT, model=d, out=out_path, feats=feats))
    env = dict(os.environ, XVEC_SEED="11")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3",
                        "--master-addr", "127.0.0.1", "--master-port", "29741", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert open(out_path, "rb").read() == single.getvalue()          # same keys, same order, same bytes


@pytest.mark.parametrize("poison", [False, True])
def test_parallel_index_of_a_large_stripe_equals_the_sequential_chain(tmp_path, poison):
    # stripes of more than a few MB are indexed by several threads from resynchronised sub-range starts and stitched
    path = str(tmp_path / "big.ark")
    utts = _random_ark(path, 21, 700, lo=100, hi=900, double_every=13, poison=poison)
    assert os.path.getsize(path) > 24 << 20
    ok, fail = _expected(utts, 25, 400)
    seq = _native.ArkReader(path, 23, 25, 400, 60000, n_threads=1, pinned=False)
    par = _native.ArkReader(path, 23, 25, 400, 60000, n_threads=6, pinned=False)
    a, b = seq.index(), par.index()
    assert a == b and a["n_ok"] == len(ok) and a["n_fail"] == len(fail)
    ka, kb = seq.keys(), par.keys()
    assert np.array_equal(ka[0], kb[0]) and np.array_equal(ka[1], kb[1])
    got, sizes = _drain(par)
    assert all(np.array_equal(rows, m[:sum(segs)]) and segs == want for (rows, segs, _), (_, m, want) in zip(got, ok))
    assert sizes[0] <= 60000                                   # (ramp-up only applies to batch_frames >= 131072)
    seq.close(); par.close()
    # two ranks, each with a parallel index of its stripe
    readers, infos, _ = _striped(path, 2, chunk=400, batch_frames=200000)
    assert sum(i["n_ok"] for i in infos) == len(ok)
    keys = []
    for rd, info in zip(readers, infos):
        blob, off = rd.keys()
        keys += [blob.tobytes()[off[i]:off[i + 1]].decode() for i in range(info["n_ok"])]
        got, sizes = _drain(rd)
        assert sizes[0] <= 200000 // 8 + 900 and sizes[1] <= 200000 // 4 + 900       # ramp-up: 1/8, 1/4, 1/2, full
        rd.close()
    assert keys == [k for k, _, _ in ok]


_WORKER_FILES = r"""
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
m = models.Model()
if rank == 0:
    with kaldi_io.open_vector_writer("ark,scp:%(ark)s,%(scp)s") as w:
        w.write_vec_entries(["preamble"], [np.arange(512, dtype=np.float32)])      # the job starts behind earlier entries
        m.make_embedding(%(feats)r, w, %(model)r, 25, 100, True, logging.getLogger("w"))
        w.write_vec_entries(["postscript"], [np.ones(512, dtype=np.float32)])      # ... and the writer goes on behind it
    with open(%(plain)r, "wb") as f:
        m.make_embedding(%(feats)r, f, %(model)r, 25, 100, True, None)
else:
    m.make_embedding(%(feats)r, None, %(model)r, 25, 100, True, None)
    m.make_embedding(%(feats)r, None, %(model)r, 25, 100, True, None)
assert m.last_job_stats["output"] == "shared_file", m.last_job_stats
dist.barrier()
dist.destroy_process_group()
"""


def test_every_rank_writes_its_byte_range_of_the_one_output_file(T_model_dir, tmp_path, monkeypatch):
    d, _ = T_model_dir
    monkeypatch.setenv("XVEC_BATCH_FRAMES", "300")
    feats = str(tmp_path / "feats.ark")
    _random_ark(feats, 10, 50, lo=20, hi=260)
    names = dict(ark=str(tmp_path / "x.ark"), scp=str(tmp_path / "x.scp"), plain=str(tmp_path / "plain.ark"))
    ref_ark, ref_scp = str(tmp_path / "ref.ark"), str(tmp_path / "ref.scp")
    with kaldi_io.ArkScpWriter(ref_ark, ref_scp, scp_ark_name=names["ark"]) as w:
        w.write_vec_entries(["preamble"], [np.arange(512, dtype=np.float32)])
        models.Model().make_embedding(io.BytesIO(open(feats, "rb").read()), w, d, 25, 100, True, None)   # one process, general path
        w.write_vec_entries(["postscript"], [np.ones(512, dtype=np.float32)])
    script = tmp_path / "worker.py"
    script.write_text(_WORKER_FILES % dict(root=ROOT, model=d, feats=feats, **names))
    env = dict(os.environ, XVEC_SEED="11")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3",
                        "--master-addr", "127.0.0.1", "--master-port", "29743", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert open(names["ark"], "rb").read() == open(ref_ark, "rb").read()
    assert open(names["scp"], "rb").read() == open(ref_scp, "rb").read()
    body = open(ref_ark, "rb").read()
    pre = len("preamble") + 11 + 2048
    assert open(names["plain"], "rb").read() == body[pre:len(body) - (len("postscript") + 11 + 2048)]
    got = dict(kaldi_io.read_vec_flt_scp(names["scp"]))
    assert len(got) > 10 and np.array_equal(got["postscript"], np.ones(512, np.float32))


def test_reader_can_round_rows_to_float16_on_the_way_into_the_batches(tmp_path):
    # feats_f16: IEEE round-to-nearest-even, the bits numpy's float16 cast gives -- for float32 and float64 payloads, any
    # alignment of the destination rows, vector and scalar converter alike
    path = str(tmp_path / "feats.ark")
    utts = _random_ark(path, 31, 200, lo=0, hi=700, double_every=5)
    ok, _ = _expected(utts, 25, 300)
    r = _native.ArkReader(path, 23, 25, 300, 9000, n_threads=3, pinned=False, feats_f16=True)
    r.index()
    got, _ = _drain(r)
    assert len(got) == len(ok) and got[0][0].dtype == np.float16
    for (rows, segs, _), (key, m, want_segs) in zip(got, ok):
        assert segs == want_segs
        assert np.array_equal(rows.view(np.uint16), m[:sum(segs)].astype(np.float16).view(np.uint16)), key
    r.close()
    lib = _native.load_library()
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.standard_normal(40001).astype(np.float32) * 60,
                        np.array([0, -0.0, 65504, 65519.9, 65520, 1e9, -1e9, np.inf, 6.1e-5, 6e-8, 2.9802322e-8, 3e-8, 1e-10], np.float32),
                        rng.integers(0, 2 ** 32, 100000, dtype=np.uint64).astype(np.uint32).view(np.float32)])
    x = x[~np.isnan(x)]
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    for fn in (lib.xv_convert_f32_to_f16_host, lib.xv_convert_f32_to_f16_host_scalar):
        for off in (0, 1, 5):
            out = np.zeros(len(x) + 16, np.uint16)
            fn(x.ctypes.data, out.ctypes.data + 2 * off, len(x))
            assert np.array_equal(out[off:off + len(x)], want)


def test_scp_line_offsets_are_known_before_a_line_is_formatted():
    # what lets every rank write its own byte range of the one scp: line lengths from key lengths and the decimal digits
    # of the marker offsets alone -- checked against the formatter across every power of ten an offset can cross
    from xvector_b200 import ark_job
    rng = np.random.default_rng(5)
    entry = 11 + 4 * 512
    for base in (0, 1, 7, 93, 998, 9_990, 99_000, 999_999, 9_999_000, 99_990_000, 10 ** 10 - 5000, 10 ** 12 + 3):
        n = 37
        klens = rng.integers(1, 30, n)
        keys = ["".join(chr(97 + int(c)) for c in rng.integers(0, 26, k)) for k in klens]
        blob = np.frombuffer("".join(keys).encode(), np.uint8)
        off = np.concatenate([[0], np.cumsum(klens)]).astype(np.int64)
        markers = off[:-1] + np.arange(n) * entry + klens + 1
        lines = _native.scp_format(blob, off, "/some/where/x.ark", base, markers)
        cum = ark_job.scp_line_offsets(off, "/some/where/x.ark", base, entry)
        assert int(cum[-1]) == lines.shape[0]
        text = lines.tobytes().decode().splitlines(keepends=True)
        assert [len(t) for t in text] == np.diff(cum).tolist()
        assert text[3] == "%s /some/where/x.ark:%d\n" % (keys[3], base + markers[3])
