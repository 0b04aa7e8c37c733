"""Kaldi ark I/O against fixtures produced by the reference's own kaldi_io (tests/golden)."""
import io
import os
import struct

import numpy as np
import pytest

from xvector_b200 import kaldi_io

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXPECTED = np.load(os.path.join(GOLD, "expected.npz"))


def _expected(fname):
    pre = fname + "/"
    return {k[len(pre):]: EXPECTED[k] for k in EXPECTED.files if k.startswith(pre)}


@pytest.mark.parametrize("fname", ["mat_f32.ark", "mat_f64.ark"])
def test_read_mat_ark_matches_reference_writer(fname):
    want = _expected(fname)
    got = list(kaldi_io.read_mat_ark(os.path.join(GOLD, fname)))
    assert [k for k, _ in got] == list(want.keys())          # order preserved, incl. the 0-row entry
    for k, m in got:
        assert m.dtype == want[k].dtype and m.shape == want[k].shape
        assert np.array_equal(m, want[k])


@pytest.mark.parametrize("fname", ["vec_f32.ark", "vec_f64.ark"])
def test_read_vec_ark_matches_reference_writer(fname):
    want = _expected(fname)
    got = dict(kaldi_io.read_vec_flt_ark(os.path.join(GOLD, fname)))
    assert list(got) == list(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and np.array_equal(got[k], want[k])


@pytest.mark.parametrize("fname,writer", [("mat_f32.ark", "mat"), ("mat_f64.ark", "mat"),
                                          ("vec_f32.ark", "vec"), ("vec_f64.ark", "vec")])
def test_writers_are_byte_identical_to_reference(fname, writer):
    want = _expected(fname)
    buf = io.BytesIO()
    buf.mode = "wb"
    for k, a in want.items():
        (kaldi_io.write_mat if writer == "mat" else kaldi_io.write_vec_flt)(buf, a, key=k)
    with open(os.path.join(GOLD, fname), "rb") as f:
        assert buf.getvalue() == f.read()


def test_vec_entry_bytes_helper_matches_writer():
    v = np.arange(5, dtype=np.float32)
    buf = io.BytesIO()
    buf.mode = "wb"
    kaldi_io.write_vec_flt(buf, v, key="k1")
    assert kaldi_io.vec_flt_entry_bytes(v, "k1") == buf.getvalue()


def test_compressed_matrix_decodes_like_reference():
    got = dict(kaldi_io.read_mat_ark(os.path.join(GOLD, "cm.ark")))
    want = np.load(os.path.join(GOLD, "cm_expected.npy"))
    assert got["cm_utt"].dtype == np.float32 and got["cm_utt"].shape == want.shape
    assert np.array_equal(got["cm_utt"], want)


def test_text_matrix_parses_like_reference():
    got = dict(kaldi_io.read_mat_ark(os.path.join(GOLD, "text.ark")))
    assert np.array_equal(got["t1"], np.load(os.path.join(GOLD, "text_expected.npy")))


def test_hand_assembled_entry_layout():
    # key ' ' \0B FM ' ' \4 rows \4 cols payload
    m = np.arange(6, dtype=np.float32).reshape(2, 3)
    blob = b"k-1 \0BFM \4" + struct.pack("<i", 2) + b"\4" + struct.pack("<i", 3) + m.tobytes()
    (key, got), = list(kaldi_io.read_mat_ark(io.BytesIO(blob)))
    assert key == "k-1" and np.array_equal(got, m)
    v = np.array([1.5, -2.0], np.float32)
    out = io.BytesIO()
    out.mode = "wb"
    kaldi_io.write_vec_flt(out, v, key="k-1")
    assert out.getvalue() == b"k-1 \0BFV \4" + struct.pack("<I", 2) + v.tobytes()


def test_key_format_is_enforced_and_eof_is_none():
    assert kaldi_io.read_key(io.BytesIO(b"")) is None
    with pytest.raises(AssertionError):
        kaldi_io.read_key(io.BytesIO(b"bad$key \0B"))
    # unbuffered stream without peek(): same result as the buffered fast path
    class Raw(io.RawIOBase):
        def __init__(self, b): self.b, self.i = b, 0
        def readable(self): return True
        def readinto(self, out):
            n = min(len(out), len(self.b) - self.i)
            out[:n] = self.b[self.i:self.i + n]; self.i += n
            return n
    r = Raw(b"utt1 rest")
    assert kaldi_io.read_key(r) == "utt1" and r.read(4) == b"rest"
    b = io.BufferedReader(Raw(b"utt1 rest"))
    assert kaldi_io.read_key(b) == "utt1" and b.read(4) == b"rest"


def test_unknown_headers_and_dtypes_raise():
    with pytest.raises(kaldi_io.UnknownMatrixHeader):
        list(kaldi_io.read_mat_ark(io.BytesIO(b"k \0BXM \4")))
    with pytest.raises(kaldi_io.UnknownVectorHeader):
        list(kaldi_io.read_vec_flt_ark(io.BytesIO(b"k \0BXV \4")))
    out = io.BytesIO()
    out.mode = "wb"
    with pytest.raises(kaldi_io.UnsupportedDataType):
        kaldi_io.write_vec_flt(out, np.zeros(3, np.int32), key="k")
    with pytest.raises(kaldi_io.UnsupportedDataType):
        kaldi_io.write_mat(out, np.zeros((2, 2), np.int32), key="k")


def test_open_or_fd_prefix_offset_gzip_and_pipes(tmp_path):
    m = np.arange(12, dtype=np.float32).reshape(3, 4)
    ark = tmp_path / "a.ark"
    with open(ark, "wb") as f:
        f.write(b"junk")
        off = f.tell() + len(b"u1 ")
        kaldi_io.write_mat(f, m, key="u1")
    # ark: prefix stripped; ':offset' seeks to the matrix body (what an scp line points at)
    assert np.array_equal(kaldi_io.read_mat("ark:%s:%d" % (ark, off)), m)
    scp = tmp_path / "a.scp"
    scp.write_text("u1 %s:%d\n" % (ark, off))
    assert np.array_equal(dict(kaldi_io.read_mat_scp(str(scp)))["u1"], m)
    # gzip
    import gzip
    gz = tmp_path / "b.ark.gz"
    with gzip.open(gz, "wb") as f:
        kaldi_io.write_mat(f, m, key="u2")
    assert np.array_equal(dict(kaldi_io.read_mat_ark(str(gz)))["u2"], m)
    # input pipe / output pipe
    with open(tmp_path / "c.ark", "wb") as f:
        kaldi_io.write_mat(f, m, key="u3")
    got = dict(kaldi_io.read_mat_ark("ark:cat %s |" % (tmp_path / "c.ark")))
    assert np.array_equal(got["u3"], m)
    out = tmp_path / "d.ark"
    fd = kaldi_io.open_or_fd("| cat > %s" % out)
    kaldi_io.write_vec_flt(fd, np.ones(4, np.float32), key="v")
    fd.close()
    import time
    for _ in range(100):
        if out.exists() and out.stat().st_size > 0:
            break
        time.sleep(0.05)
    assert np.array_equal(dict(kaldi_io.read_vec_flt_ark(str(out)))["v"], np.ones(4, np.float32))


@pytest.mark.parametrize("fname", ["mat_f32.ark", "mat_f64.ark", "cm.ark"])
def test_entry_reader_equals_read_mat_ark(fname):
    # the extractor's zero-copy reader must see exactly what read_mat_ark (reference kaldi_io.py:372-392) sees
    path = os.path.join(GOLD, fname)
    want = list(kaldi_io.read_mat_ark(path))
    got = []
    for e in kaldi_io.read_mat_ark_entries(path):
        assert (e.rows, e.cols) == (want[len(got)][1].shape if want[len(got)][1].ndim == 2 else (0, 0)) or e.rows == 0
        dst = np.full((e.rows, e.cols), np.nan, np.float32)
        e.read_into(dst)
        got.append((e.key, dst))
    assert [k for k, _ in got] == [k for k, _ in want]
    for (_, g), (_, w) in zip(got, want):
        assert np.array_equal(g, w.astype(np.float32).reshape(g.shape))
    # skipping payloads (utterances of other ranks / too short) keeps the stream aligned
    keys = [e.key for e in kaldi_io.read_mat_ark_entries(path)]
    assert keys == [k for k, _ in want]
    mixed = []
    for i, e in enumerate(kaldi_io.read_mat_ark_entries(path)):
        mixed.append(e.read() if i % 2 else None)
    for i, (_, w) in enumerate(want):
        if i % 2:
            assert np.array_equal(np.asarray(mixed[i], dtype=np.float64).reshape(w.shape), w.astype(np.float64))


def test_entry_reader_on_unbuffered_pipe_and_truncation():
    m = np.arange(7 * 23, dtype=np.float32).reshape(7, 23)
    buf = io.BytesIO()
    kaldi_io.write_mat(buf, m, key="a")
    kaldi_io.write_mat(buf, m * 2, key="b")

    class Dribble(io.RawIOBase):                 # a stream that returns at most 5 bytes per read, no peek()
        def __init__(self, data):
            self.data, self.pos = data, 0

        def readable(self):
            return True

        def readinto(self, b):
            n = min(5, len(b), len(self.data) - self.pos)
            b[:n] = self.data[self.pos:self.pos + n]
            self.pos += n
            return n

    got = {}
    for e in kaldi_io.read_mat_ark_entries(Dribble(buf.getvalue())):
        dst = np.empty((e.rows, e.cols), np.float32)
        e.read_into(dst)
        got[e.key] = dst
    assert np.array_equal(got["a"], m) and np.array_equal(got["b"], m * 2)
    with pytest.raises(kaldi_io.BadInputFormat):
        for e in kaldi_io.read_mat_ark_entries(io.BytesIO(buf.getvalue()[:-10])):
            e.read_into(np.empty((e.rows, e.cols), np.float32))


# ---- integer vectors, posteriors, confusion-network times, segments: pinned by fixtures the reference's own readers /
# writer produced (tests/golden/make_golden_misc.py) ----

def _misc():
    return np.load(os.path.join(GOLD, "misc_expected.npz"))


def test_int_vectors_match_the_reference_bytes_and_values(tmp_path):
    want = _misc()
    got = dict(kaldi_io.read_vec_int_ark(os.path.join(GOLD, "vec_int.ark")))
    assert list(got) == ["utt-a", "utt.b", "spk/utt_c"]
    for k, v in got.items():
        assert v.dtype == np.int32 and np.array_equal(v, want["vec_int/" + k])
    assert [k for k, _ in kaldi_io.read_ali_ark(os.path.join(GOLD, "vec_int.ark"))] == list(got)
    out = tmp_path / "rewritten.ark"
    with open(out, "wb") as f:
        for k, v in got.items():
            kaldi_io.write_vec_int(f, v, key=k)
    assert out.read_bytes() == open(os.path.join(GOLD, "vec_int.ark"), "rb").read()     # byte-identical to the reference's writer
    text = tmp_path / "ali.txt"
    text.write_bytes(b"u1 4 8 15 16 23 42\nu2  [ 1 2 3 ]\n")
    parsed = dict(kaldi_io.read_vec_int_ark(str(text)))
    assert parsed["u1"].tolist() == [4, 8, 15, 16, 23, 42] and parsed["u2"].tolist() == [1, 2, 3]


def test_posteriors_and_bin_times_match_the_reference_readers():
    want = _misc()
    posts = dict(kaldi_io.read_post_ark(os.path.join(GOLD, "post.ark")))
    assert list(posts) == ["p1", "p2"] and list(dict(kaldi_io.read_cnet_ark(os.path.join(GOLD, "post.ark")))) == ["p1", "p2"]
    for k, post in posts.items():
        assert [len(fr) for fr in post] == want["post/%s/lens" % k].tolist()
        assert [i for fr in post for i, _ in fr] == want["post/%s/idx" % k].tolist()
        assert np.array_equal(np.array([v for fr in post for _, v in fr], np.float64), want["post/%s/val" % k])
    for k, bins in kaldi_io.read_cntime_ark(os.path.join(GOLD, "cntime.ark")):
        assert np.array_equal(np.array(bins, np.float64), want["cntime/" + k])


def test_segments_as_bool_vec_matches_the_reference():
    got = kaldi_io.read_segments_as_bool_vec(os.path.join(GOLD, "segments.txt"))
    want = _misc()["segments"]
    assert got.dtype == bool and np.array_equal(got, want) and got.sum() == 45 + 40 + 17


def test_native_ark_scan_agrees_with_the_python_parser_on_random_arks():
    """xv_ark_scan (host-only, loads without a GPU) against read_mat_ark_entries over random archives: same keys, shapes,
    element sizes and payload bytes; a buffer cut anywhere yields exactly the entries that are whole."""
    from xvector_b200 import _native
    rng = np.random.default_rng(12)
    alphabet = "abcXYZ019._-/"
    for trial in range(20):
        buf = io.BytesIO()
        want = []
        for i in range(int(rng.integers(1, 12))):
            key = "".join(rng.choice(list(alphabet), size=int(rng.integers(1, 20))))
            rows, cols = int(rng.integers(0, 40)), int(rng.integers(1, 30))
            m = rng.standard_normal((rows, cols)).astype(np.float64 if rng.random() < 0.3 else np.float32)
            kaldi_io.write_mat(buf, m, key=key)
            want.append((key, m))
        data = buf.getvalue()
        (key_off, key_len, rows, cols, elem, pay), consumed = _native.ark_scan(data)
        assert consumed == len(data) and len(key_off) == len(want)
        parsed = [(e.key, e.read()) for e in kaldi_io.read_mat_ark_entries(io.BytesIO(data))]
        for i, (key, m) in enumerate(want):
            assert data[key_off[i]:key_off[i] + key_len[i]].decode() == key == parsed[i][0]
            assert (rows[i], cols[i], elem[i]) == (m.shape[0], m.shape[1], m.dtype.itemsize)
            got = np.frombuffer(data, m.dtype, m.size, int(pay[i])).reshape(m.shape)
            assert np.array_equal(got, m) and np.array_equal(parsed[i][1], m)
        cut = int(rng.integers(0, len(data)))
        (k2, _, _, _, _, p2), consumed2 = _native.ark_scan(data[:cut])
        whole = sum(1 for i in range(len(want)) if int(pay[i]) + want[i][1].nbytes <= cut)
        assert len(k2) == whole and consumed2 <= cut
        # an offset start: scanning from the second entry
        if len(want) > 1:
            start = int(pay[0]) + want[0][1].nbytes
            (k3, l3, _, _, _, _), c3 = _native.ark_scan(data, start)
            assert len(k3) == len(want) - 1 and c3 == len(data) and data[k3[0]:k3[0] + l3[0]].decode() == want[1][0]
    bad = b"bad key! \0BFM \4" + b"\0" * 20                             # a key the reference's regex rejects: scan stops
    assert len(_native.ark_scan(bad)[0][0]) == 0


def test_shell_pipes_are_waited_for_and_a_failing_child_fails_the_caller(tmp_path):
    # the clean-up thread of popen is non-daemon (reference kaldi_io.py:91-95) and wait_for_children /
    # ze_utils.wait_for_background_commands re-raise a non-zero exit on the caller's thread
    from xvector_b200 import ze_utils
    out = str(tmp_path / "late.ark")
    fd = kaldi_io.open_or_fd("| (sleep 0.4; cat > %s)" % out, "wb")          # a writer that is slow to flush
    kaldi_io.write_vec_flt(fd, np.arange(4, dtype=np.float32), key="v")
    fd.close()
    ze_utils.wait_for_background_commands()                                   # returns only after the child has exited
    assert dict(kaldi_io.read_vec_flt_ark(out))["v"].tolist() == [0.0, 1.0, 2.0, 3.0]
    fd = kaldi_io.open_or_fd("| (cat > /dev/null; exit 3)", "wb")
    fd.write(b"x")
    fd.close()
    with pytest.raises(kaldi_io.SubprocessFailed, match="returned 3"):
        kaldi_io.wait_for_children()
    kaldi_io.wait_for_children()                                              # reported once, then forgotten
    src = str(tmp_path / "in.ark")
    with open(src, "wb") as f:
        kaldi_io.write_vec_flt(f, np.ones(3, dtype=np.float32), key="a")
    assert [k for k, _ in kaldi_io.read_vec_flt_ark("cat %s |" % src)] == ["a"]
    kaldi_io.wait_for_children()
