"""Kaldi ark/scp I/O for the x-vector extractor.

Drop-in for the subset of the reference's ``local/tf/kaldi_io.py`` that the extraction hot
path touches plus the scp / compressed-matrix readers its
callers rely on.  Same function names, argument meaning and exception types; the byte
layouts are pinned by fixtures generated with the reference module itself
(tests/golden/make_golden_ark.py).

Wire formats (little endian, unaligned, entries concatenated):

    matrix : <key> ' ' '\\0B' 'FM ' '\\4' <i32 rows> '\\4' <i32 cols> rows*cols f32 row-major
             ('DM ' = f64, 'CM ' = Kaldi compressed matrix, ' [' starts the text form)
             -- reference kaldi_io.py:395-437 (read), :506-542 (write)
    vector : <key> ' ' '\\0B' 'FV ' '\\4' <u32 dim> dim f32      ('DV ' = f64)
             -- reference kaldi_io.py:266-305 (read), :309-343 (write)
    key    : bytes up to the first space, ``^[./a-zA-Z0-9_-]+$``; an empty key means EOF
             -- reference kaldi_io.py:120-133

Unlike the reference's ``read_key`` (one ``fd.read(1)`` per character) the readers here pull
whole headers with a single ``read`` where the stream allows ``peek``; the bytes consumed
from the stream are identical, so mixing calls with other readers stays safe.
"""
from __future__ import annotations

import gzip
import io
import mmap
import os
import re
import stat
import struct
import subprocess
import sys
import threading

import numpy as np

_KEY_RE = re.compile(r"^[./a-zA-Z0-9_-]+$")
_SPECIFIER_RE = re.compile(r"^(ark|scp)(,scp|,b|,t|,n?f|,n?p|,b?o|,n?s|,n?cs)*:")
_OFFSET_RE = re.compile(r":[0-9]+$")


class UnsupportedDataType(Exception):
    pass


class UnknownVectorHeader(Exception):
    pass


class UnknownMatrixHeader(Exception):
    pass


class BadSampleSize(Exception):
    pass


class BadInputFormat(Exception):
    pass


class SubprocessFailed(Exception):
    pass


# --------------------------------------------------------------------------------------
# opening streams

_children = []                 # one record per shell pipe opened by popen: {"cmd", "proc", "thread", "ret"}
_children_lock = threading.Lock()


def popen(cmd, mode="rb"):
    """Run ``cmd`` through the shell and return the pipe end matching ``mode``.

    As in the reference (kaldi_io.py:84-117) a NON-daemon clean-up thread waits for the child and raises SubprocessFailed
    when it exits non-zero, so the interpreter does not end while e.g. ``| copy-vector ark:- ark,scp:...`` is still
    flushing.  An exception in a thread cannot fail the job, so the exit status is also recorded and
    ``wait_for_children`` (called by ze_utils.wait_for_background_commands and before extract_embedding.py publishes its
    outputs) re-raises it on the caller's thread.
    """
    if not isinstance(cmd, str):
        raise TypeError("invalid cmd type (%s, expected string)" % type(cmd))
    if mode not in ("r", "w", "rb", "wb"):
        raise ValueError("invalid mode %s" % mode)
    reading = mode[0] == "r"
    proc = subprocess.Popen(cmd, shell=True,
                            stdout=subprocess.PIPE if reading else None,
                            stdin=None if reading else subprocess.PIPE)
    rec = dict(cmd=cmd, proc=proc, ret=None, thread=None)

    def watch():
        rec["ret"] = proc.wait()
        if rec["ret"] > 0:
            raise SubprocessFailed("cmd %s returned %d !" % (cmd, rec["ret"]))

    rec["thread"] = threading.Thread(target=watch, name="kaldi_io-popen-cleanup")
    rec["thread"].start()
    with _children_lock:
        _children.append(rec)
    pipe = proc.stdout if reading else proc.stdin
    return pipe if mode.endswith("b") else io.TextIOWrapper(pipe)


def wait_for_children(timeout=None):
    """Wait for every child started by ``popen`` to exit (their pipes must have been closed / drained by now) and raise
    SubprocessFailed for the first one that returned non-zero.  Children still running after ``timeout`` seconds (None =
    wait for ever, as the reference's thread join does) are left alone and stay registered."""
    with _children_lock:
        pending = list(_children)
    failed = None
    for rec in pending:
        rec["thread"].join(timeout)
        if rec["thread"].is_alive():
            continue
        with _children_lock:
            if rec in _children:
                _children.remove(rec)
        if rec["ret"] is not None and rec["ret"] > 0 and failed is None:      # > 0 as the reference: a SIGPIPE'd reader is fine
            failed = rec
    if failed is not None:
        raise SubprocessFailed("cmd %s returned %d !" % (failed["cmd"], failed["ret"]))


def open_or_fd(file, mode="rb"):
    """Open a Kaldi r/w-filename: plain or gzipped file, ``cmd |`` / ``| cmd`` pipe, optional
    ``ark:``/``scp:`` prefix and ``:offset`` suffix; anything that is not a string is taken
    to be an already opened stream and returned as is (reference kaldi_io.py:50-80)."""
    if not isinstance(file, str):
        return file
    offset = None
    if _SPECIFIER_RE.search(file):
        file = file.split(":", 1)[1]
    if _OFFSET_RE.search(file):
        file, offset = file.rsplit(":", 1)
    if file[-1] == "|":
        fd = popen(file[:-1], "rb")
    elif file[0] == "|":
        fd = popen(file[1:], "wb")
    elif file.split(".")[-1] == "gz":
        fd = gzip.open(file, mode)
    else:
        fd = open(file, mode)
    if offset is not None:
        fd.seek(int(offset))
    return fd


def _read_exact(fd, n):
    """``fd.read(n)`` that tolerates short reads from pipes."""
    buf = fd.read(n)
    if buf is None:
        buf = b""
    if len(buf) == n or not buf:
        return buf
    parts = [buf]
    got = len(buf)
    while got < n:
        more = fd.read(n - got)
        if not more:
            break
        parts.append(more)
        got += len(more)
    return b"".join(parts)


def read_key(fd):
    """Next utterance key from an ark stream, or None at end of stream."""
    chars = []
    peek = getattr(fd, "peek", None)
    while True:
        if peek is not None:
            window = peek(64)[:64]
            if not window:
                break
            cut = window.find(b" ")
            if cut < 0:
                chars.append(fd.read(len(window)))
                continue
            chars.append(fd.read(cut + 1)[:-1])
            break
        c = fd.read(1)
        if c == b"" or c == b" ":
            break
        chars.append(c)
    key = b"".join(chars).decode().strip()
    if key == "":
        return None
    assert _KEY_RE.match(key) is not None, "bad key %r" % key
    return key


def _check_binary_writable(fd):
    """The reference asserts ``fd.mode == 'wb'`` (kaldi_io.py:326,524); kept for text-mode
    mistakes, relaxed for streams whose ``mode`` is not a string (gzip) or absent (BytesIO)."""
    mode = getattr(fd, "mode", "wb")
    if isinstance(mode, str):
        assert mode == "wb", "open the output stream in binary mode ('wb'), got %r" % mode


# --------------------------------------------------------------------------------------
# integer vectors (alignments, utt2int-style tables), posteriors, confusion-network times, segments
# (reference kaldi_io.py:139-222, :553-700).  None of them is on the extraction or training path; they complete the
# module's surface for recipes that import it.

_INT_RECORD = np.dtype([("size", "i1"), ("value", "<i4")])                                   # '\4' <int32>
_POST_RECORD = np.dtype([("size_idx", "i1"), ("idx", "<i4"), ("size_post", "i1"), ("post", "<f4")])
_TIME_RECORD = np.dtype([("size_beg", "i1"), ("t_beg", "<f4"), ("size_end", "i1"), ("t_end", "<f4")])


def _read_sized_int(fd):
    """One '\4' <int32> token of a Kaldi binary stream."""
    tag = _read_exact(fd, 1)
    assert tag == b"\4", "expected a 4-byte integer, size tag was %r" % tag
    return struct.unpack("<i", _read_exact(fd, 4))[0]


def _read_records(fd, count, dtype, size_fields):
    rec = np.frombuffer(_read_exact(fd, count * dtype.itemsize), dtype=dtype, count=count)
    for name in size_fields:
        assert count == 0 or rec[0][name] == 4
    return rec


def read_vec_int(file_or_fd):
    """One Kaldi integer vector, binary or text (values as int32 / int)."""
    fd = open_or_fd(file_or_fd)
    try:
        flag = _read_exact(fd, 2).decode()
        if flag == "\0B":
            n = _read_sized_int(fd)
            return _read_records(fd, n, _INT_RECORD, ("size",))["value"]
        tokens = [t for t in (flag + fd.readline().decode()).split() if t not in ("[", "]")]
        return np.array(tokens, dtype=int)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vec_int_ark(file_or_fd):
    """Generator of (key, int vector) over an ark."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            yield key, read_vec_int(fd)
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_ali_ark(file_or_fd):
    """Alignments are integer vectors."""
    return read_vec_int_ark(file_or_fd)


def write_vec_int(file_or_fd, v, key=""):
    """Append one binary integer vector: [key ' '] '\0B' '\4' <dim> then '\4' <value> per element."""
    fd = open_or_fd(file_or_fd, mode="wb")
    _check_binary_writable(fd)
    try:
        v = np.asarray(v)
        body = np.empty(v.shape[0], dtype=_INT_RECORD)
        body["size"] = 4
        body["value"] = v
        head = (key + " ").encode() if key != "" else b""
        fd.write(head + b"\0B\4" + struct.pack("<i", v.shape[0]) + body.tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_post(file_or_fd):
    """One binary Kaldi Posterior (vector<vector<pair<int32, float>>>) as a list, per frame, of (index, value) tuples."""
    fd = open_or_fd(file_or_fd)
    try:
        assert _read_exact(fd, 2) == b"\0B"
        frames = []
        for _ in range(_read_sized_int(fd)):
            rec = _read_records(fd, _read_sized_int(fd), _POST_RECORD, ("size_idx", "size_post"))
            frames.append(rec[["idx", "post"]].tolist())
        return frames
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_post_ark(file_or_fd):
    """Generator of (key, posterior) over an ark."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            yield key, read_post(fd)
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_cnet_ark(file_or_fd):
    """Confusion networks are stored as posteriors."""
    return read_post_ark(file_or_fd)


def read_cntime(file_or_fd):
    """Begin / end times of the bins of one confusion network: list of (t_beg, t_end)."""
    fd = open_or_fd(file_or_fd)
    try:
        assert _read_exact(fd, 2) == b"\0B"
        rec = _read_records(fd, _read_sized_int(fd), _TIME_RECORD, ("size_beg", "size_end"))
        return rec[["t_beg", "t_end"]].tolist()
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_cntime_ark(file_or_fd):
    """Generator of (key, bin times) over an ark."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            yield key, read_cntime(fd)
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_segments_as_bool_vec(segments_file):
    """A Kaldi ``segments`` file of ONE recording ('<utt> <rec> <t-beg> <t-end>', seconds) as a per-frame bool vector at
    100 frames per second: True inside a segment, ending at the last segment's end."""
    segs = np.loadtxt(segments_file, dtype="object,object,f,f", ndmin=1)
    assert len(segs) > 0, "empty segmentation"
    assert len({rec[1] for rec in segs}) == 1, "segments of more than one recording"
    start = np.rint([100 * rec[2] for rec in segs]).astype(int)
    end = np.rint([100 * rec[3] for rec in segs]).astype(int)
    frames = np.zeros(int(end[-1]) if len(end) else 0, dtype=bool)
    for b, e in zip(start, end):
        frames[b:e] = True
    assert frames.sum() == np.sum(end - start)
    return frames


# --------------------------------------------------------------------------------------
# float vectors

def _read_vec_flt_binary(fd):
    header = _read_exact(fd, 3).decode()
    if header == "FV ":
        dtype = np.dtype("<f4")
    elif header == "DV ":
        dtype = np.dtype("<f8")
    else:
        raise UnknownVectorHeader("The header contained '%s'" % header)
    assert _read_exact(fd, 1) == b"\4"
    dim = struct.unpack("<i", _read_exact(fd, 4))[0]
    return np.frombuffer(_read_exact(fd, dim * dtype.itemsize), dtype=dtype)


def read_vec_flt(file_or_fd):
    """One Kaldi float vector, binary or text."""
    fd = open_or_fd(file_or_fd)
    try:
        flag = _read_exact(fd, 2).decode()
        if flag == "\0B":
            return _read_vec_flt_binary(fd)
        tokens = (flag + fd.readline().decode()).strip().split()
        tokens = [t for t in tokens if t not in ("[", "]")]
        return np.array(tokens, dtype=float)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vec_flt_ark(file_or_fd):
    """Generator of (key, vector) over a vector ark."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            yield key, read_vec_flt(fd)
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vec_flt_scp(file_or_fd):
    """Generator of (key, vector) following an scp of ``key rxfilename[:offset]`` lines."""
    fd = open_or_fd(file_or_fd)
    try:
        for line in fd:
            key, rxfile = line.decode().split(" ")
            yield key, read_vec_flt(rxfile.strip())
    finally:
        if fd is not file_or_fd:
            fd.close()


def write_vec_flt(file_or_fd, v, key=""):
    """Append one binary vector (float32 -> 'FV ', float64 -> 'DV ')."""
    fd = open_or_fd(file_or_fd, mode="wb")
    _check_binary_writable(fd)
    try:
        if v.dtype == "float32":
            tag = b"FV "
        elif v.dtype == "float64":
            tag = b"DV "
        else:
            raise UnsupportedDataType("'%s', please use 'float32' or 'float64'" % v.dtype)
        head = (key + " ").encode() if key != "" else b""
        fd.write(head + b"\0B" + tag + b"\4" + struct.pack("<I", v.shape[0]))
        fd.write(v.tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


def vec_flt_entry_bytes(v, key):
    """The exact bytes ``write_vec_flt`` emits for one float32 entry (used for batched writes)."""
    return (key + " ").encode() + b"\0BFV \4" + struct.pack("<I", v.shape[0]) + np.ascontiguousarray(v, "<f4").tobytes()


# --------------------------------------------------------------------------------------
# float matrices

def _read_compressed_mat(fd, fmt):
    """Kaldi CompressedMatrix, format 'CM ' only (reference kaldi_io.py:455-502):
    16-byte global header {min f32, range f32, rows i32, cols i32}, per column four u16
    percentiles, then column-major u8 payload; piecewise-linear decode."""
    assert fmt == "CM "
    gmin, grange, rows, cols = struct.unpack("<ffii", _read_exact(fd, 16))
    gmin, grange = np.float32(gmin), np.float32(grange)
    pct = np.frombuffer(_read_exact(fd, cols * 8), dtype="<u2").reshape(cols, 4)
    data = np.frombuffer(_read_exact(fd, cols * rows), dtype=np.uint8).reshape(cols, rows)
    # percentile u16 -> float (kaldi: min + range * 1/65535 * value), float32 like the reference
    q = (gmin + grange * np.float32(1.52590218966964e-05) * pct.astype(np.float32)).astype(np.float32)
    p0, p25, p75, p100 = (q[:, i:i + 1].astype(np.float32) for i in range(4))
    d = data.astype(np.float32)
    out = np.where(data <= 64, p0 + (p25 - p0) / np.float32(64.0) * d,
                   np.where(data <= 192, p25 + (p75 - p25) / np.float32(128.0) * (d - np.float32(64.0)),
                            p75 + (p100 - p75) / np.float32(63.0) * (d - np.float32(192.0))))
    return out.astype(np.float32).T


def _read_mat_binary(fd):
    header = _read_exact(fd, 3).decode()
    if header.startswith("CM"):
        return _read_compressed_mat(fd, header)
    if header == "FM ":
        dtype = np.dtype("<f4")
    elif header == "DM ":
        dtype = np.dtype("<f8")
    else:
        raise UnknownMatrixHeader("The header contained '%s'" % header)
    _, rows, _, cols = struct.unpack("<bibi", _read_exact(fd, 10))
    buf = _read_exact(fd, rows * cols * dtype.itemsize)
    return np.frombuffer(buf, dtype=dtype).reshape(rows, cols)


def _read_mat_ascii(fd):
    rows = []
    while True:
        line = fd.readline().decode()
        if len(line) == 0:
            raise BadInputFormat
        tokens = line.strip().split()
        if not tokens:
            continue
        if tokens[-1] != "]":
            rows.append(np.array(tokens, dtype="float32"))
        else:
            rows.append(np.array(tokens[:-1], dtype="float32"))
            return np.vstack(rows)


def read_mat(file_or_fd):
    """One Kaldi matrix, binary ('FM ', 'DM ', 'CM ') or text."""
    fd = open_or_fd(file_or_fd)
    try:
        flag = _read_exact(fd, 2).decode()
        if flag == "\0B":
            return _read_mat_binary(fd)
        assert flag == " ["
        return _read_mat_ascii(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_mat_ark(file_or_fd):
    """Generator of (key, matrix) over a matrix ark: the extractor's input loop
    (reference models.py:373 -> kaldi_io.py:372-392)."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            yield key, read_mat(fd)
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


_last_plain = (None, False)


def _plain_file(fd):
    """True for a buffered reader over a regular file: payloads can then be fetched with pread, by any thread,
    without touching the stream's position.  (The last answer is remembered: the reader asks once per matrix.)"""
    global _last_plain
    if fd is _last_plain[0]:
        return _last_plain[1]
    try:
        plain = isinstance(fd, io.BufferedReader) and stat.S_ISREG(os.fstat(fd.fileno()).st_mode)
    except (OSError, ValueError, AttributeError):
        plain = False
    _last_plain = (fd, plain)
    return plain


class PayloadGroup(object):
    """A few megabytes of matrix payloads to fetch with pread in ONE job of the extractor's reader pool (a job per matrix
    costs more in hand-over than the 50 KB read it does).  Every distinct file is duplicated once, while it is certainly
    still open, so that the scp reader may move on and close it; ``fetch`` closes the duplicates."""
    __slots__ = ("items", "nbytes", "_dups")

    def __init__(self):
        self.items, self.nbytes, self._dups = [], 0, {}

    def add(self, fileobj, offset, dst):
        key = id(fileobj)
        if key not in self._dups:
            self._dups[key] = (fileobj, os.dup(fileobj.fileno()))      # the reference to fileobj keeps id() unique
        self.items.append((self._dups[key][1], offset, dst))
        self.nbytes += dst.nbytes

    def fetch(self):
        """Runs on a pool thread: read system calls release the GIL, so groups are read in parallel."""
        try:
            for fd, offset, dst in self.items:
                view = memoryview(dst).cast("B")
                got, need = 0, len(view)
                while got < need:
                    n = os.preadv(fd, [view[got:]], offset + got)
                    if n <= 0:
                        raise BadInputFormat("truncated matrix payload at byte %d" % (offset + got))
                    got += n
        finally:
            for _, fd in self._dups.values():
                os.close(fd)
            self._dups = {}


class MatArkEntry(object):
    """One entry of ``read_mat_ark_entries``: header known, payload not consumed yet.  Exactly one of
    ``read_into`` / ``read`` / ``skip`` must be called before the generator is advanced."""
    __slots__ = ("key", "rows", "cols", "_fd", "_kind", "_mat", "_done")

    def __init__(self, key, rows, cols, fd, kind, mat=None):
        self.key, self.rows, self.cols, self._fd, self._kind, self._mat, self._done = key, rows, cols, fd, kind, mat, False

    def read(self):
        """The matrix as read_mat would return it."""
        assert not self._done
        self._done = True
        if self._kind == "FM":
            buf = _read_exact(self._fd, self.rows * self.cols * 4)
            return np.frombuffer(buf, dtype="<f4").reshape(self.rows, self.cols)
        if self._kind == "DM":
            buf = _read_exact(self._fd, self.rows * self.cols * 8)
            return np.frombuffer(buf, dtype="<f8").reshape(self.rows, self.cols)
        return self._mat                                    # already decoded (compressed / text)

    def read_into(self, dst):
        """Payload straight into ``dst`` (C-contiguous float32 [rows, cols], e.g. a slice of a page-locked
        staging buffer): for binary float32 matrices no intermediate copy is made."""
        assert dst.shape == (self.rows, self.cols) and dst.dtype == np.float32 and dst.flags.c_contiguous
        if self._kind == "FM" and self.rows * self.cols > 0:
            assert not self._done
            self._done = True
            view = memoryview(dst).cast("B")
            got, need = 0, len(view)
            readinto = getattr(self._fd, "readinto", None)
            while got < need:
                if readinto is not None:
                    n = readinto(view[got:])
                else:
                    chunk = self._fd.read(need - got)
                    n = len(chunk)
                    view[got:got + n] = chunk
                if not n:
                    raise BadInputFormat("truncated matrix for key %r" % self.key)
                got += n
        else:
            dst[...] = self.read()

    def detach_payload(self):
        """For a binary float32 matrix stored in a regular file: move the stream past the payload WITHOUT reading it and
        return ``(file object, offset)`` for ``PayloadGroup.add``.  ``None`` when the payload has to come through the
        stream (pipes, gzip, in-memory streams, other matrix kinds)."""
        nbytes = self.rows * self.cols * 4
        if self._kind != "FM" or self._done or nbytes == 0 or not _plain_file(self._fd):
            return None
        offset = self._fd.tell()
        self._fd.seek(nbytes, 1)
        self._done = True
        return self._fd, offset

    def skip(self):
        if self._done:
            return
        if self._kind in ("FM", "DM") and _plain_file(self._fd):
            self._done = True
            self._fd.seek(self.rows * self.cols * (4 if self._kind == "FM" else 8), 1)
        else:
            self.read()


class IndexedMatEntry(object):
    """Entry of ``read_mat_ark_entries_indexed``: same interface as MatArkEntry, but the header came from a native scan
    of the mmap'ed file and the payload is fetched with pread (nothing goes through the stream)."""
    __slots__ = ("key", "rows", "cols", "_fd", "_offset", "_elem")

    def __init__(self, key, rows, cols, fd, offset, elem_bytes):
        self.key, self.rows, self.cols, self._fd, self._offset, self._elem = key, rows, cols, fd, offset, elem_bytes

    def _pread(self, view):
        got, need, fileno = 0, len(view), self._fd.fileno()
        while got < need:
            n = os.preadv(fileno, [view[got:]], self._offset + got)
            if n <= 0:
                raise BadInputFormat("truncated matrix for key %r" % self.key)
            got += n

    def read(self):
        out = np.empty((self.rows, self.cols), dtype="<f4" if self._elem == 4 else "<f8")
        if out.size:
            self._pread(memoryview(out).cast("B"))
        return out

    def read_into(self, dst):
        assert dst.shape == (self.rows, self.cols) and dst.dtype == np.float32 and dst.flags.c_contiguous
        if self._elem == 4 and dst.size:
            self._pread(memoryview(dst).cast("B"))
        else:
            dst[...] = self.read()

    def detach_payload(self):
        if self._elem != 4 or self.rows * self.cols == 0:
            return None
        return self._fd, self._offset

    def skip(self):
        pass


def read_mat_ark_entries_indexed(fd, scan, chunk_entries=1 << 16):
    """``read_mat_ark_entries`` for an ark in a regular file (``fd``: buffered reader at an entry boundary): headers are
    indexed ``chunk_entries`` at a time by ``scan(buffer, start, max_entries)`` (the native xv_ark_scan through
    _native.ark_scan) over the mmap'ed file instead of being parsed one by one; entries the scanner does not know (text,
    compressed) and everything after them go through the general parser, which also owns the error messages."""
    pos = fd.tell()
    size = os.fstat(fd.fileno()).st_size
    if size > pos:
        with mmap.mmap(fd.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            while True:
                (key_off, key_len, rows, cols, elem, pay), consumed = scan(mm, pos, chunk_entries)
                keys = [mm[o:o + n].decode() for o, n in zip(key_off.tolist(), key_len.tolist())]
                for key, r, c, e, p in zip(keys, rows.tolist(), cols.tolist(), elem.tolist(), pay.tolist()):
                    yield IndexedMatEntry(key, r, c, fd, p, e)
                pos = consumed
                if len(keys) < chunk_entries:
                    break
    fd.seek(pos)
    for entry in read_mat_ark_entries(fd):
        yield entry


def read_mat_ark_entries(file_or_fd):
    """Generator of MatArkEntry over a matrix ark: like read_mat_ark (reference kaldi_io.py:372-392), but the
    shape is known before the payload is read, so the extractor can skip short utterances (models.py:378-387)
    cheaply and place the rows directly in its staging buffer."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            flag = _read_exact(fd, 2).decode()
            if flag == "\0B":
                header = _read_exact(fd, 3).decode()
                if header in ("FM ", "DM "):
                    _, rows, _, cols = struct.unpack("<bibi", _read_exact(fd, 10))
                    entry = MatArkEntry(key, rows, cols, fd, header[:2])
                elif header.startswith("CM"):
                    m = _read_compressed_mat(fd, header)
                    entry = MatArkEntry(key, m.shape[0], m.shape[1], fd, "decoded", m)
                else:
                    raise UnknownMatrixHeader("The header contained '%s'" % header)
            else:
                assert flag == " ["
                m = _read_mat_ascii(fd)
                entry = MatArkEntry(key, m.shape[0], m.shape[1] if m.ndim == 2 else 0, fd, "decoded", m)
            yield entry
            entry.skip()
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def _entry_at(fd, key):
    """MatArkEntry for the matrix whose '\0B' flag starts at the current position of ``fd``."""
    flag = _read_exact(fd, 2).decode()
    if flag == "\0B":
        header = _read_exact(fd, 3).decode()
        if header in ("FM ", "DM "):
            _, rows, _, cols = struct.unpack("<bibi", _read_exact(fd, 10))
            return MatArkEntry(key, rows, cols, fd, header[:2])
        if header.startswith("CM"):
            m = _read_compressed_mat(fd, header)
            return MatArkEntry(key, m.shape[0], m.shape[1], fd, "decoded", m)
        raise UnknownMatrixHeader("The header contained '%s'" % header)
    assert flag == " ["
    m = _read_mat_ascii(fd)
    return MatArkEntry(key, m.shape[0], m.shape[1] if m.ndim == 2 else 0, fd, "decoded", m)


class _RxFileCache(object):
    """Keeps the most recently used ark of an scp open: consecutive scp lines nearly always point into one file."""

    def __init__(self):
        self.path, self.fd = None, None

    def seek(self, rxfile):
        rxfile = rxfile.strip()
        if _OFFSET_RE.search(rxfile) and not rxfile.endswith("|"):
            path, offset = rxfile.rsplit(":", 1)
            if path != self.path:
                self.close()
                self.path, self.fd = path, open_or_fd(path)
            self.fd.seek(int(offset))
            return self.fd
        self.close()
        self.path, self.fd = rxfile, open_or_fd(rxfile)
        return self.fd

    def close(self):
        if self.fd is not None:
            self.fd.close()
        self.path, self.fd = None, None


class _MappedFiles(object):
    """The regular files an scp table points into, opened and mmap'ed once each and kept until ``close``: headers are
    then parsed from memory (no seek / read system calls per entry) and payloads fetched with pread."""

    def __init__(self):
        self._open = {}

    def get(self, path):
        """(file object, mmap) or (None, None) when ``path`` is not a plain, non-empty, uncompressed file."""
        hit = self._open.get(path)
        if hit is None:
            hit = (None, None)
            if not path.endswith(".gz") and os.path.isfile(path):
                f = open(path, "rb")
                if _plain_file(f) and os.fstat(f.fileno()).st_size > 0:
                    hit = (f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ))
                else:
                    f.close()
            self._open[path] = hit
        return hit

    def close(self):
        for f, mm in self._open.values():
            if mm is not None:
                mm.close()
                f.close()
        self._open = {}


def _split_rxfile(rxfile):
    """('path', offset) of an scp target of the form ``path:offset``; (None, None) for pipes and offset-less names."""
    rxfile = rxfile.strip()
    if rxfile.endswith("|") or not _OFFSET_RE.search(rxfile):
        return None, None
    path, offset = rxfile.rsplit(":", 1)
    return path, int(offset)


def _mapped_mat_entry(mm, fileobj, key, offset):
    """IndexedMatEntry for the binary float matrix whose '\0B' flag is at ``offset`` of the mapped file, or None when
    something else is stored there (text, compressed: the stream parser takes over)."""
    hdr = mm[offset:offset + 15]
    if len(hdr) < 15 or hdr[0:2] != b"\0B" or hdr[2:5] not in (b"FM ", b"DM ") or hdr[5] != 4 or hdr[10] != 4:
        return None
    rows, cols = struct.unpack_from("<i", hdr, 6)[0], struct.unpack_from("<i", hdr, 11)[0]
    elem = 4 if hdr[2:3] == b"F" else 8
    if rows < 0 or cols < 0 or offset + 15 + rows * cols * elem > len(mm):
        raise BadInputFormat("truncated matrix for key %r" % key)
    return IndexedMatEntry(key, rows, cols, fileobj, offset + 15, elem)


def read_mat_scp_entries(file_or_fd):
    """read_mat_ark_entries over an scp table (``key rxfilename[:offset]`` lines, reference kaldi_io.py:350-369): what
    ``scp:data/feats.scp`` means as a feature rspecifier when the feature front end runs on the device."""
    fd = open_or_fd(file_or_fd)
    cache = _RxFileCache()
    files = _MappedFiles()
    try:
        for line in fd:
            line = line.decode() if isinstance(line, bytes) else line
            if not line.strip():
                continue
            key, rxfile = line.rstrip("\n").split(" ", 1)
            path, offset = _split_rxfile(rxfile)
            entry = None
            if path is not None:
                fileobj, mm = files.get(path)
                if mm is not None:
                    entry = _mapped_mat_entry(mm, fileobj, key, offset)
            if entry is None:                            # pipes, gzip, text / compressed matrices
                entry = _entry_at(cache.seek(rxfile), key)
            yield entry
            entry.skip()
    finally:
        cache.close()
        files.close()
        if fd is not file_or_fd:
            fd.close()


class VecTable(object):
    """Float vectors by key: the ``scp,s,cs:vad.scp`` table select-voiced-frames opens
    (reference local/tf/extract_xvectors.sh:68).  An scp is loaded as a key -> rxfilename map and read on demand; an
    ark is walked once, front to back, assuming the caller asks in the archive's order (Kaldi's ``s,cs``)."""

    def __init__(self, rspecifier):
        self._cache = _RxFileCache()
        self._files = _MappedFiles()
        self._table = None
        self._iter = None
        self._pending = None
        spec = _SPECIFIER_RE.search(rspecifier) if isinstance(rspecifier, str) else None
        if spec is not None and spec.group(0).startswith("scp"):
            self._table = {}
            with open_or_fd(rspecifier) as fd:
                for line in fd:
                    line = line.decode() if isinstance(line, bytes) else line
                    if line.strip():
                        key, rxfile = line.rstrip("\n").split(" ", 1)
                        self._table[key] = rxfile
        else:
            self._iter = read_vec_flt_ark(rspecifier)

    def get(self, key):
        """The vector stored under ``key`` or None."""
        if self._table is not None:
            rxfile = self._table.get(key)
            if rxfile is None:
                return None
            path, offset = _split_rxfile(rxfile)
            if path is not None:
                _, mm = self._files.get(path)
                if mm is not None:
                    hdr = mm[offset:offset + 10]
                    if len(hdr) == 10 and hdr[0:2] == b"\0B" and hdr[2:5] in (b"FV ", b"DV ") and hdr[5] == 4:
                        dtype = np.dtype("<f4" if hdr[2:3] == b"F" else "<f8")
                        dim = struct.unpack_from("<i", hdr, 6)[0]
                        if dim < 0 or offset + 10 + dim * dtype.itemsize > len(mm):
                            raise BadInputFormat("truncated vector for key %r" % key)
                        return np.frombuffer(mm, dtype=dtype, count=dim, offset=offset + 10).copy()
            return read_vec_flt(self._cache.seek(rxfile))
        while True:
            if self._pending is None:
                self._pending = next(self._iter, None)
                if self._pending is None:
                    return None
            if self._pending[0] == key:
                vec, self._pending = self._pending[1], None
                return vec
            if self._pending[0] > key:               # sorted tables: the key is not there
                return None
            self._pending = None

    def close(self):
        self._cache.close()
        self._files.close()
        if self._iter is not None:
            self._iter.close()


def read_mat_scp(file_or_fd):
    """Generator of (key, matrix) following an scp file."""
    fd = open_or_fd(file_or_fd)
    try:
        for line in fd:
            key, rxfile = line.decode().split(" ")
            yield key, read_mat(rxfile.strip())
    finally:
        if fd is not file_or_fd:
            fd.close()


def write_mat(file_or_fd, m, key=""):
    """Append one binary matrix (float32 -> 'FM ', float64 -> 'DM ')."""
    fd = open_or_fd(file_or_fd, mode="wb")
    _check_binary_writable(fd)
    try:
        if m.dtype == "float32":
            tag = b"FM "
        elif m.dtype == "float64":
            tag = b"DM "
        else:
            raise UnsupportedDataType("'%s', please use 'float32' or 'float64'" % m.dtype)
        head = (key + " ").encode() if key != "" else b""
        fd.write(head + b"\0B" + tag + b"\4" + struct.pack("<I", m.shape[0]) + b"\4" + struct.pack("<I", m.shape[1]))
        fd.write(m.tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


# --------------------------------------------------------------------------------------
# native "ark,scp:" writer (what `| copy-vector ark:- ark,scp:A,S` does in the recipe,
# extract_xvectors.sh:76,86 -- without needing the Kaldi binary)

class ArkScpWriter(object):
    """Writes a float-vector ark and the scp that indexes it (``key ark_path:offset`` where
    offset is the byte position of the ``\\0B`` binary marker, as Kaldi's TableWriter emits)."""

    mode = "wb"

    def __init__(self, ark_path, scp_path, scp_ark_name=None):
        self.ark = open(ark_path, "wb")
        self.scp = open(scp_path, "wt")
        self.name = scp_ark_name if scp_ark_name is not None else ark_path
        self.pos = 0

    def write_vec_entries(self, keys, vectors):
        blobs, lines = [], []
        for key, v in zip(keys, vectors):
            blob = vec_flt_entry_bytes(v, key)
            lines.append("%s %s:%d\n" % (key, self.name, self.pos + len(key) + 1))
            self.pos += len(blob)
            blobs.append(blob)
        self.ark.write(b"".join(blobs))
        self.scp.write("".join(lines))

    def write_vec_block(self, key_blob, key_off, vectors):
        """``write_vec_entries`` for a block whose keys come as one uint8 blob + int64 offsets [n + 1]: ark bytes and scp
        lines are formatted natively (xv_vec_ark_format / xv_scp_format, include/xvec_job.h), same bytes."""
        from ._native import scp_format, vec_ark_format
        blob, markers = vec_ark_format(key_blob, key_off, vectors, with_markers=True)
        lines = scp_format(key_blob, key_off, self.name, self.pos, markers)
        self.pos += int(blob.shape[0])
        self.ark.write(memoryview(blob))
        self.scp.write(lines.tobytes().decode())

    def close(self):
        self.ark.close()
        self.scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open_vector_writer(wspecifier):
    """``ark,scp:A,S`` / ``scp,ark:S,A`` (no pipe) -> ArkScpWriter; anything else -> open_or_fd."""
    spec = wspecifier.strip()
    if spec.startswith("ark,scp:") and "|" not in spec:
        ark, scp = spec[8:].split(",")
        return ArkScpWriter(ark, scp)
    if spec.startswith("scp,ark:") and "|" not in spec:
        scp, ark = spec[8:].split(",")
        return ArkScpWriter(ark, scp)
    return open_or_fd(wspecifier, "wb")
