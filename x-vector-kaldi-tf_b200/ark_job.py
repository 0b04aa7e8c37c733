"""Host side of an extraction job over a feature ark that lies in a regular file (include/xvec_job.h).

What the reference does per utterance in Python -- ``kaldi_io.read_mat_ark`` header parsing, the skip / chunk rules of
``Model.make_embedding`` (local/tf/models.py:373-409), ``kaldi_io.write_vec_flt`` (kaldi_io.py:309-343) -- is done here
per BATCH by the native reader / formatter; this module only orchestrates:

  * ``open_striped_reader``: one ``_native.ArkReader`` per rank over its byte stripe of the archive, and the exchange
    that confirms every stripe's first entry against the previous stripe's chain (a multi-GPU job reads the file ONCE
    in total, 1/N per rank, frames balanced to within one utterance -- the reference splits the data directory into
    ``nj`` pieces beforehand, extract_xvectors.sh:63-65);
  * ``VectorSink``: rank 0's output side -- formats blocks of x-vectors natively and writes them to the caller's stream
    (any ``write``-able) or ``kaldi_io.ArkScpWriter``.
"""
from __future__ import annotations

import os

import numpy as np

from . import sharding


def _all_gather_i64(values, device):
    """[world, len(values)] int64 numpy array of every rank's ``values`` (torch.distributed; gloo or nccl)."""
    import torch
    import torch.distributed as dist
    rank, world = sharding.dist_info()
    if world == 1:
        return np.asarray([values], dtype=np.int64)
    dev = torch.device(device) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(values), dtype=torch.int64, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return np.stack([t.cpu().numpy() for t in out])


def stripe_bounds(start, size, rank, world):
    """Byte stripe [begin, end) of rank ``rank``: equal shares of [start, size)."""
    span = max(size - start, 0)
    return start + span * rank // world, start + span * (rank + 1) // world


def open_striped_reader(path, start, feat_dim, min_chunk_size, chunk_size, batch_frames, device="cpu", pinned=True,
                        n_threads=None, n_slots=3, rank=None, world=None, feats_f16=False):
    """Index this rank's stripe of ``path`` (entries from byte ``start`` on) and agree the stripe boundaries with the
    other ranks.  Returns ``(reader, counts)`` with ``counts`` = [world, 6] int64 rows
    (n_entries, n_ok, n_fail, rows_used, stopped_at, key_bytes) of every rank -- or ``(None, counts)`` when some stripe holds an entry
    the native scanner cannot parse (text / compressed matrices): the caller then takes its general path."""
    from ._native import ArkReader
    if rank is None:
        rank, world = sharding.dist_info()
    size = os.path.getsize(path)
    begin, end = stripe_bounds(start, size, rank, world)
    if n_threads is None:
        n_threads = int(os.environ.get("XVEC_READER_THREADS", str(max(2, min(8, (os.cpu_count() or 4) // max(world, 1))))))
    reader = ArkReader(path, feat_dim, min_chunk_size, chunk_size, batch_frames, byte_begin=begin,
                       byte_end=(-1 if rank == world - 1 else end), begin_is_boundary=(rank == 0), n_threads=n_threads,
                       n_slots=n_slots, pinned=pinned, feats_f16=feats_f16)
    try:
        info = reader.index()
        if world > 1:
            # Every stripe r > 0 found its first entry by pattern search; stripe r-1's chain knows the truth (the marker and
            # key offsets of the first entry beyond it).  Apply, re-gather, repeat until nothing moves: one round when every
            # candidate was right, one more per empty stripe that has to hand its predecessor's boundary on.
            applied = None
            for _ in range(world + 1):
                table = _all_gather_i64([info["next_marker_off"], info["next_key_off"]], device)
                changed = 0
                if rank > 0:
                    want = (int(table[rank - 1][0]), int(table[rank - 1][1]))
                    if want != applied:
                        before = (info["next_marker_off"], info["next_key_off"])
                        info = reader.set_first(*want)
                        applied = want
                        changed = int((info["next_marker_off"], info["next_key_off"]) != before)
                if int(_all_gather_i64([changed], device).sum()) == 0:
                    break
            else:
                raise RuntimeError("the stripes of %s did not settle on common boundaries" % path)
        counts = _all_gather_i64([info["n_entries"], info["n_ok"], info["n_fail"], info["rows_used"], info["stopped_at"],
                                  info["key_bytes"]], device)
        if (counts[:, 4] >= 0).any():
            reader.close()
            return None, counts
        return reader, counts
    except BaseException:
        reader.close()
        raise


class VectorSink(object):
    """Rank 0's output side: blocks of float32 x-vectors + their keys -> the exact bytes ``write_vec_flt`` would emit,
    formatted natively, written with one ``write`` per block (plus the scp lines for an ``ArkScpWriter``)."""

    def __init__(self, output_stream):
        self.out = output_stream
        self.is_pair = hasattr(output_stream, "write_vec_block")

    def write(self, key_blob, key_off, vectors):
        """key_off: int64 [n + 1] offsets into ``key_blob`` (may be a window of a longer table)."""
        from ._native import vec_ark_format
        if len(vectors) == 0:
            return
        if self.is_pair:
            self.out.write_vec_block(key_blob, key_off, vectors)
        else:
            self.out.write(memoryview(vec_ark_format(key_blob, key_off, vectors)))


def gather_bytes_to_rank0(data, device):
    """Rank 0 gets the list of every rank's uint8 array ``data`` (None elsewhere): sizes first, then one padded gather."""
    import torch
    import torch.distributed as dist
    rank, world = sharding.dist_info()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    if world == 1:
        return [data]
    sizes = _all_gather_i64([data.shape[0]], device)[:, 0]
    dev = torch.device(device) if dist.get_backend() == "nccl" else torch.device("cpu")
    pad = int(max(int(sizes.max()), 1))
    mine = torch.zeros(pad, dtype=torch.uint8, device=dev)
    if data.shape[0]:
        mine[:data.shape[0]] = torch.from_numpy(data).to(dev)
    bufs = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, bufs, dst=0)
    if rank != 0:
        return None
    return [bufs[r][:int(sizes[r])].cpu().numpy() for r in range(world)]


def shared_output_spec(output_stream, total_bytes):
    """Rank 0: if the job's output lies in regular files that every rank of the node can open -- a
    ``kaldi_io.ArkScpWriter`` or a buffered writer over a named regular file -- the description the other ranks need to
    write their byte ranges themselves: ``dict(ark=path, base=offset where this job's first entry goes, scp_name=name the
    scp lines use for the ark or None)``.  The file is extended by ``total_bytes`` (the size of the job's entries) so that the
    ranks can map their ranges.  None for pipes, in-memory streams, gzip, anything else."""
    import io
    import stat
    try:
        if hasattr(output_stream, "write_vec_block") and hasattr(output_stream, "ark"):        # ArkScpWriter
            output_stream.ark.flush()
            output_stream.scp.flush()
            os.ftruncate(output_stream.ark.fileno(), int(output_stream.pos) + int(total_bytes))
            return dict(ark=os.path.abspath(output_stream.ark.name), base=int(output_stream.pos), scp_name=output_stream.name,
                        scp=os.path.abspath(output_stream.scp.name), scp_base=int(os.fstat(output_stream.scp.fileno()).st_size))
        if isinstance(output_stream, io.BufferedWriter) and isinstance(output_stream.name, str) and \
                stat.S_ISREG(os.fstat(output_stream.fileno()).st_mode):
            output_stream.flush()
            os.ftruncate(output_stream.fileno(), int(output_stream.tell()) + int(total_bytes))
            return dict(ark=os.path.abspath(output_stream.name), base=int(output_stream.tell()), scp_name=None)
    except (OSError, ValueError, AttributeError):
        return None
    return None


def populate_pages(mapped, offset, length):
    """Fault in the pages of ``mapped[offset : offset + length]`` (a writable mmap of the output file) from a background
    thread, so that the formatter's first touch of a page is not a page fault on the job's critical path
    (MADV_POPULATE_WRITE, Linux >= 5.14; silently nothing where it is not supported)."""
    import ctypes
    import mmap
    import threading
    if length <= 0:
        return None
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        base = ctypes.addressof(ctypes.c_char.from_buffer(mapped))
    except (OSError, TypeError, ValueError):
        return None
    page = mmap.PAGESIZE
    lo = (base + offset) // page * page
    hi = base + offset + length

    def run():
        step = 8 << 20
        for a in range(lo, hi, step):
            if libc.madvise(ctypes.c_void_p(a), ctypes.c_size_t(min(step, hi - a)), 23) != 0:      # MADV_POPULATE_WRITE
                return

    t = threading.Thread(target=run, name="xvec-populate", daemon=True)
    t.start()
    return t


_POW10 = np.array([10 ** k for k in range(1, 19)], dtype=np.int64)


def scp_line_offsets(key_off, ark_name, first_marker_base, entry_bytes):
    """Byte offset of every scp line ``key ark_name:offset\n`` of a run of entries (int64 [n + 1], starting at 0), computed
    WITHOUT formatting them: entry i starts ``key_off[i] + i * entry_bytes`` bytes behind ``first_marker_base`` and its
    offset is that of its binary marker, ``len(key) + 1`` further on.  This is what lets every rank of a job write its own
    byte range of the ONE scp (sizes are exchanged before a single line exists)."""
    key_off = np.asarray(key_off, dtype=np.int64)
    klen = np.diff(key_off)
    n = int(klen.shape[0])
    marker = first_marker_base + (key_off[:-1] - key_off[0]) + np.arange(n, dtype=np.int64) * entry_bytes + klen + 1
    digits = 1 + np.searchsorted(_POW10, marker, side="right")
    out = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(klen + len(os.fsencode(ark_name)) + 3 + digits, out=out[1:])
    return out


def extend_shared_scp(output_stream, spec, total_scp_bytes):
    """Rank 0: make room for the job's scp lines behind what the writer already holds (the ranks map their ranges of it)."""
    output_stream.scp.flush()
    os.ftruncate(output_stream.scp.fileno(), int(spec["scp_base"]) + int(total_scp_bytes))


def finish_shared_output(output_stream, spec, total_bytes):
    """Rank 0, after every rank has written its byte ranges (ark and scp): move the writer behind the job's entries."""
    end = spec["base"] + total_bytes
    if spec["scp_name"] is not None:
        output_stream.ark.seek(end)
        output_stream.pos = end
        output_stream.scp.seek(0, os.SEEK_END)
    else:
        output_stream.seek(end)


class _SynthBatch(object):
    __slots__ = ("slot", "n_seg", "n_utt", "n_rows", "feats", "seg_len", "utt_first_seg", "utt_dst_row", "first_ok_index", "ready_event")


class SyntheticSource(object):
    """A batch source with ``_native.ArkReader``'s interface whose utterances are GENERATED on the device
    (xv_synth_mfcc: counter-based, reproducible on the host with ``synthetic.counter_mfcc``): the workload of
    BASELINE configs[3] -- 1 M utterances of 200-1000 frames, 55 GB of features, more than a host can hold or feed
    (hard part 7).  Rank r owns the contiguous block of utterance ids [n * r / world, n * (r + 1) / world);
    keys are ``utt%07d``.  Lengths: ``synthetic.lengths_uniform(seed, n)``."""
    feats_on_device = True

    def __init__(self, device, n_utts, seed, batch_frames, feat_dim=23, rank=None, world=None, slots=3):
        import torch
        from . import synthetic
        if rank is None:
            rank, world = sharding.dist_info()
        self.device, self.seed, self.feat_dim = int(device), int(seed), int(feat_dim)
        lo, hi = n_utts * rank // world, n_utts * (rank + 1) // world
        self.ids = np.arange(lo, hi, dtype=np.int64)
        self.lens = synthetic.lengths_uniform(seed, n_utts)[lo:hi].astype(np.int32)
        n = len(self.ids)
        # batches: greedy runs of utterances of at most batch_frames rows
        ends = np.cumsum(self.lens, dtype=np.int64)
        self.bounds = [0]
        while self.bounds[-1] < n:
            start_rows = ends[self.bounds[-1] - 1] if self.bounds[-1] else 0
            nxt = int(np.searchsorted(ends, start_rows + batch_frames, side="right"))
            self.bounds.append(max(nxt, self.bounds[-1] + 1))
        self.bounds[-1] = n
        cap = max(int(batch_frames), int(self.lens.max()) if n else 1)
        self.bufs = [torch.empty((cap, feat_dim), dtype=torch.float32, device="cuda:%d" % device) for _ in range(min(slots, max(len(self.bounds) - 1, 1)))]
        self.stream = torch.cuda.Stream(device)
        self.next_batch, self.released, self.base = 0, 0, 0
        self.info = dict(n_entries=n, n_ok=n, n_fail=0, rows_used=int(self.lens.sum()), stopped_at=-1, key_bytes=10 * n)

    def counts(self, device_name):
        i = self.info
        return _all_gather_i64([i["n_entries"], i["n_ok"], i["n_fail"], i["rows_used"], i["stopped_at"], i["key_bytes"]], device_name)

    def failures(self):
        return []

    def keys(self):
        n = len(self.ids)
        blob = np.empty((n, 10), np.uint8)
        blob[:, :3] = np.frombuffer(b"utt", np.uint8)
        v = self.ids.copy()
        for k in range(7):
            blob[:, 9 - k] = 48 + (v % 10)
            v //= 10
        return blob.reshape(-1), np.arange(n + 1, dtype=np.int64) * 10

    def start(self, dst_row_base=0):
        self.base = int(dst_row_base)

    def next(self):
        from ._native import synth_mfcc
        if self.next_batch >= len(self.bounds) - 1:
            return None
        assert self.next_batch - self.released < len(self.bufs), "release a batch first"
        u0, u1 = self.bounds[self.next_batch], self.bounds[self.next_batch + 1]
        b = _SynthBatch()
        b.slot = self.next_batch % len(self.bufs)
        b.n_utt = b.n_seg = u1 - u0
        b.seg_len = self.lens[u0:u1]
        b.n_rows = int(b.seg_len.sum())
        b.utt_first_seg = np.arange(b.n_utt + 1, dtype=np.int32)
        b.utt_dst_row = self.base + np.arange(u0, u1, dtype=np.int64)
        b.first_ok_index = u0
        b.feats = self.bufs[b.slot][:b.n_rows]
        import torch
        synth_mfcc(self.device, self.bufs[b.slot], self.ids[u0:u1], b.seg_len, self.seed, stream=self.stream)
        b.ready_event = torch.cuda.Event()            # the submission runs on the engine's own stream: it waits for this on the device
        b.ready_event.record(self.stream)
        self.next_batch += 1
        return b

    def release(self, slot):
        self.released += 1

    def close(self):
        self.bufs = []
