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
                        n_threads=None, n_slots=3, rank=None, world=None):
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
        n_threads = int(os.environ.get("XVEC_READER_THREADS", str(max(2, min(4, (os.cpu_count() or 4) // max(world, 1))))))
    reader = ArkReader(path, feat_dim, min_chunk_size, chunk_size, batch_frames, byte_begin=begin,
                       byte_end=(-1 if rank == world - 1 else end), begin_is_boundary=(rank == 0), n_threads=n_threads,
                       n_slots=n_slots, pinned=pinned)
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
            return dict(ark=os.path.abspath(output_stream.ark.name), base=int(output_stream.pos), scp_name=output_stream.name)
        if isinstance(output_stream, io.BufferedWriter) and isinstance(output_stream.name, str) and \
                stat.S_ISREG(os.fstat(output_stream.fileno()).st_mode):
            output_stream.flush()
            os.ftruncate(output_stream.fileno(), int(output_stream.tell()) + int(total_bytes))
            return dict(ark=os.path.abspath(output_stream.name), base=int(output_stream.tell()), scp_name=None)
    except (OSError, ValueError, AttributeError):
        return None
    return None


def finish_shared_output(output_stream, spec, total_bytes, scp_parts):
    """Rank 0, after every rank has written its byte range: move the writer behind the job's entries and append the scp
    lines the ranks produced (in rank order)."""
    end = spec["base"] + total_bytes
    if spec["scp_name"] is not None:
        output_stream.ark.seek(end)
        output_stream.pos = end
        for part in scp_parts:
            if len(part):
                output_stream.scp.write(part.tobytes().decode())
    else:
        output_stream.seek(end)
