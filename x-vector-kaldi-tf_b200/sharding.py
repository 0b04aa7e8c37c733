"""Utterance sharding across the GPUs of one box and the gather of embeddings to rank 0.

The reference scales extraction by running ``nj`` unrelated OS processes over a pre-split
data directory and concatenating their outputs (local/tf/extract_xvectors.sh:63-65,83-95);
there is no exchange step inside the forward pass.  Here: one process per GPU
(``torch.distributed``) and no data-path collective.  An archive in a regular file is cut into byte stripes (``ark_job.py``:
every rank reads 1/N of it, and either writes its own byte range of the one output ark or stores its rows into rank 0's
peer-memory table); what is left here serves streams, whose length is unknown ahead of time: utterances dealt block-cyclically
and -- for process groups without peer memory (gloo in the CPU tests) -- ONE gather of ``[n_r, emb_dim]`` fp32 + int64 indices
to rank 0, which writes the ark.
"""
from __future__ import annotations

import numpy as np


def dist_info():
    """(rank, world_size) of the initialised default process group, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def block_cyclic_rank(index, world_size, block=16):
    """Rank of the index-th utterance of a *stream* (lengths unknown ahead of time)."""
    return (index // block) % world_size


def gather_to_rank0(local_index, local_emb, total, emb_dim, device=None):
    """Gather (index, embedding) pairs of every rank to rank 0.

    local_index: int64 [n_r]; local_emb: float32 [n_r, emb_dim] (numpy).  Returns on rank 0 a
    float32 [total, emb_dim] array with every row filled (asserted), on other ranks None.
    Backend nccl -> tensors staged on ``device``; gloo -> CPU tensors.
    """
    import torch
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        out = np.empty((total, emb_dim), np.float32)
        out[np.asarray(local_index, dtype=np.int64)] = local_emb
        return out
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device(device if device is not None else "cuda") if use_cuda else torch.device("cpu")
    n_local = torch.tensor([len(local_index)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    n_max = max(max(counts), 1)
    idx = torch.full((n_max,), -1, dtype=torch.int64, device=dev)
    emb = torch.zeros((n_max, emb_dim), dtype=torch.float32, device=dev)
    if len(local_index):
        idx[:len(local_index)] = torch.as_tensor(np.asarray(local_index, dtype=np.int64)).to(dev)
        emb[:len(local_index)] = torch.as_tensor(np.ascontiguousarray(local_emb, dtype=np.float32)).to(dev)
    if rank == 0:
        idx_all = [torch.empty_like(idx) for _ in range(world)]
        emb_all = [torch.empty_like(emb) for _ in range(world)]
    else:
        idx_all = emb_all = None
    dist.gather(idx, idx_all, dst=0)
    dist.gather(emb, emb_all, dst=0)
    if rank != 0:
        return None
    out = np.empty((total, emb_dim), np.float32)
    seen = np.zeros(total, dtype=bool)
    for r in range(world):
        n = counts[r]
        if n == 0:
            continue
        i = idx_all[r][:n].cpu().numpy()
        out[i] = emb_all[r][:n].cpu().numpy()
        seen[i] = True
    assert seen.all(), "gather lost %d embeddings" % int((~seen).sum())
    return out
