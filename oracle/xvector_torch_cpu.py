"""Second, independent CPU restatement of the reference forward  --  TEST INFRASTRUCTURE ONLY.

Same contract as ``oracle/xvector_oracle.py`` (see its header: forward parity is UNPINNED,
TensorFlow 1.x is absent) but through a different code path: ``torch.nn.functional.conv1d``
(oneDNN) instead of explicit per-tap matmuls.  Two roles:

  1. cross-check of the numpy oracle (tests/test_oracle.py requires agreement <= 1e-6 in
     fp64 and reports the fp32 distance as the noise floor);
  2. the timed CPU baseline of ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``):
     fp32, run in the reference's operating mode -- one utterance per call
     (local/tf/models.py:410-414) -- since TensorFlow itself cannot be installed here.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .xvector_oracle import BN_EPSILON, TOPOLOGIES, VAR2STD_EPSILON, chunk_plan


class TorchCpuXvector:
    """Holds torch copies of the parameters; ``forward(x[T,D]) -> emb[512]``."""

    def __init__(self, params, topology="ModelWithoutDropout", dtype=torch.float32):
        topo = TOPOLOGIES[topology] if isinstance(topology, str) else topology
        self.dtype = dtype
        self.act = topo.get("act", "relu")       # relu | lrelu (0.2, models.py:912) | prelu (tf_block.py:38-47)
        self.layers = []
        for i, d in enumerate(topo["dilations"]):
            s = "frame_level_info_layer-%d/" % i
            w = torch.as_tensor(np.asarray(params[s + "w:0"])).to(dtype)       # [k, Cin, Cout]  (models.py:472)
            k = w.shape[0]
            g = lambda n: torch.as_tensor(np.asarray(params[s + n])).to(dtype)
            inv = g("gamma:0") * torch.rsqrt(g("variance:0") + BN_EPSILON)       # tf_block.py:26
            self.layers.append(dict(
                w=w.permute(2, 1, 0).contiguous(),                               # torch wants [Cout, Cin, k]
                b=g("b:0"), inv=inv, shift=g("beta:0") - g("mean:0") * inv,
                alpha=(g("prelu/prelu:0") if self.act == "prelu" else None),
                dilation=d, pad=((k - 1) * d) // 2))
        self.attention = None
        if topo.get("pooling") == "attention":       # models.py:1037-1051
            self.attention = tuple(torch.as_tensor(np.asarray(params["attention/" + n])).to(dtype) for n in ("w:0", "b:0", "v:0"))
        self.w0 = torch.as_tensor(np.asarray(params["embed_layer-0/w:0"])).to(dtype)
        self.b0 = torch.as_tensor(np.asarray(params["embed_layer-0/b:0"])).to(dtype)

    @torch.no_grad()
    def forward_batch(self, x):
        """x: [B, T, D] (all the same length) -> [B, 512].  models.py:470-495."""
        h = torch.as_tensor(x).to(self.dtype).transpose(1, 2)                    # NWC -> NCW
        for L in self.layers:
            h = F.conv1d(h, L["w"], L["b"], padding=L["pad"], dilation=L["dilation"])
            if self.act == "relu":
                h = torch.relu(h)
            elif self.act == "lrelu":
                h = F.leaky_relu(h, 0.2)
            else:
                h = torch.clamp(h, min=0) + L["alpha"][None, :, None] * torch.clamp(h, max=0)
            h = h * L["inv"][None, :, None] + L["shift"][None, :, None]
        if self.attention is not None:
            w, b, v = self.attention
            C = h.shape[1] // 2
            h1, h2 = h[:, :C, :].transpose(1, 2), h[:, C:, :].transpose(1, 2)        # [B, T, C]
            att = torch.softmax(torch.einsum("ijk,k->ij", torch.tanh(torch.einsum("ijk,kl->ijl", h1, w) + b), v), dim=1)
            h_m = torch.einsum("ijk,ij->ik", h2, att)
            h_s = torch.einsum("ijk,ij->ik", h2 * h2, att) - h_m * h_m
            stats = torch.cat([h_m, torch.sqrt(h_s + VAR2STD_EPSILON)], dim=1)
            return stats @ self.w0 + self.b0
        mean = h.mean(dim=2)
        var = ((h - mean[:, :, None]) ** 2).mean(dim=2)
        stats = torch.cat([mean, torch.sqrt(var + VAR2STD_EPSILON)], dim=1)
        return stats @ self.w0 + self.b0

    def forward(self, x):
        return self.forward_batch(np.asarray(x)[None])[0].numpy()

    def make_embedding_one(self, mat, min_chunk_size, chunk_size):
        """models.py:378-421 for one utterance (B=1 per call, as the reference runs it)."""
        plan = chunk_plan(mat.shape[0], min_chunk_size, chunk_size)
        if plan is None:
            return None
        acc, tot = 0, 0.0
        for start, length in plan:
            xv = self.forward(mat[start:start + length])
            tot += length
            acc = acc + length * xv
        return acc / tot
