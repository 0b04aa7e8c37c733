"""CPU oracle for the x-vector extraction hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the forward pass that the reference
(BUTSpeechFIT/x-vector-kaldi-tf @ 5249e7d) evaluates through TensorFlow 1.x in
``sess.run(self.embedding[0], ...)`` (local/tf/models.py:412-415).  It is the
checker for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does.

PARITY STATUS
    * forward arithmetic: **parity unpinned**.  The arithmetic lives in the
      third-party dependency ``tensorflow`` 1.x (version not pinned by the
      reference: no requirements file; README.md:29-33), which is absent from
      /root/reference and from this image, and the reference ships no tests,
      golden vectors or fixtures for this path.  The
      restatement is anchored on the reference's own call sites (cited per
      function below) and de-risked by an independent second restatement
      (``oracle/xvector_torch_cpu.py``, different code path: oneDNN conv1d) that
      must agree with this one, plus property tests (tests/test_oracle.py).
    * Kaldi ark I/O: pinned.  The reference's ``local/tf/kaldi_io.py`` is pure
      numpy and was imported in the build container to generate the byte-level
      fixtures under ``tests/golden/`` (tests/golden/make_golden_ark.py).

TensorFlow op semantics restated here (TF 1.x public documentation):
    * ``tf.nn.conv1d(padding="SAME", stride=1)`` / ``tf.nn.convolution(dilation_rate=d,
      padding="SAME")``: cross-correlation (no kernel flip), filter layout
      ``[k, C_in, C_out]``, zero padding of ``(k-1)*d`` in total, split
      ``left=(k-1)*d//2`` / ``right=rest`` -- symmetric for the odd kernel sizes the
      reference uses; output length equals input length.
    * ``tf.nn.batch_normalization(x, mean, var, offset, scale, eps)``:
      ``inv = rsqrt(var + eps) * scale ; y = x * inv + (offset - mean * inv)``.
    * ``tf.nn.moments(h, 1)``: mean and *population* variance over the time axis.
    * ``tf.nn.xw_plus_b(x, w, b)``: ``x @ w + b``.
"""
from __future__ import annotations

import numpy as np

VAR2STD_EPSILON = 0.00001      # local/tf/models.py:16
BN_EPSILON = 1e-3              # local/tf/tf_block.py:9  (batch_norm_wrapper default epsilon)

# Topologies (class name -> per-layer kernel sizes / dilations / widths).
#   ModelWithoutDropout      local/tf/models.py:443-445   (recipe default, run_xvector.sh:90)
#   ModelWithoutDropoutTdnn  local/tf/models.py:545-548   (north-star splice)
#   Model                    local/tf/models.py:27-29     (same layer shapes as ModelWithoutDropout)
TOPOLOGIES = {
    "Model": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                  layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512]),
    "ModelWithoutDropout": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                                layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512]),
    "ModelWithoutDropoutTdnn": dict(kernel_sizes=[5, 3, 3, 1, 1], dilations=[1, 2, 3, 1, 1],
                                    layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512]),
    # frame-layer nonlinearity variants (local/tf/models.py:643-744 prelu, :866-983 leaky_relu(0.2))
    "ModelWithoutDropoutPRelu": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                                     layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512], act="prelu"),
    "ModelL2LossWithoutDropoutLRelu": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                                           layer_sizes=[512, 512, 512, 512, 1536], embedding_sizes=[512, 512], act="lrelu",
                                           l2_beta=0.0002),          # training only: models.py:876,961
    # self-attention pooling (local/tf/models.py:990-1051): last frame layer 6*512 wide, split into scores input | pooled half
    "ModelL2LossWithoutDropoutLReluAttention": dict(kernel_sizes=[5, 5, 7, 1, 1], dilations=[1, 1, 1, 1, 1],
                                                    layer_sizes=[512, 512, 512, 512, 3072], embedding_sizes=[512, 512],
                                                    act="lrelu", pooling="attention"),
}


def conv1d_same(x, w, dilation=1):
    """Time convolution of one utterance, TF "SAME" padding.

    Follows local/tf/models.py:476 (``tf.nn.conv1d``) and :579 (``tf.nn.convolution`` with
    ``dilation_rate``).  x: [T, C_in], w: [k, C_in, C_out] -> [T, C_out].
    ``y[t] = sum_j x[t + j*d - pad_left] @ w[j]`` with out-of-range rows reading zero.
    """
    T = x.shape[0]
    k, _, c_out = w.shape
    pad_left = ((k - 1) * dilation) // 2
    y = np.zeros((T, c_out), dtype=np.result_type(x.dtype, w.dtype))
    for j in range(k):
        shift = j * dilation - pad_left
        lo = max(0, -shift)
        hi = min(T, T - shift)
        if hi > lo:
            y[lo:hi] += x[lo + shift:hi + shift] @ w[j]
    return y


def batch_norm_eval(z, gamma, beta, mean, variance, eps=BN_EPSILON):
    """Evaluation branch of batch_norm_wrapper (local/tf/tf_block.py:25-26)."""
    inv = gamma / np.sqrt(variance + eps)
    return z * inv + (beta - mean * inv)


def activation(y, act="relu", alpha=None):
    """relu: models.py:479; lrelu(0.2): models.py:912; prelu: tf_block.py:38-47."""
    if act == "relu":
        return np.maximum(y, 0.0)
    if act == "lrelu":
        return np.maximum(y, 0.0) + 0.2 * np.minimum(y, 0.0)
    if act == "prelu":
        return np.maximum(y, 0.0) + alpha * np.minimum(y, 0.0)
    raise ValueError(act)


def frame_layer(x, params, i, dilation, act="relu"):
    """One frame-level layer: conv -> +bias -> ReLU -> BatchNorm (local/tf/models.py:470-482)."""
    s = "frame_level_info_layer-%d/" % i
    y = conv1d_same(x, params[s + "w:0"], dilation) + params[s + "b:0"]
    y = activation(y, act, params.get(s + "prelu/prelu:0"))
    return batch_norm_eval(y, params[s + "gamma:0"], params[s + "beta:0"],
                           params[s + "mean:0"], params[s + "variance:0"])


def stats_pool(h):
    """Statistics pooling (local/tf/models.py:485-486): [T, C] -> [2C], mean then std."""
    mean = h.mean(axis=0)
    var = ((h - mean) ** 2).mean(axis=0)          # tf.nn.moments: population variance
    return np.concatenate([mean, np.sqrt(var + VAR2STD_EPSILON)])


def attention_pool(h, params):
    """Self-attention statistics pooling (local/tf/models.py:1037-1051): [T, 2C] -> [2C].

    ``h1, h2 = split(h, 2)``; ``attention = softmax_t( tanh(h1 @ w + b) @ v )``; ``h_m = sum_t a_t h2[t]``;
    ``h_s = sum_t a_t h2[t]^2 - h_m^2``; result ``[h_m | sqrt(h_s + 1e-5)]``.
    """
    C = h.shape[1] // 2
    h1, h2 = h[:, :C], h[:, C:]
    non_linearity = np.tanh(h1 @ params["attention/w:0"] + params["attention/b:0"])
    score = non_linearity @ params["attention/v:0"]
    e = np.exp(score - score.max())
    attention = e / e.sum()                                   # tf.nn.softmax over the time axis
    h_m = attention @ h2
    h_s = attention @ (h2 ** 2) - h_m ** 2
    return np.concatenate([h_m, np.sqrt(h_s + VAR2STD_EPSILON)])


def forward(x, params, topology="ModelWithoutDropout", dtype=np.float64, return_layers=False):
    """x-vector of ONE chunk: what ``sess.run(embedding[0])`` returns (models.py:158,414).

    x: [T, D].  ``embedding[0]`` is ``embed_layer-0/scores:0`` = the first segment-level
    affine output *before* its ReLU/BN (models.py:495).
    """
    topo = TOPOLOGIES[topology] if isinstance(topology, str) else topology
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    h = np.asarray(x, dtype=dtype)
    layers = []
    for i, d in enumerate(topo["dilations"]):
        h = frame_layer(h, p, i, d, topo.get("act", "relu"))
        layers.append(h)
    stats = attention_pool(h, p) if topo.get("pooling") == "attention" else stats_pool(h)
    emb = stats @ p["embed_layer-0/w:0"] + p["embed_layer-0/b:0"]     # tf.nn.xw_plus_b, models.py:495
    if return_layers:
        return emb, layers, stats
    return emb


def chunk_plan(num_rows, min_chunk_size, chunk_size):
    """Chunk boundaries chosen by make_embedding (local/tf/models.py:378-409).

    Returns None when the utterance is skipped (zero rows or < min_chunk_size), else a
    list of (start, length) for the chunks that are actually run through the network.
    """
    if num_rows == 0 or num_rows < min_chunk_size:
        return None
    this_chunk = chunk_size
    if num_rows < chunk_size:
        this_chunk = num_rows
    elif chunk_size == -1:
        this_chunk = num_rows
    n = int(np.ceil(num_rows / float(this_chunk)))
    plan = []
    for c in range(n):
        offset = min(this_chunk, num_rows - c * this_chunk)
        if offset < min_chunk_size:
            continue
        plan.append((c * this_chunk, offset))
    return plan


def make_embedding_one(mat, params, topology, min_chunk_size, chunk_size, dtype=np.float64):
    """Frame-weighted average of chunk x-vectors (local/tf/models.py:398-421); None if skipped."""
    plan = chunk_plan(mat.shape[0], min_chunk_size, chunk_size)
    if plan is None:
        return None
    xvector_avg = 0
    tot_weight = 0.0
    for start, length in plan:
        xvector = forward(mat[start:start + length], params, topology, dtype=dtype)
        tot_weight += length
        xvector_avg = xvector_avg + length * xvector
    return xvector_avg / tot_weight


def parity_metrics(e, r):
    """SURVEY.md section 8(d) parity metric.  e, r: [N, 512] (computed, oracle)."""
    e = np.atleast_2d(np.asarray(e, dtype=np.float64))
    r = np.atleast_2d(np.asarray(r, dtype=np.float64))
    per_utt = np.abs(e - r).max(axis=1) / np.abs(r).max(axis=1)
    l2 = np.linalg.norm(e - r) / np.linalg.norm(r)
    cos = (e * r).sum(axis=1) / (np.linalg.norm(e, axis=1) * np.linalg.norm(r, axis=1))
    return dict(max_rel=float(per_utt.max()), l2_rel=float(l2), min_cos=float(cos.min()))
