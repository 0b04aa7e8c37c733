"""CPU oracle for ONE training minibatch  --  TEST INFRASTRUCTURE ONLY (same rules as xvector_oracle.py).

Restates, in torch fp64 with autograd, what the reference evaluates in
``sess.run([self.optimizer, self.loss, self.accuracy], feed_dict)`` (local/tf/models.py:263) on the
graph built by ``ModelWithoutDropout.build_model`` / ``ModelWithoutDropoutTdnn.build_model``
(local/tf/models.py:441-534 / :543-639) with ``phase=True``:

  * frame layers: conv SAME -> +b -> relu -> batch_norm_wrapper TRAINING branch
    (local/tf/tf_block.py:18-23: batch mean / *population* variance over [B, T] via tf.nn.moments,
    moving statistics ``pop = pop*decay + batch*(1-decay)`` with decay=0.95 (models.py:480), then
    tf.nn.batch_normalization with the BATCH statistics, eps 1e-3);
  * statistics pooling (models.py:485-486), two segment layers xw_plus_b -> relu -> BN(train)
    (models.py:489-499), output xw_plus_b (models.py:502-508);
  * loss = mean softmax cross-entropy (models.py:512-514), accuracy (models.py:521-523);
  * tf.train.AdamOptimizer(learning_rate).minimize (models.py:516-519): TF's Adam --
    ``lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    var -= lr_t * m / (sqrt(v) + eps)`` with b1=0.9, b2=0.999, eps=1e-8 (TF 1.x documentation).

PARITY STATUS: unpinned, for the same reason as the forward oracle (TensorFlow absent, the reference
ships no fixtures).  The autograd gradients are cross-checked against central finite differences of the
restated loss (tests/test_train_oracle.py); the forward half must agree with oracle/xvector_oracle.py
when the batch statistics are substituted for the moving ones.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .xvector_oracle import BN_EPSILON, TOPOLOGIES, VAR2STD_EPSILON

BN_DECAY = 0.95                # models.py:480,497  batch_norm_wrapper(h, decay=0.95, ...)
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8


def trainable_names(topology, params):
    """Trainable variables in graph-construction order (w, b, gamma, beta per layer; output w, b)."""
    topo = TOPOLOGIES[topology] if isinstance(topology, str) else topology
    names = []
    for i in range(len(topo["kernel_sizes"])):
        s = "frame_level_info_layer-%d/" % i
        names += [s + "w:0", s + "b:0", s + "gamma:0", s + "beta:0"]
    for i in range(len(topo["embedding_sizes"])):
        s = "embed_layer-%d/" % i
        names += [s + "w:0", s + "b:0", s + "gamma:0", s + "beta:0"]
    names += ["output/w:0", "output/b:0"]
    return [n for n in names if n in params]


def _act(h, act):
    """relu (models.py:479) or tf.nn.leaky_relu(alpha=0.2) (models.py:912), frame and segment layers alike."""
    if act == "relu":
        return torch.relu(h)
    if act == "lrelu":
        return F.leaky_relu(h, 0.2)
    raise ValueError("the training oracle covers relu / lrelu, not %r" % act)


def l2_term(p, beta):
    """ModelL2Loss* (models.py:930-951, 962): beta * (0.1 l2(embed-0 w, b) + l2(embed-1 w, b) + l2(output w, b)), l2 = sum(x^2)/2."""
    l2 = 0.1 * ((p["embed_layer-0/w:0"] ** 2).sum() + (p["embed_layer-0/b:0"] ** 2).sum()) / 2
    l2 = l2 + ((p["embed_layer-1/w:0"] ** 2).sum() + (p["embed_layer-1/b:0"] ** 2).sum()) / 2
    l2 = l2 + ((p["output/w:0"] ** 2).sum() + (p["output/b:0"] ** 2).sum()) / 2
    return beta * l2


def _bn_train(h, gamma, beta, axes, eps=BN_EPSILON):
    mean = h.mean(dim=axes)
    var = ((h - mean) ** 2).mean(dim=axes)                      # tf.nn.moments: population variance
    inv = gamma / torch.sqrt(var + eps)
    return h * inv + (beta - mean * inv), mean, var             # tf.nn.batch_normalization


def _fp16_storage(t):
    """Round to fp16 and back with a straight-through gradient: emulates WHERE the CUDA path stores 16-bit values
    (spliced input, conv weights, relu output, BatchNorm output of the frame layers) so that the discrete ReLU masks
    of both computations coincide; gradients themselves stay fp64."""
    return t + (t.half().to(t.dtype) - t).detach()


def forward_backward(x, labels, params, topology="ModelWithoutDropoutTdnn", dtype=torch.float64,
                     return_intermediates=False, fp16_storage=False, bn_eps=BN_EPSILON, l2_beta=None):
    """One minibatch.  x: [B, T, D]; labels: [B] ints; params: dict of numpy arrays by TF variable name.

    Returns dict(loss, accuracy, grads{name: np}, batch_stats{scope: (mean, var)}, moving{name: np})
    where ``moving`` holds the updated moving mean/variance (tf_block.py:20-21).
    ``fp16_storage=True`` rounds the frame-level operands to fp16 at the points the CUDA path does (see
    _fp16_storage); the default is the exact fp64 restatement of the reference.  ``bn_eps`` is the reference's 1e-3;
    tests also use a large value to make the gradient well conditioned (small-variance channels amplify any forward
    rounding ~100x at 1e-3), which checks the backward arithmetic tightly.
    """
    q = _fp16_storage if fp16_storage else (lambda t: t)
    act = (TOPOLOGIES[topology] if isinstance(topology, str) else topology).get("act", "relu")
    if l2_beta is None:
        l2_beta = (TOPOLOGIES[topology] if isinstance(topology, str) else topology).get("l2_beta", 0.0)
    topo = TOPOLOGIES[topology] if isinstance(topology, str) else topology
    names = trainable_names(topo, params)
    p = {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in params.items()}
    for n in names:
        p[n].requires_grad_(True)
    h = q(torch.tensor(np.asarray(x), dtype=dtype))             # [B, T, D]
    n_frame = len(topo["kernel_sizes"])
    inter = {}
    batch_stats = {}
    for i, (k, d) in enumerate(zip(topo["kernel_sizes"], topo["dilations"])):
        s = "frame_level_info_layer-%d/" % i
        w = p[s + "w:0"]                                          # [k, Cin, Cout]
        z = F.conv1d(h.transpose(1, 2), q(w).permute(2, 1, 0), padding=(k - 1) // 2 * d, dilation=d).transpose(1, 2)
        r = q(_act(z + p[s + "b:0"], act))
        h, mean, var = _bn_train(r, p[s + "gamma:0"], p[s + "beta:0"], (0, 1), bn_eps)
        if i < n_frame - 1:
            h = q(h)
        batch_stats[s] = (mean.detach().numpy(), var.detach().numpy())
        if return_intermediates:
            r.retain_grad(); h.retain_grad()
            inter[s + "relu"] = r
            inter[s + "bn"] = h
    mean_t = h.mean(dim=1)
    var_t = ((h - mean_t[:, None, :]) ** 2).mean(dim=1)
    h = torch.cat([mean_t, torch.sqrt(var_t + VAR2STD_EPSILON)], dim=1)    # [B, 2C]
    if return_intermediates:
        h.retain_grad()
        inter["stats"] = h
    for i in range(len(topo["embedding_sizes"])):
        s = "embed_layer-%d/" % i
        z = h @ p[s + "w:0"] + p[s + "b:0"]
        r = _act(z, act)
        h, mean, var = _bn_train(r, p[s + "gamma:0"], p[s + "beta:0"], (0,), bn_eps)
        batch_stats[s] = (mean.detach().numpy(), var.detach().numpy())
        if return_intermediates:
            inter[s + "scores"] = z
            inter[s + "bn"] = h
    logits = h @ p["output/w:0"] + p["output/b:0"]
    lab = torch.tensor(np.asarray(labels), dtype=torch.long)
    loss = F.cross_entropy(logits, lab, reduction="mean")       # one-hot labels (models.py:164-169)
    if l2_beta:
        loss = loss + l2_term(p, l2_beta)                       # tf.reduce_mean(loss + beta * l2_loss), models.py:961
    acc = (logits.argmax(dim=1) == lab).double().mean()
    loss.backward()
    grads = {n: p[n].grad.detach().numpy().copy() for n in names}
    moving = {}
    for s, (mean, var) in batch_stats.items():
        moving[s + "mean:0"] = np.asarray(params[s + "mean:0"], np.float64) * BN_DECAY + mean * (1 - BN_DECAY)
        moving[s + "variance:0"] = np.asarray(params[s + "variance:0"], np.float64) * BN_DECAY + var * (1 - BN_DECAY)
    out = dict(loss=float(loss.detach()), accuracy=float(acc), grads=grads, batch_stats=batch_stats, moving=moving,
               logits=logits.detach().numpy())
    if return_intermediates:
        out["intermediates"] = {k: v.detach().numpy() for k, v in inter.items()}
        out["intermediate_grads"] = {k: v.grad.numpy() for k, v in inter.items()
                                     if (k.endswith("/relu") or k.endswith("/bn") or k == "stats") and k.startswith(("frame", "stats"))}
    return out


def loss_only(x, labels, params, topology, l2_beta=None):
    """Loss of the restated graph, numpy fp64, no autograd (for finite differences)."""
    with torch.no_grad():
        topo = TOPOLOGIES[topology] if isinstance(topology, str) else topology
        act = topo.get("act", "relu")
        if l2_beta is None:
            l2_beta = topo.get("l2_beta", 0.0)
        p = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in params.items()}
        h = torch.tensor(np.asarray(x), dtype=torch.float64)
        for i, (k, d) in enumerate(zip(topo["kernel_sizes"], topo["dilations"])):
            s = "frame_level_info_layer-%d/" % i
            z = F.conv1d(h.transpose(1, 2), p[s + "w:0"].permute(2, 1, 0), padding=(k - 1) // 2 * d, dilation=d).transpose(1, 2)
            h, _, _ = _bn_train(_act(z + p[s + "b:0"], act), p[s + "gamma:0"], p[s + "beta:0"], (0, 1))
        mean_t = h.mean(dim=1)
        var_t = ((h - mean_t[:, None, :]) ** 2).mean(dim=1)
        h = torch.cat([mean_t, torch.sqrt(var_t + VAR2STD_EPSILON)], dim=1)
        for i in range(len(topo["embedding_sizes"])):
            s = "embed_layer-%d/" % i
            h, _, _ = _bn_train(_act(h @ p[s + "w:0"] + p[s + "b:0"], act), p[s + "gamma:0"], p[s + "beta:0"], (0,))
        logits = h @ p["output/w:0"] + p["output/b:0"]
        loss = F.cross_entropy(logits, torch.tensor(np.asarray(labels), dtype=torch.long))
        if l2_beta:
            loss = loss + l2_term(p, l2_beta)
        return float(loss)


def adam_init(params, names):
    return dict(t=0, m={n: np.zeros_like(np.asarray(params[n], np.float64)) for n in names},
                v={n: np.zeros_like(np.asarray(params[n], np.float64)) for n in names})


def adam_step(params, grads, slots, lr):
    """tf.train.AdamOptimizer._apply_dense (defaults b1=.9, b2=.999, eps=1e-8); updates in place, fp64."""
    slots["t"] += 1
    t = slots["t"]
    lr_t = lr * np.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t)
    for n, g in grads.items():
        g = np.asarray(g, np.float64)
        slots["m"][n] = ADAM_B1 * slots["m"][n] + (1 - ADAM_B1) * g
        slots["v"][n] = ADAM_B2 * slots["v"][n] + (1 - ADAM_B2) * g * g
        params[n] = np.asarray(params[n], np.float64) - lr_t * slots["m"][n] / (np.sqrt(slots["v"][n]) + ADAM_EPS)
    return params


def evaluate(x, labels, params, topology="ModelWithoutDropoutTdnn"):
    """(loss, accuracy) with ``phase=False``: the moving statistics are used everywhere (tf_block.py:25-26) -- what
    ``Model.eval`` fetches per minibatch (models.py:338-339).  fp64, no gradients."""
    from .xvector_oracle import batch_norm_eval, forward
    topo = TOPOLOGIES[topology] if isinstance(topology, str) else topology
    p = {k: np.asarray(v, np.float64) for k, v in params.items()}
    losses, correct = [], []
    for xb, lab in zip(np.asarray(x, np.float64), np.asarray(labels)):
        h = forward(xb, p, topo)                                   # embed_layer-0/scores
        for i in range(len(topo["embedding_sizes"])):
            s = "embed_layer-%d/" % i
            if i > 0:
                h = h @ p[s + "w:0"] + p[s + "b:0"]
            h = batch_norm_eval(np.maximum(h, 0.0), p[s + "gamma:0"], p[s + "beta:0"], p[s + "mean:0"], p[s + "variance:0"])
        logits = h @ p["output/w:0"] + p["output/b:0"]
        m = logits.max()
        losses.append(np.log(np.exp(logits - m).sum()) + m - logits[int(lab)])
        correct.append(float(int(np.argmax(logits)) == int(lab)))
    return float(np.mean(losses)), float(np.mean(correct))
