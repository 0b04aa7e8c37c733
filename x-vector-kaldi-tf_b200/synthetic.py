"""Deterministic synthetic MFCC features and model weights.

There is no network for corpora or checkpoints, so every test/bench input is generated
here from fixed seeds.  Two weight sets:

  * ``"A"`` (model_0-like): what ``build_model`` initialises (reference
    local/tf/models.py:473-474,492-493 and local/tf/tf_block.py:10-15): truncated normal
    (sigma=0.1, resampled beyond 2 sigma) weights, bias 0.1, identity BatchNorm.
  * ``"B"`` (trained-like): He-normal weights, random bias and non-trivial BatchNorm
    statistics, so the fused ``relu(acc+b)*inv + shift`` epilogue is really exercised.
"""
from __future__ import annotations

import numpy as np

FEAT_DIM = 23          # conf/mfcc.conf: --num-ceps=23


def mfcc(seed, num_frames, feat_dim=FEAT_DIM):
    """One utterance [T, feat_dim] float32: c0-heavy, zero-mean (post-CMN-like)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    scale = (12.0 / np.sqrt(1.0 + np.arange(feat_dim))).astype(np.float32)
    return rng.standard_normal((num_frames, feat_dim), dtype=np.float32) * scale


def mfcc_batch(seed, lengths, feat_dim=FEAT_DIM):
    """Concatenated utterances [sum(T), feat_dim] float32 drawn from ONE stream."""
    lengths = np.asarray(lengths, dtype=np.int64)
    return mfcc(seed, int(lengths.sum()), feat_dim)


def lengths_uniform(seed, n, lo=200, hi=1000):
    """BASELINE config 3/4 length law: integers uniform in [lo, hi]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(lo, hi + 1, size=n).astype(np.int32)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return z ^ (z >> np.uint64(31))


def counter_mfcc(seed, utt_id, num_frames, feat_dim=FEAT_DIM):
    """The utterance the DEVICE generator (xv_synth_mfcc, csrc/synth.cuh) produces for (seed, utt_id): same integer recipe,
    bit for bit -- Irwin-Hall(4) of the 32-bit halves of two splitmix64 words per value, scaled like ``mfcc``."""
    with np.errstate(over="ignore"):
        frame = np.arange(num_frames, dtype=np.uint64)[:, None]
        d = np.arange(feat_dim, dtype=np.uint64)[None, :]
        base = np.uint64(int(seed) & (2 ** 64 - 1)) ^ (((np.uint64(utt_id) << np.uint64(20)) | frame) * np.uint64(0x9E3779B97F4A7C15))
        h1 = _splitmix64(base + d)
        h2 = _splitmix64(h1)
        m32 = np.uint64(0xFFFFFFFF)
        s = ((h1 & m32) + (h1 >> np.uint64(32)) + (h2 & m32) + (h2 >> np.uint64(32))).astype(np.int64) - (np.int64(1) << np.int64(33))
    k = (np.sqrt(3.0) / 4294967296.0 * 12.0 / np.sqrt(1.0 + np.arange(feat_dim, dtype=np.float64))).astype(np.float32)
    return s.astype(np.float32) * k[None, :]


def _trunc_normal(rng, shape, sigma):
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * sigma).astype(np.float32)


def make_params(kernel_sizes, layer_sizes, embedding_sizes, feat_dim=FEAT_DIM, num_classes=0,
                weight_set="A", seed=100, activation="relu", init="trunc_normal", pooling="stats"):
    """Parameter dict keyed by the reference's TF variable names (models.py:199-210)."""
    p = {}
    prev = feat_dim

    def bn(scope, dim, rng):
        if weight_set == "A":
            p[scope + "gamma:0"] = np.ones(dim, np.float32)
            p[scope + "beta:0"] = np.zeros(dim, np.float32)
            p[scope + "mean:0"] = np.zeros(dim, np.float32)
            p[scope + "variance:0"] = np.ones(dim, np.float32)
        else:
            p[scope + "gamma:0"] = rng.uniform(0.5, 1.5, dim).astype(np.float32)
            p[scope + "beta:0"] = (0.1 * rng.standard_normal(dim)).astype(np.float32)
            p[scope + "mean:0"] = (0.3 + 0.1 * rng.standard_normal(dim)).astype(np.float32)
            p[scope + "variance:0"] = rng.uniform(0.5, 2.0, dim).astype(np.float32)

    for i, (k, width) in enumerate(zip(kernel_sizes, layer_sizes)):
        rng = np.random.Generator(np.random.PCG64(seed + i))
        s = "frame_level_info_layer-%d/" % i
        if weight_set == "A" and init == "he":       # ...ReluHeInit: models.py:1155-1163 (he_normal w, he_uniform b)
            fan_in = k * prev
            p[s + "w:0"] = _trunc_normal(rng, (k, prev, width), np.sqrt(2.0 / fan_in))
            lim = np.sqrt(6.0 / fan_in)
            p[s + "b:0"] = rng.uniform(-lim, lim, width).astype(np.float32)
        elif weight_set == "A":
            p[s + "w:0"] = _trunc_normal(rng, (k, prev, width), 0.1)
            p[s + "b:0"] = np.full(width, 0.1, np.float32)
        else:
            p[s + "w:0"] = (rng.standard_normal((k, prev, width)) * np.sqrt(2.0 / (k * prev))).astype(np.float32)
            p[s + "b:0"] = rng.uniform(-0.1, 0.1, width).astype(np.float32)
        if activation == "prelu":      # tf_block.py:38-47: per-channel slope, constant_initializer(0.1)
            p[s + "prelu/prelu:0"] = (np.full(width, 0.1, np.float32) if weight_set == "A"
                                      else rng.uniform(-0.2, 0.5, width).astype(np.float32))
        bn(s, width, rng)
        prev = width
    if pooling == "attention":       # models.py:1035-1041: square "attention/w" (trunc normal 0.1), b = v = 0.1; stats width = prev
        C = prev // 2
        rng = np.random.Generator(np.random.PCG64(seed + 40))
        if weight_set == "A":
            p["attention/w:0"] = _trunc_normal(rng, (C, C), 0.1)
            p["attention/b:0"] = np.full(C, 0.1, np.float32)
            p["attention/v:0"] = np.full(C, 0.1, np.float32)
        else:
            p["attention/w:0"] = (rng.standard_normal((C, C)) * np.sqrt(1.0 / C)).astype(np.float32)
            p["attention/b:0"] = rng.uniform(-0.1, 0.1, C).astype(np.float32)
            p["attention/v:0"] = (rng.standard_normal(C) * 0.05).astype(np.float32)
    else:
        prev *= 2
    for i, width in enumerate(embedding_sizes):
        rng = np.random.Generator(np.random.PCG64(seed + 50 + i))
        s = "embed_layer-%d/" % i
        if activation == "prelu":
            p[s + "prelu/prelu:0"] = np.full(width, 0.1, np.float32)
        if weight_set == "A" and init == "he":
            p[s + "w:0"] = _trunc_normal(rng, (prev, width), np.sqrt(2.0 / prev))
            lim = np.sqrt(6.0 / prev)
            p[s + "b:0"] = rng.uniform(-lim, lim, width).astype(np.float32)
        elif weight_set == "A":
            p[s + "w:0"] = _trunc_normal(rng, (prev, width), 0.1)
            p[s + "b:0"] = np.full(width, 0.1, np.float32)
        else:
            p[s + "w:0"] = (rng.standard_normal((prev, width)) * np.sqrt(2.0 / prev)).astype(np.float32)
            p[s + "b:0"] = rng.uniform(-0.1, 0.1, width).astype(np.float32)
        bn(s, width, rng)
        prev = width
    if num_classes > 0:
        rng = np.random.Generator(np.random.PCG64(seed + 90))
        lim = np.sqrt(6.0 / (prev + num_classes))            # xavier uniform (models.py:504-505)
        p["output/w:0"] = rng.uniform(-lim, lim, (prev, num_classes)).astype(np.float32)
        p["output/b:0"] = np.full(num_classes, 0.1, np.float32)
    return p


def synthetic_vad(rng, num_frames, voiced_fraction=0.75, mean_run=40):
    """A 0/1 float32 VAD track of alternating voiced / unvoiced runs (what compute-vad's energy decisions look like):
    the second input of the feature front end (select-voiced-frames, reference local/tf/extract_xvectors.sh:68)."""
    out = np.zeros(num_frames, np.float32)
    t = 0
    state = rng.random() < voiced_fraction
    while t < num_frames:
        mean = mean_run * (voiced_fraction if state else (1.0 - voiced_fraction)) * 2.0
        run = 1 + int(rng.exponential(max(mean, 1.0)))
        out[t:t + run] = 1.0 if state else 0.0
        t += run
        state = not state
    return out
