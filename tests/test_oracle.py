"""The oracle is unpinned by the reference (no TF, no reference tests), so it is de-risked here:
two independent restatements must agree, plus properties the TF ops guarantee."""
import numpy as np
import pytest
import torch

from oracle import xvector_oracle as orc
from oracle.xvector_torch_cpu import TorchCpuXvector
from xvector_b200 import synthetic


def _params(topology, weight_set):
    t = orc.TOPOLOGIES[topology]
    return synthetic.make_params(t["kernel_sizes"], t["layer_sizes"], t["embedding_sizes"], weight_set=weight_set,
                                 activation=t.get("act", "relu"), pooling=t.get("pooling", "stats"))


@pytest.mark.parametrize("topology", ["ModelWithoutDropout", "ModelWithoutDropoutTdnn", "ModelWithoutDropoutPRelu",
                                      "ModelL2LossWithoutDropoutLRelu", "ModelL2LossWithoutDropoutLReluAttention"])
@pytest.mark.parametrize("weight_set", ["A", "B"])
def test_two_restatements_agree(topology, weight_set):
    # BASELINE config 1: single 200-frame x 23 utterance, seed 1
    x = synthetic.mfcc(1, 200)
    p = _params(topology, weight_set)
    ref = orc.forward(x, p, topology)
    assert ref.shape == (512,) and np.isfinite(ref).all()
    t64 = TorchCpuXvector(p, topology, dtype=torch.float64).forward(x)
    m = orc.parity_metrics(t64, ref)
    assert m["max_rel"] < 1e-9, m                      # fp64 vs fp64: independent code paths agree
    t32 = TorchCpuXvector(p, topology, dtype=torch.float32).forward(x)
    m32 = orc.parity_metrics(t32, ref)
    assert m32["max_rel"] < 2e-5 and m32["l2_rel"] < 2e-5, m32   # fp32 noise floor (SURVEY: 4-5e-7 typ.)


def test_conv_same_matches_explicit_definition():
    rng = np.random.default_rng(0)
    for k, d in [(5, 1), (3, 2), (3, 3), (7, 1), (1, 1)]:
        x = rng.standard_normal((11, 4))
        w = rng.standard_normal((k, 4, 3))
        y = orc.conv1d_same(x, w, d)
        half = (k - 1) // 2
        want = np.zeros((11, 3))
        for t in range(11):
            for j in range(k):
                tt = t + (j - half) * d
                if 0 <= tt < 11:
                    want[t] += x[tt] @ w[j]
        assert np.allclose(y, want, atol=1e-12)


def test_dense_with_zeroed_taps_equals_dilated():
    # kernel 5 dense with taps 1,3 zeroed == kernel 3 dilation 2 (SURVEY 8c de-risking property)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((40, 6))
    w3 = rng.standard_normal((3, 6, 5))
    w5 = np.zeros((5, 6, 5))
    w5[0], w5[2], w5[4] = w3[0], w3[1], w3[2]
    assert np.allclose(orc.conv1d_same(x, w5, 1), orc.conv1d_same(x, w3, 2), atol=1e-12)


def test_time_shift_equivariance_away_from_edges():
    p = _params("ModelWithoutDropout", "B")
    x = synthetic.mfcc(3, 120).astype(np.float64)
    _, layers_a, _ = orc.forward(x, p, return_layers=True)
    _, layers_b, _ = orc.forward(x[5:], p, return_layers=True)
    # receptive field is +-(2+2+3)=7 frames: interior frames agree after the 5-frame shift
    a, b = layers_a[-1], layers_b[-1]
    assert np.allclose(a[5 + 8:-8], b[8:-8], atol=1e-9)
    assert not np.allclose(a[5:5 + 7], b[:7], atol=1e-6)      # edges differ: zero SAME padding


def test_zero_input_k1_stack_gives_constant_frames_and_eps_std():
    p = _params("ModelWithoutDropout", "B")
    x = np.zeros((50, 23))
    _, layers, stats = orc.forward(x, p, return_layers=True)
    h = layers[-1]
    # interior frames are identical, so pooled std of the k=1 tail would be sqrt(eps) without edges;
    # here edges (zero padding) perturb at most 7 frames each side
    assert np.allclose(h[10:40], h[10], atol=1e-12)
    mean, std = stats[:1536], stats[1536:]
    assert (std >= np.sqrt(orc.VAR2STD_EPSILON) - 1e-12).all()
    assert mean.shape == (1536,)


def test_batchnorm_eval_formula():
    z = np.array([[1.0, -2.0]])
    y = orc.batch_norm_eval(z, gamma=np.array([2.0, 0.5]), beta=np.array([0.1, -0.1]),
                            mean=np.array([0.5, 1.0]), variance=np.array([4.0, 0.25]))
    want = (z - np.array([0.5, 1.0])) / np.sqrt(np.array([4.0, 0.25]) + 1e-3) * np.array([2.0, 0.5]) + np.array([0.1, -0.1])
    assert np.allclose(y, want, atol=1e-12)


def test_stats_pool_is_population_variance_mean_first():
    h = np.array([[1.0, 10.0], [3.0, 10.0]])
    s = orc.stats_pool(h)
    assert np.allclose(s, [2.0, 10.0, np.sqrt(1.0 + 1e-5), np.sqrt(1e-5)])


@pytest.mark.parametrize("rows,min_c,chunk,want", [
    (0, 25, 10000, None), (24, 25, 10000, None),
    (25, 25, 10000, [(0, 25)]), (10000, 25, 10000, [(0, 10000)]),
    (10001, 25, 10000, [(0, 10000)]),                       # trailing 1-frame chunk dropped
    (25000, 25, 10000, [(0, 10000), (10000, 10000), (20000, 5000)]),
    (300, 100, -1, [(0, 300)]),                              # CLI defaults
])
def test_chunk_plan_follows_make_embedding(rows, min_c, chunk, want):
    assert orc.chunk_plan(rows, min_c, chunk) == want


def test_chunk_average_is_frame_weighted():
    p = _params("ModelWithoutDropoutTdnn", "A")
    x = synthetic.mfcc(9, 130).astype(np.float64)
    got = orc.make_embedding_one(x, p, "ModelWithoutDropoutTdnn", 25, 50)
    parts = [orc.forward(x[s:s + n], p, "ModelWithoutDropoutTdnn") for s, n in [(0, 50), (50, 50), (100, 30)]]
    want = (50 * parts[0] + 50 * parts[1] + 30 * parts[2]) / 130.0
    assert np.allclose(got, want, rtol=1e-12)
    assert orc.make_embedding_one(x[:10], p, "ModelWithoutDropoutTdnn", 25, 50) is None


def test_leaky_and_parametric_relu_definitions():
    # tf.nn.leaky_relu(x, 0.2) = max(x, 0.2 x); prelu = max(0,x) + alpha*min(0,x) per channel (tf_block.py:47)
    y = np.array([[-2.0, 3.0], [0.5, -1.0]])
    assert np.allclose(orc.activation(y, "lrelu"), [[-0.4, 3.0], [0.5, -0.2]])
    assert np.allclose(orc.activation(y, "prelu", np.array([0.1, -0.5])), [[-0.2, 3.0], [0.5, 0.5]])
    assert np.allclose(orc.activation(y, "relu"), [[0.0, 3.0], [0.5, 0.0]])


def test_attention_pooling_definition():
    # reference models.py:1037-1051
    rng = np.random.default_rng(4)
    T, C = 37, 6
    h = rng.standard_normal((T, 2 * C))
    p = {"attention/w:0": rng.standard_normal((C, C)), "attention/b:0": rng.standard_normal(C), "attention/v:0": rng.standard_normal(C)}
    got = orc.attention_pool(h, p)
    score = np.array([np.tanh(h[t, :C] @ p["attention/w:0"] + p["attention/b:0"]) @ p["attention/v:0"] for t in range(T)])
    a = np.exp(score) / np.exp(score).sum()
    assert abs(a.sum() - 1) < 1e-12
    h_m = sum(a[t] * h[t, C:] for t in range(T))
    h_s = sum(a[t] * h[t, C:] ** 2 for t in range(T)) - h_m ** 2
    np.testing.assert_allclose(got, np.concatenate([h_m, np.sqrt(h_s + 1e-5)]), rtol=1e-12)
    # v = 0 -> uniform attention -> plain statistics pooling of the second half
    p0 = dict(p, **{"attention/v:0": np.zeros(C)})
    np.testing.assert_allclose(orc.attention_pool(h, p0), orc.stats_pool(h[:, C:]), rtol=1e-10)
    # constant frames -> zero weighted variance -> std = sqrt(1e-5)
    hc = np.tile(rng.standard_normal(2 * C), (T, 1))
    np.testing.assert_allclose(orc.attention_pool(hc, p)[C:], np.sqrt(1e-5), rtol=1e-6)
