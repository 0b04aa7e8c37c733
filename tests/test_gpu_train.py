"""Parity of the sm_100a TRAINING step (through the C ABI of include/xvec_train.h) against the CPU oracle.
Run on a B200: ``python -m pytest tests -m gpu``.

Two kinds of checks:
  * end to end against the fp64 oracle of the reference graph (oracle/xvector_train_oracle.py): loss, pooled
    statistics, logits, moving statistics to 1e-3 / 5e-3; GRADIENTS to 2e-1 (relative L2).  The gradient of this
    network is ill conditioned in its inputs: the fp64 oracle itself moves by ~5e-2 when its frame-level
    activations are merely rounded to fp16 (test_train_oracle.py::test_fp16_storage_emulation_only_perturbs and
    tools/diag_train.py), which is the storage precision of the CUDA path -- so the end-to-end gradient check is
    a sanity bound, and
  * stage by stage, TIGHT: every backward stage is recomputed on the CPU in fp64 from the stage's own inputs as
    the GPU stored them (xv_train_debug_tensor), so the only difference left is the stage's arithmetic
    (fp16 operand products accumulated in fp32; results rounded to fp16 where they are stored).
"""
import os
import re
import sys

import numpy as np
import pytest
import torch

from oracle import xvector_oracle as orc
from oracle import xvector_train_oracle as tro
from xvector_b200 import synthetic

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def rel_max(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


class Problem:
    def __init__(self, topology, ws, B, T, NC, seed=5):
        from xvector_b200 import _native
        self.native = _native
        self.topology, self.B, self.T, self.NC = topology, B, T, NC
        self.topo = orc.TOPOLOGIES[topology]
        self.P = synthetic.make_params(self.topo["kernel_sizes"], self.topo["layer_sizes"], self.topo["embedding_sizes"],
                                       num_classes=NC, weight_set=ws, activation=self.topo.get("act", "relu"))
        self.x = synthetic.mfcc(seed, B * T).reshape(B, T, 23)
        self.labels = np.random.default_rng(seed).integers(0, NC, B).astype(np.int32)
        self.eng = _native.XvecEngine(self.topo["kernel_sizes"], self.topo["dilations"], self.topo["layer_sizes"], 512, 23, device=0,
                                      activation=self.topo.get("act", "relu"))
        if os.environ.get("XVEC_TEST_PDL") is not None:                   # diagnostics: programmatic dependent launch on / off
            self.eng.set_option("pdl", int(os.environ["XVEC_TEST_PDL"]))
        self.tr = _native.XvecTrainer(self.eng, NC, 512)
        self.tr.set_params(self.P)
        self.feats = torch.from_numpy(self.x.reshape(B * T, 23)).cuda()
        self.lab = torch.from_numpy(self.labels).cuda()
        S = 1.0
        while S < 8.0 * B * T:
            S *= 2.0
        self.S = S

    def step(self):
        la = self.tr.forward_backward(self.feats, self.lab, self.B, self.T)
        torch.cuda.synchronize()
        self.eng.check_overflow()
        return la.cpu().numpy()

    def grad(self, name):
        return self.tr.get_param(name, which=self.native.TRAIN_GRAD).reshape(np.asarray(self.P[name]).shape)

    def dbg(self, name, C=None):
        a = self.tr.debug_tensor(name)
        return a.reshape(self.B, self.T, C) if C else a.reshape(self.B, -1)

    def close(self):
        self.tr.close()
        self.eng.close()


@pytest.fixture(scope="module", params=[("ModelWithoutDropoutTdnn", "B", 12, 77), ("ModelWithoutDropout", "A", 8, 64)],
                ids=["tdnn-B-12x77", "dense-A-8x64"])
def prob(request):
    topology, ws, B, T = request.param
    p = Problem(topology, ws, B, T, 300)
    p.la = p.step()
    p.ref = tro.forward_backward(p.x, p.labels, p.P, topology, return_intermediates=True)
    yield p
    p.close()


def test_forward_against_the_fp64_oracle(prob):
    ref = prob.ref
    inter = ref["intermediates"]
    m = {"loss": (abs(prob.la[0] - ref["loss"]) / ref["loss"], 1e-3)}
    for i in range(5):
        C = prob.topo["layer_sizes"][i]
        m["r%d" % i] = (rel_l2(prob.dbg("r%d" % i, C), inter["frame_level_info_layer-%d/relu" % i]), 3e-3)
        if i < 4:
            m["y%d" % i] = (rel_l2(prob.dbg("y%d" % i, C), inter["frame_level_info_layer-%d/bn" % i]), 3e-3)
    m["h0"] = (rel_l2(prob.dbg("h0"), inter["stats"]), 1e-3)
    m["z5"] = (rel_l2(prob.dbg("z5"), inter["embed_layer-0/scores"]), 1e-3)
    m["logits"] = (rel_l2(prob.dbg("logits"), ref["logits"]), 2e-2)      # two small-batch BatchNorms amplify (see module doc)
    for s in ref["batch_stats"]:
        for leaf in ("mean:0", "variance:0"):
            m[s + leaf] = (rel_l2(prob.tr.get_param(s + leaf), ref["moving"][s + leaf]), 1e-3 if s.startswith("frame") else 5e-3)
    bad = {k: v for k, v in m.items() if not v[0] <= v[1]}
    assert not bad, bad
    assert prob.la[1] == pytest.approx(ref["accuracy"], abs=1e-6)


def test_gradients_against_the_fp64_oracle_sanity_bound(prob):
    worst = 0.0
    for name in tro.trainable_names(prob.topo, prob.P):
        worst = max(worst, rel_l2(prob.grad(name), prob.ref["grads"][name]))
    print("worst gradient error vs fp64 oracle: %.3e" % worst)
    assert worst <= 2e-1


def _segment_level_cpu(prob, h0):
    """fp64 autograd of everything above the pooled statistics, starting from the GPU's own h0.  The ReLU masks are
    the GPU's (sign of its fp32 pre-activations): a pre-activation within rounding of zero (|z| ~ 1e-7 happens) would
    otherwise flip one unit of a 12-row minibatch between fp32 and fp64 and move every gradient below it by percents."""
    masks = [torch.tensor(prob.dbg("z5") > 0, dtype=torch.float64), torch.tensor(prob.dbg("z6") > 0, dtype=torch.float64)]
    P = prob.P
    names = [n for n in tro.trainable_names(prob.topo, P) if n.startswith(("embed_layer", "output"))]
    p = {n: torch.tensor(np.asarray(P[n]), dtype=torch.float64, requires_grad=True) for n in names}
    h = torch.tensor(h0, dtype=torch.float64, requires_grad=True)
    z = h
    for i in range(2):
        s = "embed_layer-%d/" % i
        z, _, _ = tro._bn_train((z @ p[s + "w:0"] + p[s + "b:0"]) * masks[i], p[s + "gamma:0"], p[s + "beta:0"], (0,))
    logits = z @ p["output/w:0"] + p["output/b:0"]
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor(prob.labels, dtype=torch.long))
    loss.backward()
    return float(loss.detach()), h.grad.numpy(), {n: p[n].grad.numpy() for n in names}


def test_stage_segment_level_backward_is_tight(prob):
    loss, dh0, grads = _segment_level_cpu(prob, prob.dbg("h0"))
    assert abs(prob.la[0] - loss) / loss <= 2e-5
    assert rel_l2(prob.dbg("dh0"), dh0) <= 2e-4
    for n, g in grads.items():
        assert rel_l2(prob.grad(n), g) <= 2e-4, n


def test_stage_pooling_and_last_batchnorm_backward_is_tight(prob):
    C = prob.topo["layer_sizes"][4]
    s = "frame_level_info_layer-4/"
    r4 = torch.tensor(prob.dbg("r4", C), dtype=torch.float64, requires_grad=True)
    gamma = torch.tensor(np.asarray(prob.P[s + "gamma:0"]), dtype=torch.float64, requires_grad=True)
    beta = torch.tensor(np.asarray(prob.P[s + "beta:0"]), dtype=torch.float64, requires_grad=True)
    y, _, _ = tro._bn_train(r4, gamma, beta, (0, 1))
    m = y.mean(dim=1)
    v = ((y - m[:, None, :]) ** 2).mean(dim=1)
    h0 = torch.cat([m, torch.sqrt(v + 1e-5)], dim=1)
    assert rel_l2(prob.dbg("h0"), h0.detach().numpy()) <= 1e-5          # pooled statistics from the GPU's own r4
    h0.backward(torch.tensor(prob.dbg("dh0"), dtype=torch.float64))
    want = (r4.grad * (r4.detach() > 0)).numpy()
    got = prob.dbg("dz4", C) / prob.S
    assert rel_l2(got, want) <= 1e-3                                     # dz4 is stored in fp16 (2^-11 rounding)
    assert rel_l2(prob.grad(s + "gamma:0"), gamma.grad.numpy()) <= 1e-4
    assert rel_l2(prob.grad(s + "beta:0"), beta.grad.numpy()) <= 1e-4
    assert rel_l2(prob.grad(s + "b:0"), want.reshape(-1, C).sum(0)) <= 2e-4   # bias gradient = column sums of dz (before fp16 rounding)


@pytest.mark.parametrize("i", [3, 2, 1, 0])
def test_stage_batchnorm_relu_backward_is_tight(prob, i):
    C = prob.topo["layer_sizes"][i]
    s = "frame_level_info_layer-%d/" % i
    r = torch.tensor(prob.dbg("r%d" % i, C), dtype=torch.float64, requires_grad=True)
    gamma = torch.tensor(np.asarray(prob.P[s + "gamma:0"]), dtype=torch.float64, requires_grad=True)
    beta = torch.tensor(np.asarray(prob.P[s + "beta:0"]), dtype=torch.float64, requires_grad=True)
    y, _, _ = tro._bn_train(r, gamma, beta, (0, 1))
    assert rel_l2(prob.dbg("y%d" % i, C), y.detach().numpy()) <= 6e-4    # BN output, stored in fp16
    dy = prob.dbg("dy%d" % i, C) / prob.S
    y.backward(torch.tensor(dy, dtype=torch.float64))
    want = (r.grad * (r.detach() > 0)).numpy()
    got = prob.dbg("dz%d" % i, C) / prob.S
    assert rel_l2(got, want) <= 1e-3
    assert rel_l2(prob.grad(s + "gamma:0"), gamma.grad.numpy()) <= 1e-4
    assert rel_l2(prob.grad(s + "beta:0"), beta.grad.numpy()) <= 1e-4
    assert rel_l2(prob.grad(s + "b:0"), want.reshape(-1, C).sum(0)) <= 2e-4


def _shift_rows(a, off):
    """out[b, t] = a[b, t + off], zero outside the segment (SAME padding)."""
    out = np.zeros_like(a)
    T = a.shape[1]
    lo, hi = max(0, -off), min(T, T - off)
    if hi > lo:
        out[:, lo:hi] = a[:, lo + off:hi + off]
    return out


@pytest.mark.parametrize("i", [4, 3, 2, 1, 0])
def test_stage_weight_gradient_is_tight(prob, i):
    """wgrad_pair_kernel (tcgen05, MN-major operands): dW[j] = sum_rows x[row + (j-h)d]^T dz[row] from the stored fp16 operands."""
    k, d = prob.topo["kernel_sizes"][i], prob.topo["dilations"][i]
    C = prob.topo["layer_sizes"][i]
    dz = prob.dbg("dz%d" % i, C).astype(np.float64) / prob.S
    got = prob.grad("frame_level_info_layer-%d/w:0" % i)
    if i == 0:
        x0 = prob.dbg("x0", 128).astype(np.float64)[:, :, :k * 23]        # spliced first-layer input
        want = np.einsum("btk,bto->ko", x0, dz).reshape(k, 23, C)
    else:
        Cin = prob.topo["layer_sizes"][i - 1]
        x = prob.dbg("y%d" % (i - 1), Cin).astype(np.float64)
        want = np.stack([np.einsum("btc,bto->co", _shift_rows(x, (j - (k - 1) // 2) * d), dz) for j in range(k)])
    assert rel_l2(got, want) <= 1e-4
    assert rel_max(got, want) <= 1e-4


@pytest.mark.parametrize("i", [4, 3, 2, 1])
def test_stage_data_gradient_is_tight(prob, i):
    """dy_{i-1}[t] = sum_j dz_i[t - (j-h)d] W_i[j]^T with W rounded to fp16 (tdnn_pair_kernel with the flipped kernel)."""
    k, d = prob.topo["kernel_sizes"][i], prob.topo["dilations"][i]
    C, Cin = prob.topo["layer_sizes"][i], prob.topo["layer_sizes"][i - 1]
    dz = prob.dbg("dz%d" % i, C).astype(np.float64)
    W = np.asarray(prob.P["frame_level_info_layer-%d/w:0" % i]).astype(np.float16).astype(np.float64)
    want = np.zeros((prob.B, prob.T, Cin))
    for j in range(k):
        want += _shift_rows(dz, -(j - (k - 1) // 2) * d) @ W[j].T
    got = prob.dbg("dy%d" % (i - 1), Cin)
    assert rel_l2(got, want) <= 6e-4                                       # stored in fp16


def test_step_is_bit_reproducible(prob):
    g1 = prob.tr.download(prob.native.TRAIN_GRAD)
    moving = {k: v for k, v in prob.P.items() if k.endswith(("mean:0", "variance:0"))}
    prob.tr.set_params(moving)
    la = prob.step()
    g2 = prob.tr.download(prob.native.TRAIN_GRAD)
    assert np.array_equal(g1, g2) and np.array_equal(la, prob.la)


def test_tap_reuse_weight_gradient_kernel_agrees():
    """Option wgrad_reuse: x slab loaded once per K chunk, taps addressed by descriptor row offsets (MN-major operands)."""
    for topology, ws in (("ModelWithoutDropoutTdnn", "B"), ("ModelWithoutDropout", "A")):
        p = Problem(topology, ws, 8, 64, 40)
        try:
            p.step()
            g0 = {i: p.grad("frame_level_info_layer-%d/w:0" % i) for i in range(5)}
            p.tr.set_params({k: v for k, v in p.P.items() if k.endswith(("mean:0", "variance:0"))})
            p.tr.set_option("wgrad_reuse", 1)
            p.step()
            for i in range(5):
                # same fp16 operands, fp32 accumulation in another order (other K-split count)
                assert rel_l2(p.grad("frame_level_info_layer-%d/w:0" % i), g0[i]) <= 2e-5, (topology, i)
        finally:
            p.close()


def test_the_step_in_two_halves_equals_the_whole_step_bit_for_bit():
    # xv_train_forward_backward_part: part 1 (forward, loss, segment-level backward) + part 2 (pooling / frame-level backward)
    # enqueue exactly the launches of the whole step, in the same order; the segment-level gradients are final after part 1
    whole, split = Problem("ModelWithoutDropoutTdnn", "B", 8, 96, 50), Problem("ModelWithoutDropoutTdnn", "B", 8, 96, 50)
    try:
        for _ in range(3):                                   # the third round runs from captured graphs
            la0 = whole.tr.forward_backward(whole.feats, whole.lab, whole.B, whole.T)
            torch.cuda.synchronize()
            g0 = whole.tr.download(whole.native.TRAIN_GRAD)
            split.tr.download(split.native.TRAIN_GRAD)
            la1 = split.tr.forward_backward(split.feats, split.lab, split.B, split.T, part=1)
            torch.cuda.synchronize()
            off, n = split.tr.seg_grad_offset, split.tr.n_params
            g_half = split.tr.download(split.native.TRAIN_GRAD)
            assert 0 < off < n and np.array_equal(g_half[off:n], g0[off:n])          # segment-level gradients: final already
            split.tr.forward_backward(split.feats, split.lab, split.B, split.T, part=2)
            torch.cuda.synchronize()
            assert np.array_equal(split.tr.download(split.native.TRAIN_GRAD), g0)
            assert np.array_equal(la0.cpu().numpy(), la1.cpu().numpy())
            assert np.array_equal(split.tr.download(split.native.TRAIN_MOVING), whole.tr.download(whole.native.TRAIN_MOVING))
            whole.tr.apply(1e-3)
            split.tr.apply(1e-3)
            torch.cuda.synchronize()
            assert np.array_equal(split.tr.download(split.native.TRAIN_PARAMS), whole.tr.download(whole.native.TRAIN_PARAMS))
    finally:
        whole.close()
        split.close()


def test_the_frame_level_backward_in_layer_slices_equals_the_whole_step_bit_for_bit():
    # XV_TRAIN_PART_FRAME + i: part 2 cut where frame layer i's gradients are final, so that a data-parallel caller can start
    # that layer's all-reduce under the backward of the layers below.  The spans tile the frame-level gradients exactly.
    whole, split = Problem("ModelWithoutDropoutTdnn", "B", 8, 96, 50), Problem("ModelWithoutDropoutTdnn", "B", 8, 96, 50)
    try:
        spans = split.tr.frame_grad_spans
        assert spans[0][0] == 0 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
        assert spans[-1][0] + spans[-1][1] == split.tr.seg_grad_offset
        for _ in range(3):                                   # the third round runs from captured graphs
            la0 = whole.tr.forward_backward(whole.feats, whole.lab, whole.B, whole.T)
            torch.cuda.synchronize()
            g0 = whole.tr.download(whole.native.TRAIN_GRAD)
            la1 = split.tr.forward_backward(split.feats, split.lab, split.B, split.T, part=1)
            for i in reversed(range(len(spans))):
                split.tr.forward_backward(split.feats, split.lab, split.B, split.T, part=split.tr.PART_FRAME + i)
                torch.cuda.synchronize()
                g = split.tr.download(split.native.TRAIN_GRAD)
                off, cnt = spans[i]
                assert np.array_equal(g[off:off + cnt], g0[off:off + cnt])             # layer i's gradients: final already
            assert np.array_equal(split.tr.download(split.native.TRAIN_GRAD), g0)      # incl. the overflow flag behind them
            assert np.array_equal(la0.cpu().numpy(), la1.cpu().numpy())
            whole.tr.apply(1e-3)
            split.tr.apply(1e-3)
            torch.cuda.synchronize()
            assert np.array_equal(split.tr.download(split.native.TRAIN_PARAMS), whole.tr.download(whole.native.TRAIN_PARAMS))
        with pytest.raises(Exception):
            split.tr.forward_backward(split.feats, split.lab, split.B, split.T, part=split.tr.PART_FRAME + len(spans))
    finally:
        whole.close()
        split.close()


def test_adam_matches_tf_formula_exactly():
    p = Problem("ModelWithoutDropoutTdnn", "B", 4, 40, 50)
    try:
        names = tro.trainable_names(p.topo, p.P)
        rng = np.random.default_rng(11)
        Pn = {k: np.asarray(v, np.float64) for k, v in p.P.items()}
        slots = tro.adam_init(Pn, names)
        for step in range(3):
            g = {n: rng.standard_normal(np.asarray(p.P[n]).shape) * 10.0 ** rng.integers(-6, 0) for n in names}
            for n in names:
                _, off, _ = p.tr.span(n)
                p.tr.upload(p.native.TRAIN_GRAD, g[n], off)
            p.tr.apply(1e-3)
            torch.cuda.synchronize()
            g32 = {n: g[n].astype(np.float32).astype(np.float64) for n in names}
            tro.adam_step(Pn, g32, slots, 1e-3)
        assert p.tr.step == 3
        for n in names:
            got = p.tr.get_param(n).reshape(Pn[n].shape)
            assert np.abs(got - Pn[n]).max() <= 2e-6 * max(1.0, np.abs(Pn[n]).max()), n
            assert rel_l2(p.tr.get_param(n, which=p.native.TRAIN_ADAM_V), slots["v"][n]) <= 1e-4
    finally:
        p.close()


def test_eval_mode_uses_moving_statistics():
    p = Problem("ModelWithoutDropoutTdnn", "B", 6, 90, 40)
    try:
        la = p.tr.evaluate(p.feats, p.lab, p.B, p.T)
        torch.cuda.synchronize()
        la = la.cpu().numpy()
        loss, acc = tro.evaluate(p.x, p.labels, p.P, p.topology)
        assert abs(la[0] - loss) / loss <= 2e-3 and la[1] == pytest.approx(acc, abs=1e-6)
        # nothing was updated
        for k in ("frame_level_info_layer-0/mean:0", "embed_layer-1/variance:0"):
            assert np.array_equal(p.tr.get_param(k), np.asarray(p.P[k], np.float32).ravel())
    finally:
        p.close()


def test_three_training_steps_track_the_oracle_and_sync_to_the_extractor():
    p = Problem("ModelWithoutDropoutTdnn", "B", 16, 64, 100)
    try:
        names = tro.trainable_names(p.topo, p.P)
        Pn = {k: np.asarray(v, np.float64) for k, v in p.P.items()}
        slots = tro.adam_init(Pn, names)
        for it in range(3):
            o = tro.forward_backward(p.x, p.labels, Pn, p.topology)
            tro.adam_step(Pn, o["grads"], slots, 1e-3)
            Pn.update(o["moving"])
            la = p.tr.forward_backward(p.feats, p.lab, p.B, p.T)
            p.tr.apply(1e-3)
            torch.cuda.synchronize()
            # the trajectories drift apart slowly (fp16 storage, see module doc): 2 % on the first two steps, 15 % after
            assert abs(float(la[0]) - o["loss"]) / o["loss"] <= (2e-2 if it < 2 else 1.5e-1), (it, float(la[0]), o["loss"])
        # the trained variables reach the extraction path: xv_forward == oracle forward with the trainer's parameters
        p.tr.sync_model()
        cur = {k: p.tr.get_param(k).reshape(np.asarray(v).shape) for k, v in p.P.items()}
        emb = p.eng.forward(p.feats[:64].contiguous(), np.array([64], np.int32))
        torch.cuda.synchronize()
        ref = orc.forward(p.x[0], cur, p.topology)
        m = orc.parity_metrics(emb.cpu().numpy(), ref[None])
        assert m["max_rel"] <= 1e-3 and m["l2_rel"] <= 1e-3, m
    finally:
        p.close()


def test_train_one_iteration_through_the_model_surface(tmp_path):
    """Model.build_model -> train_one_iteration (tar egs, fp16 minibatches of two lengths) -> eval -> resume."""
    import logging
    from types import SimpleNamespace
    from xvector_b200 import examples_io, models
    logger = logging.getLogger("test_train")
    os.environ["XVEC_SEED"] = "123"
    m0 = str(tmp_path / "model_0")
    models.ModelWithoutDropoutTdnn().build_model(60, 23, m0, logger)
    rng = np.random.default_rng(2)
    spk_means = rng.standard_normal((60, 23)) * 6.0                       # separable synthetic speakers
    mbs, labs = [], []
    for i in range(12):
        T = 48 if i % 2 == 0 else 64
        lab = rng.integers(0, 60, 16)
        mbs.append((spk_means[lab][:, None, :] + synthetic.mfcc(100 + i, 16 * T).reshape(16, T, 23)).astype(np.float32))
        labs.append(lab)
    tar = str(tmp_path / "egs.1.tar")
    examples_io.write_egs_tar(tar, mbs, labs)
    args = SimpleNamespace(learning_rate=2e-3, print_interval=4, dropout_proportion=0.0, input_dir=m0,
                           output_dir=str(tmp_path / "model_1"), random_seed=0)
    st = models.ModelWithoutDropoutTdnn().train_one_iteration(examples_io.TarFileDataLoader(tar), args, logger)
    assert st["minibatch_count"] == 12 and st["total_segments"] == 192
    assert st["losses"][-3:].mean() < 0.7 * st["losses"][:3].mean()      # it learns
    from xvector_b200 import ze_utils
    assert ze_utils.is_correct_model_dir(args.output_dir)
    with np.load(os.path.join(args.output_dir, "model.npz")) as z:
        assert "frame_level_info_layer-1/w/Adam:0" in z.files and "output/b/Adam_1:0" in z.files
        assert abs(float(z["beta1_power:0"][0]) - 0.9 ** 12) < 1e-6
        assert not np.array_equal(z["frame_level_info_layer-0/mean:0"], np.zeros(512, np.float32))
    ev = models.ModelWithoutDropoutTdnn().eval(examples_io.TarFileDataLoader(tar), args.output_dir, True, logger)
    assert np.isfinite(ev["total_loss"]) and ev["minibatch_count"] == 12
    # the eval_dnn.py command line writes the same summary to its --log-file
    from xvector_b200 import eval_dnn
    log_file = str(tmp_path / "compute_prob_valid.1.log")
    ev2 = eval_dnn.eval_dnn(eval_dnn.get_args(["--tar-file", tar, "--input-dir", args.output_dir, "--log-file", log_file]))
    assert ev2["total_loss"] == ev["total_loss"]
    assert "Overall average loss is %.4f over 192 segments." % (ev["total_loss"] / 12) in open(log_file).read()
    # resume: Adam's step counter and slots continue from the checkpoint
    args2 = SimpleNamespace(**{**vars(args), "input_dir": args.output_dir, "output_dir": str(tmp_path / "model_2")})
    models.ModelWithoutDropoutTdnn().train_one_iteration(examples_io.TarFileDataLoader(tar), args2, logger)
    with np.load(os.path.join(args2.output_dir, "model.npz")) as z:
        assert abs(float(z["beta1_power:0"][0]) - 0.9 ** 24) < 1e-6
        assert z["global_adam_step:0"].dtype == np.int64 and int(z["global_adam_step:0"][0]) == 24
    # the trained model extracts
    emb_model = models.ModelWithoutDropoutTdnn()
    emb_model.load_model(None, args2.output_dir, logger)
    eng = emb_model._get_engine(0)
    e = eng.forward(torch.from_numpy(mbs[0][0]).cuda().contiguous(), np.array([48], np.int32))
    torch.cuda.synchronize()
    assert np.isfinite(e.cpu().numpy()).all()
    eng.close()


def test_train_dnn_driver_end_to_end(tmp_path):
    """train_dnn.py over a synthetic egs directory: 3 archives, 2 epochs, jobs 1 -> 2 (model averaging), nnet-dir layout."""
    from test_train_oracle import _make_egs_dir
    from xvector_b200 import train_dnn, ze_utils
    os.environ["XVEC_SEED"] = "7"
    egs = _make_egs_dir(str(tmp_path / "egs"), num_archives=3, minibatches=5, B=16, T=48, classes=30)
    nnet = str(tmp_path / "nnet")
    args = train_dnn.get_args(["--tf-model-class", "ModelWithoutDropoutTdnn", "--dir", nnet, "--egs-dir", egs, "--num-targets", "30",
                               "--minibatch-size", "16", "--num-epochs", "2", "--num-jobs-initial", "1", "--num-jobs-final", "2",
                               "--initial-effective-lrate", "0.002", "--final-effective-lrate", "0.0005", "--cleanup", "false",
                               "--print-interval", "2"])
    num_iters = train_dnn.train(args)
    assert num_iters == (int(2 * 3) * 2) // 3 == 4
    assert open(os.path.join(nnet, "model_name.txt")).read() == "ModelWithoutDropoutTdnn"
    assert os.path.islink(os.path.join(nnet, "model_final")) and os.readlink(os.path.join(nnet, "model_final")) == "model_4"
    for it in range(5):
        assert ze_utils.is_correct_model_dir(os.path.join(nnet, "model_%d" % it))
    assert not [d for d in os.listdir(nnet) if re.match(r"model_\d+\.\d+$", d)]         # transient job dirs removed
    rep = open(os.path.join(nnet, "accuracy.report")).read().strip().split("\n")
    rows = [r.split("\t") for r in rep[1:]]
    assert len(rows) == 1 + 1 + 2 + 2                                                    # jobs per iteration: 1, 1, 2, 2
    assert float(rows[-1][2]) < 0.8 * float(rows[0][2])                                  # the loss goes down over the run
    valid = [open(os.path.join(nnet, "log", "compute_prob_valid.%d.log" % it)).read() for it in range(4)]
    assert all("Overall average loss is" in v for v in valid)                           # eval_trained_dnn, every iteration
    assert not os.path.exists(os.path.join(nnet, "log", "compute_prob_train_subset.0.log"))   # no such archive in this egs dir
    log0 = open(os.path.join(nnet, "log", "train.0.1.log")).read()
    assert "Average training loss for minibatches 1-2 is" in log0 and "Overall average objective function is" in log0
    # a second call resumes: every iteration's output exists, nothing is retrained
    before = os.path.getmtime(os.path.join(nnet, "model_4", "model.npz"))
    train_dnn.train(args)
    assert os.path.getmtime(os.path.join(nnet, "model_4", "model.npz")) == before


def test_overflowing_gradients_skip_the_update_and_a_lower_loss_scale_recovers():
    p = Problem("ModelWithoutDropoutTdnn", "B", 8, 64, 50)
    try:
        before = p.tr.download(p.native.TRAIN_PARAMS)
        assert p.tr.n_grad == p.tr.n_params + 4 and p.tr.skipped_updates(blocking=True) == 0
        p.tr.set_option("loss_scale", 2.0 ** 40)                 # absurd: every frame-level gradient overflows fp16
        p.tr.forward_backward(p.feats, p.lab, p.B, p.T)
        p.tr.apply(1e-3)
        torch.cuda.synchronize()
        assert p.tr.skipped_updates(blocking=True) == 1          # counted on the device by adam_kernel
        assert p.tr.download(p.native.TRAIN_GRAD)[p.tr.n_params] == 1.0      # the flag rides behind the gradient
        p.eng.check_overflow()                                   # no ACTIVATION overflowed: the model's own flag stays clear
        assert np.array_equal(p.tr.download(p.native.TRAIN_PARAMS), before)       # adam_kernel left everything alone
        assert not p.tr.download(p.native.TRAIN_ADAM_M).any()
        p.tr.set_option("loss_scale", 0)                         # automatic again; the step flag is reset by every step
        la = p.step()
        p.tr.apply(1e-3)
        torch.cuda.synchronize()
        assert p.tr.skipped_updates(blocking=True) == 1 and p.tr.download(p.native.TRAIN_GRAD)[p.tr.n_params] == 0.0
        assert np.isfinite(la).all() and not np.array_equal(p.tr.download(p.native.TRAIN_PARAMS), before)
    finally:
        p.close()


def test_another_ranks_overflow_flag_in_the_reduced_gradient_skips_the_update_here_too():
    # data parallel: only the gradient is all-reduced, so the overflow flag travels INSIDE it (grad[n_params]);
    # a replica whose own step was clean must skip as well when the summed flag is non-zero -- emulated here by handing
    # xv_train_apply a reduced buffer whose tail another rank has raised
    p = Problem("ModelWithoutDropoutTdnn", "B", 8, 64, 50)
    try:
        before = p.tr.download(p.native.TRAIN_PARAMS)
        grad = torch.zeros(p.tr.n_grad, dtype=torch.float32, device="cuda")
        p.tr.forward_backward(p.feats, p.lab, p.B, p.T, grad_dev=grad)
        torch.cuda.synchronize()
        assert float(grad[p.tr.n_params]) == 0.0 and bool(torch.isfinite(grad).all())
        reduced = grad.clone()
        reduced[p.tr.n_params] += 1.0                            # what the sum all-reduce leaves when one other rank overflowed
        p.tr.apply(1e-3, grad_dev=reduced, grad_scale=0.5)
        torch.cuda.synchronize()
        assert np.array_equal(p.tr.download(p.native.TRAIN_PARAMS), before) and p.tr.skipped_updates(blocking=True) == 1
        p.tr.apply(1e-3, grad_dev=grad, grad_scale=1.0)          # the clean buffer updates
        torch.cuda.synchronize()
        assert not np.array_equal(p.tr.download(p.native.TRAIN_PARAMS), before) and p.tr.skipped_updates(blocking=True) == 1
        # polling form never blocks and catches up
        p.tr.skipped_updates(); torch.cuda.synchronize()
        assert p.tr.skipped_updates() == 1
    finally:
        p.close()


def test_config5_full_size_properties():
    """BASELINE configs[4] size (64 x 400 frames, 5000 speakers): size-independent identities instead of an oracle run."""
    p = Problem("ModelWithoutDropoutTdnn", "A", 64, 400, 5000)
    try:
        la = p.step()
        assert abs(la[0] - np.log(5000.0)) < 0.5 and 0.0 <= la[1] <= 1.0          # model_0 initialisation: near-uniform softmax
        gbo, gWo = p.grad("output/b:0").astype(np.float64), p.grad("output/w:0").astype(np.float64)
        # every row of (softmax - onehot) sums to zero -> so do d output/b and every row of d output/w
        assert abs(gbo.sum()) <= 1e-5 * np.abs(gbo).sum()
        assert np.abs(gWo.sum(axis=1)).max() <= 1e-4 * np.abs(gWo).sum(axis=1).max()
        dl = p.dbg("dlogits").astype(np.float64)
        assert np.abs(dl.sum(axis=1)).max() <= 1e-6 and abs(np.abs(dl).sum() - 2.0 * (1.0 - np.exp(-la[0]))) < 0.2
        # BatchNorm output of a training step: per channel, mean beta and variance gamma^2 var/(var+eps) over the valid frames
        y3 = p.dbg("y3", 512).reshape(-1, 512).astype(np.float64)
        np.testing.assert_allclose(y3.mean(0), np.asarray(p.P["frame_level_info_layer-3/beta:0"], np.float64), atol=2e-3)
        r3 = p.dbg("r3", 512).reshape(-1, 512).astype(np.float64)
        want_var = np.asarray(p.P["frame_level_info_layer-3/gamma:0"], np.float64) ** 2 * r3.var(0) / (r3.var(0) + 1e-3)
        np.testing.assert_allclose(y3.var(0), want_var, rtol=5e-3, atol=1e-4)
        # pooled statistics: second half is a standard deviation (>= sqrt(1e-5)); moving statistics moved 5 % towards the batch
        h0 = p.dbg("h0")
        assert (h0[:, 1536:] >= np.sqrt(1e-5) * 0.999).all()
        mm = p.tr.get_param("frame_level_info_layer-3/mean:0")
        np.testing.assert_allclose(mm, 0.05 * r3.mean(0), rtol=2e-3, atol=1e-5)      # set A starts from mean 0
        g1 = p.tr.download(p.native.TRAIN_GRAD)
        assert np.isfinite(g1).all()
        p.tr.set_params({k: v for k, v in p.P.items() if k.endswith(("mean:0", "variance:0"))})
        p.step()
        assert np.array_equal(g1, p.tr.download(p.native.TRAIN_GRAD))                # bit-reproducible at full size
    finally:
        p.close()


def test_leaky_relu_l2_loss_topology_trains():
    """ModelL2LossWithoutDropoutLRelu (reference models.py:866-983): leaky_relu(0.2) everywhere + beta * L2 of the segment weights."""
    topology = "ModelL2LossWithoutDropoutLRelu"
    p = Problem(topology, "B", 12, 70, 200)
    try:
        p.tr.set_option("l2_beta", 0.0002)
        la = p.step()
        ref = tro.forward_backward(p.x, p.labels, p.P, topology, return_intermediates=True)       # l2_beta from the topology table
        assert abs(la[0] - ref["loss"]) / ref["loss"] <= 1e-3
        plain = tro.forward_backward(p.x, p.labels, p.P, topology, l2_beta=0.0)["loss"]
        assert ref["loss"] - plain > 1e-3                                                          # the L2 term is really in the loss
        inter = ref["intermediates"]
        for i in range(5):
            C = p.topo["layer_sizes"][i]
            r = p.dbg("r%d" % i, C)
            assert (r < 0).any()                                                                   # negative half kept (slope 0.2)
            assert rel_l2(r, inter["frame_level_info_layer-%d/relu" % i]) <= 3e-3
        assert rel_l2(p.dbg("h0"), inter["stats"]) <= 1e-3
        # tight, from the GPU's own stored tensors: BatchNorm + leaky backward of layer 2, and the weight gradient of layer 3
        s = "frame_level_info_layer-2/"
        r2 = torch.tensor(p.dbg("r2", 512), dtype=torch.float64, requires_grad=True)
        y, _, _ = tro._bn_train(r2, torch.tensor(np.asarray(p.P[s + "gamma:0"]), dtype=torch.float64),
                                torch.tensor(np.asarray(p.P[s + "beta:0"]), dtype=torch.float64), (0, 1))
        y.backward(torch.tensor(p.dbg("dy2", 512) / p.S, dtype=torch.float64))
        want = (r2.grad * torch.where(r2.detach() > 0, 1.0, 0.2)).numpy()
        assert rel_l2(p.dbg("dz2", 512) / p.S, want) <= 1e-3
        x = p.dbg("y2", 512).astype(np.float64)
        dz3 = p.dbg("dz3", 512).astype(np.float64) / p.S
        assert rel_l2(p.grad("frame_level_info_layer-3/w:0")[0], np.einsum("btc,bto->co", x, dz3)) <= 1e-4
        # end to end (sanity bound, see module doc) incl. the L2 gradient of the output layer
        worst = max(rel_l2(p.grad(n), ref["grads"][n]) for n in tro.trainable_names(p.topo, p.P))
        assert worst <= 2e-1, worst
        assert rel_l2(p.grad("output/b:0"), ref["grads"]["output/b:0"]) <= 1e-2
    finally:
        p.close()


def test_changing_minibatch_geometry_leaves_no_stale_state():
    """Egs archives mix minibatch lengths: a geometry run again after other (larger and smaller) geometries gives the same bits
    (workspace re-laid-out without re-allocation, device-written metadata, graph captured on the second visit)."""
    p = Problem("ModelWithoutDropoutTdnn", "B", 8, 96, 60)
    try:
        moving = {k: v for k, v in p.P.items() if k.endswith(("mean:0", "variance:0"))}
        rng = np.random.default_rng(3)

        def run(B, T):
            x = torch.from_numpy(synthetic.mfcc(40 + T, B * T)).cuda()
            lab = torch.from_numpy(rng.integers(0, 60, B).astype(np.int32)).cuda()
            p.tr.set_params(moving)
            la = p.tr.forward_backward(x, lab, B, T)
            torch.cuda.synchronize()
            p.eng.check_overflow()
            return la.cpu().numpy().copy(), p.tr.download(p.native.TRAIN_GRAD)

        la0, g0 = p.step(), p.tr.download(p.native.TRAIN_GRAD)
        for B, T in ((16, 200), (4, 33), (8, 95), (16, 200), (4, 33)):
            la, g = run(B, T)
            assert np.isfinite(la).all() and np.isfinite(g).all()
        p.tr.set_params(moving)
        for _ in range(3):                                            # plain launch, capture, replay
            p.tr.set_params(moving)
            la1 = p.step()
            assert np.array_equal(la0, la1) and np.array_equal(g0, p.tr.download(p.native.TRAIN_GRAD))
    finally:
        p.close()
