"""CPU tests of the training oracle and of the host-side training plumbing (no GPU)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import xvector_oracle as orc                      # noqa: E402
from oracle import xvector_train_oracle as tro               # noqa: E402
from xvector_b200 import examples_io, synthetic              # noqa: E402

TOPO = "ModelWithoutDropoutTdnn"


def _problem(B=4, T=40, NC=30, ws="B", seed=7):
    topo = orc.TOPOLOGIES[TOPO]
    P = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], num_classes=NC, weight_set=ws)
    x = synthetic.mfcc(seed, B * T).reshape(B, T, 23)
    labels = np.random.default_rng(seed).integers(0, NC, B)
    return P, x, labels


def test_autograd_gradients_match_finite_differences():
    P, x, labels = _problem()
    out = tro.forward_backward(x, labels, P, TOPO)
    assert abs(out["loss"] - tro.loss_only(x, labels, P, TOPO)) < 1e-12
    rng = np.random.default_rng(0)
    for name in ["frame_level_info_layer-0/w:0", "frame_level_info_layer-2/w:0", "frame_level_info_layer-1/gamma:0",
                 "frame_level_info_layer-3/b:0", "embed_layer-0/w:0", "embed_layer-1/beta:0", "output/w:0"]:
        g = out["grads"][name]
        flat = np.argsort(-np.abs(g).ravel())[:50]
        idx = np.unravel_index(int(rng.choice(flat)), g.shape)      # a coordinate with a sizeable gradient
        eps = 1e-5
        Pp = {k: np.array(v, np.float64) for k, v in P.items()}
        Pm = {k: np.array(v, np.float64) for k, v in P.items()}
        Pp[name][idx] += eps
        Pm[name][idx] -= eps
        fd = (tro.loss_only(x, labels, Pp, TOPO) - tro.loss_only(x, labels, Pm, TOPO)) / (2 * eps)
        assert abs(fd - g[idx]) <= 1e-5 * max(1.0, abs(g[idx]) * 100), (name, idx, fd, g[idx])


def test_moving_statistics_update_and_population_variance():
    P, x, labels = _problem(ws="B")
    out = tro.forward_backward(x, labels, P, TOPO, return_intermediates=True)
    r0 = out["intermediates"]["frame_level_info_layer-0/relu"]
    mean, var = out["batch_stats"]["frame_level_info_layer-0/"]
    np.testing.assert_allclose(mean, r0.reshape(-1, r0.shape[-1]).mean(0), rtol=1e-12)
    np.testing.assert_allclose(var, r0.reshape(-1, r0.shape[-1]).var(0), rtol=1e-10)     # ddof = 0
    want = 0.95 * np.asarray(P["frame_level_info_layer-0/variance:0"], np.float64) + 0.05 * var
    np.testing.assert_allclose(out["moving"]["frame_level_info_layer-0/variance:0"], want, rtol=1e-12)
    y0 = out["intermediates"]["frame_level_info_layer-0/bn"].reshape(-1, r0.shape[-1])
    g = np.asarray(P["frame_level_info_layer-0/gamma:0"], np.float64)
    b = np.asarray(P["frame_level_info_layer-0/beta:0"], np.float64)
    np.testing.assert_allclose(y0.mean(0), b, atol=1e-9)                                   # BN output: mean beta ...
    np.testing.assert_allclose(y0.var(0), g ** 2 * var / (var + 1e-3), rtol=1e-9)         # ... variance gamma^2 var/(var+eps)


def test_eval_mode_matches_extraction_oracle_and_train_mode_when_stats_agree():
    P, x, labels = _problem(B=3, T=30)
    loss, acc = tro.evaluate(x, labels, P, TOPO)
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
    # substitute the batch statistics for the moving ones: eval forward == training forward (frame level, one segment batch)
    out = tro.forward_backward(x, labels, P, TOPO, return_intermediates=True)
    Q = dict(P)
    for s, (mean, var) in out["batch_stats"].items():
        Q[s + "mean:0"], Q[s + "variance:0"] = mean, var
    loss2, _ = tro.evaluate(x, labels, Q, TOPO)
    assert abs(loss2 - out["loss"]) < 1e-9


def test_adam_known_answer():
    p = {"w": np.array([1.0, -2.0])}
    slots = tro.adam_init(p, ["w"])
    g = {"w": np.array([0.5, -0.25])}
    tro.adam_step(p, g, slots, 0.1)
    # t = 1: m = 0.1 g, v = 0.001 g^2, lr_t = 0.1*sqrt(0.001)/0.1 -> step = lr_t * m/(sqrt(v)+eps) ~= 0.1*sign(g)
    np.testing.assert_allclose(p["w"], [1.0 - 0.1, -2.0 + 0.1], atol=1e-6)
    p1 = p["w"].copy()
    tro.adam_step(p, g, slots, 0.1)
    assert slots["t"] == 2
    m = 0.9 * 0.1 * g["w"] + 0.1 * g["w"]
    v = 0.999 * 0.001 * g["w"] ** 2 + 0.001 * g["w"] ** 2
    lr_t = 0.1 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    np.testing.assert_allclose(p["w"], p1 - lr_t * m / (np.sqrt(v) + 1e-8), rtol=1e-12)


def test_fp16_storage_emulation_only_perturbs():
    P, x, labels = _problem()
    a = tro.forward_backward(x, labels, P, TOPO)
    b = tro.forward_backward(x, labels, P, TOPO, fp16_storage=True)
    assert abs(a["loss"] - b["loss"]) / a["loss"] < 1e-3
    e = np.linalg.norm(a["grads"]["output/w:0"] - b["grads"]["output/w:0"]) / np.linalg.norm(a["grads"]["output/w:0"])
    assert 0 < e < 5e-2


def test_tar_loader_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    mbs = [rng.standard_normal((4, 20 + 5 * i, 23)).astype(np.float32) for i in range(3)]
    labs = [rng.integers(0, 50, 4) for _ in range(3)]
    tar = str(tmp_path / "egs.1.tar")
    examples_io.write_egs_tar(tar, mbs, labs)
    dl = examples_io.TarFileDataLoader(tar, queue_size=2)
    assert dl.count == 3
    for i in range(3):
        mat, lab = dl.pop(timeout=10)
        assert mat.dtype == np.float16 and mat.shape == mbs[i].shape          # stored as float16 (examples_io.py:165)
        np.testing.assert_array_equal(mat, mbs[i].astype(np.float16))
        np.testing.assert_array_equal(lab, labs[i])
    al = examples_io.ArrayDataLoader(mbs, labs)
    m, l = al.pop()
    assert m is mbs[2] and al.count == 3                                       # list.pop(): from the end
    al.pop(); al.pop()
    assert al.pop() == (None, None)


def test_one_iteration_cli_rejects_bad_arguments(tmp_path):
    from xvector_b200 import train_dnn_one_iteration as cli
    base = ["--feature-dim", "23", "--minibatch-size", "4", "--minibatch-count", "2", "--output-dir", str(tmp_path / "out")]
    with pytest.raises(Exception, match="expects the input model"):
        cli.get_args(base + ["--input-dir", str(tmp_path / "nope"), "--tar-file", "x.tar"])
    mdir = tmp_path / "model_0"
    mdir.mkdir()
    (mdir / "model.meta").write_text("{}")
    with pytest.raises(Exception, match="tar file"):
        cli.get_args(base + ["--input-dir", str(mdir), "--tar-file", str(tmp_path / "missing.tar")])
    with pytest.raises(Exception, match="--tar-file archives only"):
        cli.get_args(base + ["--input-dir", str(mdir)])
    tar = str(tmp_path / "egs.1.tar")
    examples_io.write_egs_tar(tar, [np.zeros((2, 5, 23))], [np.zeros(2, np.int32)])
    with pytest.raises(Exception, match="dropout-proportion"):
        cli.get_args(base + ["--input-dir", str(mdir), "--tar-file", tar, "--dropout-proportion", "1.5"])
    args = cli.get_args(base + ["--input-dir", str(mdir), "--tar-file", tar, "--learning-rate", "0.001"])
    assert args.learning_rate == 0.001 and args.print_interval == 10 and args.sequential_loading is True


def _make_egs_dir(root, num_archives=3, minibatches=4, B=8, T=40, classes=20, seed=0):
    rng = np.random.default_rng(seed)
    means = rng.standard_normal((classes, 23)) * 6.0
    os.makedirs(os.path.join(root, "info"))
    os.makedirs(os.path.join(root, "temp"))
    open(os.path.join(root, "info", "feat_dim"), "w").write("23\n")
    open(os.path.join(root, "info", "num_archives"), "w").write("%d\n" % num_archives)
    with open(os.path.join(root, "temp", "archive_minibatch_count"), "w") as f:
        for a in range(1, num_archives + 1):
            f.write("%d %d\n" % (a, minibatches))
            labs = [rng.integers(0, classes, B) for _ in range(minibatches)]
            mbs = [(means[l][:, None, :] + synthetic.mfcc(1000 * a + i, B * T).reshape(B, T, 23)).astype(np.float32)
                   for i, l in enumerate(labs)]
            examples_io.write_egs_tar(os.path.join(root, "egs.%d.tar" % a), mbs, labs)
            if a == 1:                                       # diagnostics archive (get_egs.sh writes valid_egs.1.tar)
                examples_io.write_egs_tar(os.path.join(root, "valid_egs.1.tar"), mbs[:2], labs[:2])
    return root


def test_train_dnn_schedules_and_bookkeeping(tmp_path):
    from xvector_b200 import train_dnn
    # learning-rate law (reference ze_utils.py:111-120): geometric in the archives processed, times the job count; last iter = final
    lr0, lr1 = 1e-3, 1e-4
    assert train_dnn.get_learning_rate(0, 1, 10, 0, 20, lr0, lr1) == pytest.approx(1e-3)
    assert train_dnn.get_learning_rate(3, 2, 10, 10, 20, lr0, lr1) == pytest.approx(2 * 1e-3 * np.sqrt(0.1))
    assert train_dnn.get_learning_rate(9, 3, 10, 19, 20, lr0, lr1) == pytest.approx(3 * 1e-4)
    assert train_dnn.get_successful_models([-1.0, -0.5, -2.1]) == [[1, 2], 2]          # within 1.0 of the best
    egs = _make_egs_dir(str(tmp_path / "egs"))
    assert train_dnn.verify_egs_dir(egs) == [3, 23, {1: 4, 2: 4, 3: 4}]
    with pytest.raises((IOError, ValueError)):
        train_dnn.verify_egs_dir(str(tmp_path / "missing"))
    # model averaging: element-wise mean of the jobs' arrays, Adam step counters from the first job
    for j, val in ((1, 1.0), (2, 3.0)):
        d = tmp_path / ("model_5.%d" % j)
        d.mkdir()
        np.savez(str(d / "model.npz"), **{"a/w:0": np.full((2, 3), val, np.float32), "beta1_power:0": np.float32([0.5 * j])})
        (d / "model.meta").write_text('{"format": "xvec-b200-v1"}')
    train_dnn.average_model_dirs([str(tmp_path / "model_5.1"), str(tmp_path / "model_5.2")], str(tmp_path / "model_5"))
    with np.load(str(tmp_path / "model_5" / "model.npz")) as z:
        np.testing.assert_array_equal(z["a/w:0"], np.full((2, 3), 2.0, np.float32))
        assert float(z["beta1_power:0"][0]) == 0.5
    from xvector_b200 import ze_utils
    assert ze_utils.is_correct_model_dir(str(tmp_path / "model_5"))
    # the report is parsed from the job logs' own summary line (reference models.py:290-292)
    os.makedirs(str(tmp_path / "nnet" / "log"))
    open(str(tmp_path / "nnet" / "log" / "train.0.1.log"), "w").write(
        "2026 [x.py:1 - f - INFO ] Overall average training loss is 3.2100 over 64 segments. Also, the overall "
        "average training accuracy is 0.2500.\n")
    rep = train_dnn.generate_report(str(tmp_path / "nnet"))
    assert "0\t1\t3.2100\t-3.2100\t0.2500" in rep
    with pytest.raises(Exception, match="not implemented"):
        train_dnn.get_args(["--tf-model-class", "Model", "--dir", "x", "--egs-dir", "y", "--num-targets", "5",
                            "--minibatch-size", "8", "--do-final-combination", "true"])
    a = train_dnn.get_args(["--tf-model-class", "ModelWithoutDropout", "--dir", "x", "--egs-dir", "y", "--num-targets", "5",
                            "--minibatch-size", "8", "--num-epochs", "2", "--cmd", "run.pl --long 0", "--momentum", "0.5"])
    assert a.num_epochs == 2.0 and a.cleanup is True and a.initial_effective_lrate == 0.0003


def test_eval_dnn_cli_rejects_bad_arguments(tmp_path):
    from xvector_b200 import eval_dnn
    with pytest.raises(Exception, match="expects the input model"):
        eval_dnn.get_args(["--tar-file", "x.tar", "--input-dir", str(tmp_path / "nope"), "--log-file", str(tmp_path / "l.log")])
    mdir = tmp_path / "model_3"
    mdir.mkdir()
    (mdir / "model.meta").write_text("{}")
    with pytest.raises(Exception, match="tar file"):
        eval_dnn.get_args(["--tar-file", str(tmp_path / "valid_egs.1.tar"), "--input-dir", str(mdir), "--log-file", str(tmp_path / "l.log")])


@pytest.mark.parametrize("topology", ["ModelWithoutDropoutTdnn", "ModelWithoutDropout", "ModelL2LossWithoutDropoutLRelu"])
def test_golden_minibatch(topology):
    """The oracle against its own committed outputs (tests/golden/make_golden_train.py): guards refactorings of the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden_train
    got = make_golden_train.compute(topology)
    with np.load(os.path.join(ROOT, "tests", "golden", "train_golden.npz")) as z:
        want = {k.split("|", 1)[1]: z[k] for k in z.files if k.startswith(topology + "|")}
    assert set(got) == set(want)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=1e-12, err_msg=k)
