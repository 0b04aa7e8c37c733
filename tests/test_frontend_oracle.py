"""CPU tests of the feature front-end oracle (oracle/kaldi_frontend_oracle.py): sliding-window CMVN + voiced-frame
selection, the Kaldi pipe of the reference's local/tf/extract_xvectors.sh:68.  Kaldi is absent (parity unpinned): the
restatement is checked against hand-computed known answers, an independent second restatement and invariants."""
import numpy as np
import pytest

from oracle import kaldi_frontend_oracle as fe


def test_window_placement_known_answers():
    # centred window of 3 over 5 frames: shifted (not shrunk) at both edges
    got = [fe.window_bounds(t, 5, cmn_window=3, center=True) for t in range(5)]
    assert got == [(0, 3), (0, 3), (1, 4), (2, 5), (2, 5)]
    # the reference's options on a 1000-frame utterance
    assert fe.window_bounds(0, 1000) == (0, 300)
    assert fe.window_bounds(149, 1000) == (0, 300)
    assert fe.window_bounds(150, 1000) == (0, 300)
    assert fe.window_bounds(151, 1000) == (1, 301)
    assert fe.window_bounds(849, 1000) == (699, 999)
    assert fe.window_bounds(850, 1000) == (700, 1000)
    assert fe.window_bounds(999, 1000) == (700, 1000)
    # shorter than the window: every frame sees the whole utterance
    assert {fe.window_bounds(t, 200) for t in range(200)} == {(0, 200)}
    # causal window: looks ahead until min_window frames exist, then trails
    assert fe.window_bounds(0, 1000, 300, center=False, min_window=100) == (0, 100)
    assert fe.window_bounds(99, 1000, 300, center=False, min_window=100) == (0, 100)
    assert fe.window_bounds(100, 1000, 300, center=False, min_window=100) == (0, 101)
    assert fe.window_bounds(500, 1000, 300, center=False, min_window=100) == (200, 501)
    assert fe.window_bounds(10, 50, 300, center=False, min_window=100) == (0, 50)


def test_windows_slide_by_at_most_one_frame():
    for T in (1, 7, 150, 299, 300, 301, 777):
        for center in (True, False):
            prev = None
            for t in range(T):
                ws, we = fe.window_bounds(t, T, 300, center, 100)
                assert 0 <= ws <= t < we <= T
                if prev is not None:
                    assert ws - prev[0] in (0, 1) and we - prev[1] in (0, 1)
                prev = (ws, we)


def test_known_answer_small():
    x = np.array([[1.0, 10.0], [2.0, 20.0], [6.0, 0.0], [4.0, 40.0], [5.0, -10.0]], np.float32)
    want = np.array([x[0] - x[0:3].mean(0), x[1] - x[0:3].mean(0), x[2] - x[1:4].mean(0),
                     x[3] - x[2:5].mean(0), x[4] - x[2:5].mean(0)], np.float32)
    got = fe.sliding_window_cmn(x, cmn_window=3, center=True)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)


def test_short_utterance_is_global_mean_subtraction():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((120, 23)) * 5 + 3).astype(np.float32)
    got = fe.sliding_window_cmn(x)
    np.testing.assert_allclose(got, x - x.astype(np.float64).mean(0), rtol=0, atol=2e-6)


@pytest.mark.parametrize("center,norm_vars,window", [(True, False, 300), (True, True, 300), (False, False, 300),
                                                     (False, True, 64), (True, False, 7)])
def test_two_restatements_agree(center, norm_vars, window):
    rng = np.random.default_rng(1)
    for T in (1, 2, 33, 299, 300, 301, 1000):
        x = (rng.standard_normal((T, 23)) * 12 / np.sqrt(1 + np.arange(23)) - 40.0).astype(np.float32)
        a = fe.sliding_window_cmn(x, window, center, norm_vars, min_window=min(100, window))
        b = fe.sliding_window_cmn_direct(x, window, center, norm_vars, min_window=min(100, window))
        assert a.dtype == np.float32 and a.shape == x.shape
        # both are double arithmetic narrowed to float: they may differ by the last float bit at most
        np.testing.assert_allclose(a, b, rtol=3e-7, atol=1e-6)
        assert (a == b).mean() > 0.99


def test_variance_normalisation_gives_unit_variance_over_the_window():
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((900, 5)) * np.array([1, 3, 10, 0.1, 50])).astype(np.float32)
    y = fe.sliding_window_cmn(x, 300, True, True)
    t = 450
    ws, we = fe.window_bounds(t, 900)
    xw = x[ws:we].astype(np.float64)
    np.testing.assert_allclose(y[t], (x[t] - xw.mean(0)) / xw.std(0), rtol=1e-5)
    # a constant feature: variance floored at 1e-10, output 0
    c = np.full((400, 2), 3.25, np.float32)
    assert np.all(fe.sliding_window_cmn(c, 300, True, True) == 0.0)


def test_select_voiced_frames():
    x = np.arange(12, dtype=np.float32).reshape(6, 2)
    vad = np.array([0, 1, 1, 0, 0, 1], np.float32)
    np.testing.assert_array_equal(fe.select_voiced_frames(x, vad), x[[1, 2, 5]])
    assert fe.select_voiced_frames(x, np.zeros(6, np.float32)) is None           # no voiced frame: utterance skipped
    assert fe.select_voiced_frames(x, np.ones(5, np.float32)) is None            # length mismatch: skipped
    out = fe.frontend(x, vad, cmn_window=3)
    np.testing.assert_array_equal(out, fe.sliding_window_cmn(x, 3)[[1, 2, 5]])     # CMVN sees ALL frames, selection after


def test_synthetic_vad_is_binary_and_mixed():
    rng = np.random.default_rng(3)
    v = fe.synthetic_vad(rng, 5000)
    assert set(np.unique(v)) == {0.0, 1.0}
    assert 0.5 < v.mean() < 0.95


@pytest.mark.parametrize("center,norm_vars,window", [(True, False, 300), (True, True, 300), (False, False, 300),
                                                     (False, True, 64), (True, False, 7)])
def test_c_restatement_is_bit_identical_to_the_numpy_one(center, norm_vars, window):
    """oracle/kaldi_frontend_oracle.c (gcc -O2 -ffp-contract=off) follows the same recursion in the same order: the third
    restatement must agree with the first bit for bit, selection included."""
    from oracle import kaldi_frontend_c as kfc
    rng = np.random.default_rng(4)
    for T in (1, 2, 33, 299, 300, 301, 1000):
        x = (rng.standard_normal((T, 23)) * 12 / np.sqrt(1 + np.arange(23)) - 40.0 * (np.arange(23) == 0)).astype(np.float32)
        v = fe.synthetic_vad(rng, T)
        mw = min(100, window)
        assert np.array_equal(kfc.sliding_window_cmn(x, window, center, norm_vars, mw),
                              fe.sliding_window_cmn(x, window, center, norm_vars, mw))
        a, b = kfc.frontend(x, v, window, center, norm_vars, mw), fe.frontend(x, v, window, center, norm_vars, mw)
        assert (a is None and b is None) or np.array_equal(a, b)
    assert kfc.frontend(np.ones((4, 3), np.float32), np.zeros(4, np.float32)) is None
    assert kfc.frontend(np.ones((4, 3), np.float32), np.ones(5, np.float32)) is None
