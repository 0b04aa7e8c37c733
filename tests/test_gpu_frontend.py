"""Parity of the on-device feature front end (csrc/frontend.cuh through xv_frontend / xv_submit_host_raw) against the
CPU restatement of Kaldi's apply-cmvn-sliding | select-voiced-frames (oracle/kaldi_frontend_oracle.py).  Run on a B200:
``python -m pytest tests -m gpu``.

Tolerance: both sides compute in double and narrow to float.  Window sums of floats are exact in double whatever
the order, and the kernel rounds the product and the sum of ``x + (-1/N) * sum`` separately like the oracle (an FMA
there flips ~0.1 % of the outputs by one float ulp: x - mean lands exactly on float rounding midpoints that often), so
mean normalisation must be BIT-IDENTICAL apart from a stray element; with variance normalisation (sums of squares are
not exact, pow(v, -0.5) vs 1/sqrt(v)) outputs agree to one float ulp and > 99.9 % are bit-identical.
"""
import numpy as np
import pytest

from oracle import kaldi_frontend_oracle as fe
from oracle import xvector_oracle as orc
from xvector_b200 import synthetic

pytestmark = pytest.mark.gpu


def _engine(feat_dim=23, with_params=False):
    from xvector_b200 import _native
    t = orc.TOPOLOGIES["ModelWithoutDropoutTdnn"]
    eng = _native.XvecEngine(t["kernel_sizes"], t["dilations"], t["layer_sizes"], 512, feat_dim, device=0)
    params = None
    if with_params:
        params = synthetic.make_params(t["kernel_sizes"], t["layer_sizes"], t["embedding_sizes"], weight_set="B")
        eng.set_params(params)
    return eng, params


def _corpus(seed, lens, dim=23, offset=-35.0):
    rng = np.random.default_rng(seed)
    scale = 12.0 / np.sqrt(1.0 + np.arange(dim))
    feats = [(rng.standard_normal((n, dim)) * scale + offset * (np.arange(dim) == 0)).astype(np.float32) for n in lens]
    vads = [fe.synthetic_vad(rng, n) for n in lens]
    return feats, vads


def _device_frontend(eng, feats, vads, keep, opts):
    import torch
    lens = np.array([f.shape[0] for f in feats], np.int32)
    x = torch.from_numpy(np.concatenate(feats)).cuda()
    v = None if vads is None else torch.from_numpy(np.concatenate(vads)).cuda()
    out = eng.frontend(x, v, lens, keep, opts)
    torch.cuda.synchronize()
    eng.check_overflow()
    return out.cpu().numpy()


def _assert_same(got, want, ref_mag, identical=0.99999):
    assert got.shape == want.shape and got.dtype == np.float32
    ulp = np.spacing(np.maximum(np.abs(want), ref_mag).astype(np.float32))
    assert np.all(np.abs(got - want) <= ulp), float(np.abs(got - want).max())
    assert (got == want).mean() >= identical, float((got != want).mean())


@pytest.mark.parametrize("center,norm_vars,window", [(True, False, 300), (True, True, 300), (False, False, 300),
                                                     (False, True, 64), (True, False, 7), (True, False, 1000)])
def test_cmvn_and_selection_match_the_oracle(center, norm_vars, window):
    from xvector_b200._native import XvCmvnOpts
    eng, _ = _engine()
    lens = [1, 5, 127, 128, 129, 0, 300, 301, 1000, 16, 2500, 449, 150, 151]
    feats, vads = _corpus(7, lens)
    vads[2][:] = 0.0                                   # an utterance without voiced frames: nothing written
    vads[3][:] = 1.0                                   # all voiced
    opts = XvCmvnOpts(window, min(100, window), center, norm_vars)
    keep = np.array([int(np.count_nonzero(v)) for v in vads], np.int32)
    keep[8] -= 37                                      # a tail the chunking dropped
    got = _device_frontend(eng, feats, vads, keep, opts)
    want_rows, mags = [], []
    for f, v, k in zip(feats, vads, keep):
        if f.shape[0] == 0:
            continue
        cm = fe.sliding_window_cmn(f, window, center, norm_vars, min(100, window))
        sel = cm[v != 0][:k]
        want_rows.append(sel)
        mags.append(np.abs(f[v != 0][:k]) if not norm_vars else np.ones_like(sel))
    identical = 0.999 if norm_vars else 0.99999
    _assert_same(got, np.concatenate(want_rows), np.concatenate(mags), identical)
    # apply-cmvn-sliding alone (no VAD track): every row kept
    got = _device_frontend(eng, feats, None, None, opts)
    want = np.concatenate([fe.sliding_window_cmn(f, window, center, norm_vars, min(100, window)) for f in feats if f.shape[0]])
    mag = np.concatenate([np.abs(f) for f in feats if f.shape[0]]) if not norm_vars else np.ones_like(want)
    _assert_same(got, want, mag, identical)
    eng.close()


def test_long_utterances_take_the_tile_count_pass():
    """Above 4096 frames the voiced rows in front of a tile come from vad_tile_count_kernel instead of being counted
    in place: same answer, one more launch."""
    from xvector_b200._native import XvCmvnOpts
    eng, _ = _engine()
    feats, vads = _corpus(13, [5000, 300, 4097, 77])
    keep = np.array([int(np.count_nonzero(v)) for v in vads], np.int32)
    keep[0] -= 100
    got = _device_frontend(eng, feats, vads, keep, XvCmvnOpts())
    want = np.concatenate([fe.frontend(f, v)[:k] for f, v, k in zip(feats, vads, keep)])
    mag = np.concatenate([np.abs(f[v != 0][:k]) for f, v, k in zip(feats, vads, keep)])
    _assert_same(got, want, mag)
    eng.close()


def test_device_output_equals_the_c_restatement_bit_for_bit():
    """Mean normalisation + selection at the recipe's options against oracle/kaldi_frontend_oracle.c: no tolerance."""
    from oracle import kaldi_frontend_c as kfc
    eng, _ = _engine()
    feats, vads = _corpus(21, [400, 25, 999, 311, 150, 64])
    keep = np.array([int(np.count_nonzero(v)) for v in vads], np.int32)
    got = _device_frontend(eng, feats, vads, keep, None)
    want = np.concatenate([kfc.frontend(f, v) for f, v in zip(feats, vads)])
    assert got.shape == want.shape
    assert (got != want).sum() <= 2, int((got != want).sum())      # a stray last-bit flip at most (none observed)
    eng.close()


def test_other_feature_dims():
    from xvector_b200._native import XvCmvnOpts
    for dim in (13, 40):      # (wider inputs are refused by xv_create: the pack kernel stages (32 + 2*halo) * feat_dim floats)
        eng, _ = _engine(feat_dim=dim)
        feats, vads = _corpus(dim, [700, 90, 257], dim=dim)
        keep = np.array([int(np.count_nonzero(v)) for v in vads], np.int32)
        got = _device_frontend(eng, feats, vads, keep, XvCmvnOpts())
        want = np.concatenate([fe.frontend(f, v) for f, v in zip(feats, vads)])
        mag = np.concatenate([np.abs(f[v != 0]) for f, v in zip(feats, vads)])
        _assert_same(got, want, mag)
        eng.close()


def test_vad_mismatch_is_reported_not_silently_wrong():
    from xvector_b200 import _native
    eng, _ = _engine()
    feats, vads = _corpus(9, [400, 300])
    keep = np.array([int(np.count_nonzero(v)) for v in vads], np.int32)
    keep[1] = min(300, keep[1] + 5)                    # the caller claims more voiced rows than the track holds
    assert keep[1] > np.count_nonzero(vads[1])
    import torch
    lens = np.array([400, 300], np.int32)
    eng.frontend(torch.from_numpy(np.concatenate(feats)).cuda(), torch.from_numpy(np.concatenate(vads)).cuda(), lens, keep)
    with pytest.raises(_native.XvecError) as err:
        eng.check_overflow()
    assert err.value.code == _native.XV_EINVAL and "voiced" in str(err.value)
    eng.check_overflow()                               # the flag is cleared once reported
    with pytest.raises(_native.XvecError):             # host-side argument checks
        eng.frontend(torch.zeros((10, 23)).cuda(), torch.ones(10).cuda(), np.array([10], np.int32), np.array([11], np.int32))
    eng.close()


def test_raw_submission_equals_the_kaldi_pipe_then_the_network():
    """xv_submit_host_raw (raw rows + VAD in, x-vectors out) against: front-end oracle on the CPU, then the same
    engine's xv_submit_host -- and against the fp64 network oracle at the north-star tolerance."""
    import torch
    eng, params = _engine(with_params=True)
    lens = [523, 200, 1000, 64, 777]
    feats, vads = _corpus(11, lens, offset=-20.0)
    voiced = [int(np.count_nonzero(v)) for v in vads]
    keep = np.array(voiced, np.int32)
    keep[2] -= 11
    seg_lens = []
    for k in keep:                                     # two chunks for the longer utterances
        seg_lens += [int(k)] if k < 300 else [int(k) // 2, int(k) - int(k) // 2]
    raw = torch.from_numpy(np.concatenate(feats)).pin_memory()
    vad = torch.from_numpy(np.concatenate(vads)).pin_memory()
    emb = torch.empty((len(seg_lens), 512)).pin_memory()
    eng.collect(eng.submit_host_raw(raw, vad, lens, keep, seg_lens, emb))
    got = emb.numpy().copy()
    assert eng.last_launch_count >= 8                  # the front-end kernel + the 7 of a forward (first layer spliced in-kernel, last two layers fused)

    piped = np.concatenate([fe.frontend(f, v)[:k] for f, v, k in zip(feats, vads, keep)])
    ref_emb = eng.extract_host(piped, np.asarray(seg_lens, np.int32))
    m = orc.parity_metrics(got, ref_emb)
    assert m["max_rel"] <= 1e-5 and m["l2_rel"] <= 1e-5, m            # same kernels, inputs equal up to rare last-bit flips
    off, want = 0, []
    for n in seg_lens:
        want.append(orc.forward(piped[off:off + n], params, "ModelWithoutDropoutTdnn"))
        off += n
    m = orc.parity_metrics(got, np.stack(want))
    assert m["max_rel"] <= 1e-3 and m["l2_rel"] <= 1e-3, m
    eng.close()


def test_full_size_properties():
    """configs[1] size (256 utterances x 400 raw frames): shift invariance (adding a constant to every cepstral bin
    changes nothing: the window mean absorbs it), and zero column mean over utterances shorter than the window."""
    import torch
    from xvector_b200._native import XvCmvnOpts
    eng, _ = _engine()
    lens = np.full(256, 400, np.int32)
    x = synthetic.mfcc_batch(2, lens)
    xd = torch.from_numpy(x).cuda()
    opts = XvCmvnOpts(300, 100, True, False)
    a = eng.frontend(xd, None, lens, None, opts).clone()
    b = eng.frontend(xd + 64.0, None, lens, None, opts)               # 64 + x is exact for |x| < 64 up to input rounding
    torch.cuda.synchronize()
    shifted_in = (xd + 64.0) - 64.0                                    # what the shifted input really holds
    a2 = eng.frontend(shifted_in.contiguous(), None, lens, None, opts)
    torch.cuda.synchronize()
    assert float((a2 - b).abs().max()) <= 1e-5                         # constant shift: same output
    assert float((a - a2).abs().max()) <= 1e-4                         # and the rounding of the shift itself is tiny
    short = np.full(64, 250, np.int32)                                 # T < window: global mean subtraction
    xs = torch.from_numpy(synthetic.mfcc_batch(3, short)).cuda()
    y = eng.frontend(xs, None, short, None, opts).reshape(64, 250, 23)
    torch.cuda.synchronize()
    assert float(y.double().mean(dim=1).abs().max()) <= 1e-5
    eng.check_overflow()
    eng.close()
