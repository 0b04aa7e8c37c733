"""Utterance-level output of the sm_100a path (xv_forward_utts / xv_submit_host_utts, include/xvec.h) and the peer-memory
result table (xv_peer_*).  The device finishes make_embedding's chunk loop (reference local/tf/models.py:398-421):
the averaged row must be BIT-identical to that loop run on the host, in the reference's float32 arithmetic, over the
chunk x-vectors the plain forward returns.  Run on a B200: ``python -m pytest tests -m gpu``.
"""
import numpy as np
import pytest

from oracle import xvector_oracle as orc
from xvector_b200 import synthetic

pytestmark = pytest.mark.gpu


def _engine():
    from xvector_b200 import _native
    t = orc.TOPOLOGIES["ModelWithoutDropoutTdnn"]
    params = synthetic.make_params(t["kernel_sizes"], t["layer_sizes"], t["embedding_sizes"], weight_set="B")
    eng = _native.XvecEngine(t["kernel_sizes"], t["dilations"], t["layer_sizes"], 512, 23, device=0)
    eng.set_params(params)
    return eng


def _reference_average(seg_emb, first_seg, seg_lens):
    """The loop of models.py:398-421, literally (python int `offset`, python float `tot_weight`, float32 x-vectors)."""
    out = []
    for u in range(len(first_seg) - 1):
        xvector_avg = 0
        tot_weight = 0.0
        for s in range(first_seg[u], first_seg[u + 1]):
            offset = int(seg_lens[s])
            tot_weight += offset
            xvector_avg += offset * seg_emb[s]
        xvector_avg /= tot_weight
        out.append(xvector_avg)
    return np.stack(out)


def _plan():
    # utterances of 1, 3, 1, 2 and 5 chunks; chunk lengths as make_embedding cuts them (full chunks + a shorter tail)
    chunks = [[140], [50, 50, 37], [25], [400, 399], [64, 64, 64, 64, 31]]
    seg_lens = np.array([n for c in chunks for n in c], np.int32)
    first = np.concatenate([[0], np.cumsum([len(c) for c in chunks])]).astype(np.int32)
    return seg_lens, first


def test_forward_utts_equals_the_reference_chunk_loop_bit_for_bit():
    import torch
    eng = _engine()
    seg_lens, first = _plan()
    feats = torch.from_numpy(synthetic.mfcc_batch(61, seg_lens)).cuda()
    seg_emb = eng.forward(feats, seg_lens)
    torch.cuda.synchronize()
    want = _reference_average(seg_emb.cpu().numpy(), first, seg_lens)
    assert want.dtype == np.float32
    n_utt = len(first) - 1
    out = torch.full((n_utt + 3, 512), float("nan"), dtype=torch.float32, device="cuda")
    dst = np.array([4, 0, 7, 2, 5], np.int64)                    # scattered destination rows
    assert eng.forward_utts(feats, seg_lens, out, utt_first_seg=first, dst_rows=dst) == n_utt
    torch.cuda.synchronize()
    eng.check_overflow()
    got = out.cpu().numpy()
    assert np.array_equal(got[dst], want)
    untouched = np.setdiff1d(np.arange(n_utt + 3), dst)
    assert np.isnan(got[untouched]).all()
    # default plan: every segment its own utterance, rows in order: (w * x) / w in float32, as the reference computes it
    out1 = torch.empty((len(seg_lens), 512), dtype=torch.float32, device="cuda")
    eng.forward_utts(feats, seg_lens, out1)
    torch.cuda.synchronize()
    single = _reference_average(seg_emb.cpu().numpy(), np.arange(len(seg_lens) + 1), seg_lens)
    assert np.array_equal(out1.cpu().numpy(), single)
    eng.close()


def test_submit_host_utts_host_rows_and_device_rows_agree():
    import torch
    eng = _engine()
    seg_lens, first = _plan()
    feats_np = synthetic.mfcc_batch(62, seg_lens)
    feats = torch.from_numpy(feats_np).pin_memory()
    n_utt = len(first) - 1
    host = torch.zeros((n_utt, 512), dtype=torch.float32).pin_memory()
    table = torch.zeros((10, 512), dtype=torch.float32, device="cuda")
    dst = np.array([9, 8, 1, 0, 3], np.int64)
    t = eng.submit_host_utts(feats, seg_lens, utt_first_seg=first, dst_rows=dst, out_dev=table, out_host=host)
    eng.collect(t)
    seg_emb = eng.extract_host(feats_np, seg_lens)
    want = _reference_average(seg_emb, first, seg_lens)
    assert np.array_equal(host.numpy(), want)
    assert np.array_equal(table.cpu().numpy()[dst], want)
    # host-only and device-only forms
    host2 = torch.zeros_like(host).pin_memory()
    eng.collect(eng.submit_host_utts(feats, seg_lens, utt_first_seg=first, out_host=host2))
    assert np.array_equal(host2.numpy(), want)
    with pytest.raises(Exception):
        eng.submit_host_utts(feats, seg_lens, utt_first_seg=first)          # no destination
    with pytest.raises(Exception):
        eng.submit_host_utts(feats, seg_lens, utt_first_seg=np.array([0, 3, 3, 12], np.int32), out_host=host2)   # empty utterance
    eng.close()


def test_peer_table_round_trip_in_one_process():
    # xv_peer_alloc'ed memory as the destination of the averaged rows, read back with xv_peer_read (the owner's view; the
    # cross-process mapping is exercised by tools/multi_gpu_extract_check.py and bench.py --gpus N on a multi-GPU box)
    import torch
    from xvector_b200 import _native
    eng = _engine()
    seg_lens, first = _plan()
    feats = torch.from_numpy(synthetic.mfcc_batch(63, seg_lens)).cuda()
    n_utt = len(first) - 1
    table = _native.PeerTable.create(0, 2 * n_utt, 512)
    assert len(table.handle) == 64
    local = torch.empty((n_utt, 512), dtype=torch.float32, device="cuda")
    eng.forward_utts(feats, seg_lens, local, utt_first_seg=first)
    eng.forward_utts(feats, seg_lens, table.data_ptr(n_utt), utt_first_seg=first)       # second half of the table
    torch.cuda.synchronize()
    got = table.read(n_utt, n_utt)
    assert np.array_equal(got, local.cpu().numpy())
    table.close()
    eng.close()


def test_device_generated_utterances_equal_the_numpy_restatement_and_run_through_the_job_loop(tmp_path):
    # xv_synth_mfcc (the workload generator of BASELINE configs[3]) against synthetic.counter_mfcc, bit for bit; then a small
    # SyntheticSource job through Model._run_extraction_job: keys in order, rows equal to a direct forward of the regenerated rows
    import io
    import time
    import torch
    from xvector_b200 import _native, ark_job, kaldi_io, models
    ids = np.array([0, 7, 123456, 999999], np.int64)
    lens = np.array([200, 333, 25, 1000], np.int32)
    out = torch.empty((int(lens.sum()), 23), dtype=torch.float32, device="cuda")
    _native.synth_mfcc(0, out, ids, lens, seed=4)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    want = np.concatenate([synthetic.counter_mfcc(4, int(i), int(n)) for i, n in zip(ids, lens)])
    assert np.array_equal(got, want)
    assert abs(float(got[:, 0].std()) - 12.0) < 0.8 and abs(float(got.mean())) < 0.3

    import os
    os.environ["XVEC_SEED"] = "5"
    model_dir = str(tmp_path / "model_0")
    models.ModelWithoutDropoutTdnn().build_model(10, 23, model_dir, None)
    model = models.Model()
    model.load_model(None, model_dir, None)
    engine = model._get_engine(0)
    src = ark_job.SyntheticSource(0, 300, 4, 40000, rank=0, world=1)
    buf = io.BytesIO()
    t0 = time.time()
    model._run_extraction_job(src, src.counts("cuda:0"), buf, engine, 0, 25, None, t0, t0)
    rows = list(kaldi_io.read_vec_flt_ark(io.BytesIO(buf.getvalue())))
    assert [k for k, _ in rows] == ["utt%07d" % i for i in range(300)]
    all_lens = synthetic.lengths_uniform(4, 300)
    for i in (0, 151, 299):
        x = torch.from_numpy(synthetic.counter_mfcc(4, i, int(all_lens[i]))).cuda()
        o = torch.empty((1, 512), dtype=torch.float32, device="cuda")
        engine.forward_utts(x, np.array([all_lens[i]], np.int32), o)
        torch.cuda.synchronize()
        assert np.array_equal(o.cpu().numpy()[0], rows[i][1])
