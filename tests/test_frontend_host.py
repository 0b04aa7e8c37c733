"""CPU tests of the host side of the on-device feature front end: make_embedding with raw features + a VAD table
(skip rules of select-voiced-frames, chunking over the VOICED frames, the out_keep / segment bookkeeping handed to
xv_submit_host_raw), the scp entry reader, the VAD vector table and the additive CLI flags.  The device engine is
replaced by a stand-in that answers with the CPU oracles (tests are the one place allowed to call them)."""
import io
import logging
import os

import numpy as np
import pytest

from oracle import kaldi_frontend_oracle as fe
from oracle import xvector_oracle as orc
from xvector_b200 import extract_embedding, kaldi_io, models, synthetic
from test_host_logic import OracleEngine

TOPOLOGY = "ModelWithoutDropoutTdnn"


class OracleFrontendEngine(OracleEngine):
    """submit_host_raw answered by the front-end oracle followed by the network oracle; checks the bookkeeping the
    real xv_submit_host_raw checks (include/xvec_frontend.h)."""

    def submit_host_raw(self, feats, vad, utt_lens, out_keep, seg_lens, emb, opts=None):
        feats = np.array(feats, copy=True)
        vad = None if vad is None else np.array(vad, copy=True)
        utt_lens = np.asarray(utt_lens)
        keep = utt_lens if out_keep is None else np.asarray(out_keep)
        assert feats.shape[0] == utt_lens.sum() and (vad is None or vad.shape[0] == utt_lens.sum())
        assert int(np.sum(seg_lens)) == int(keep.sum())
        rows, off = [], 0
        for n, k in zip(utt_lens, keep):
            sel = fe.frontend(feats[off:off + n], None if vad is None else vad[off:off + n], opts.cmn_window,
                              bool(opts.center), bool(opts.normalize_variance), opts.min_window)
            assert sel is not None and sel.shape[0] >= k       # the host promised k voiced rows
            rows.append(sel[:k])
            off += n
        return self.submit_host(np.concatenate(rows), seg_lens, emb)


@pytest.fixture()
def model_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("XVEC_SEED", "11")
    d = str(tmp_path / "model_0")
    models.ModelWithoutDropoutTdnn().build_model(7, 23, d, None)
    engines = []

    def fake(meta, params, device):
        engines.append(OracleFrontendEngine(meta, params))
        return engines[-1]

    monkeypatch.setattr(models, "_create_engine", fake)
    return d, engines


def _raw_corpus():
    rng = np.random.default_rng(5)
    lens = dict(a=400, allsil=120, b=90, fewvoiced=200, c=333, novad=150, badlen=80, empty=0, d=640)
    feats = {k: (synthetic.mfcc(60 + i, n) + 7.0 if n else np.zeros((0, 23), np.float32)) for i, (k, n) in enumerate(lens.items())}
    vads = {k: fe.synthetic_vad(rng, n) for k, n in lens.items()}
    vads["allsil"][:] = 0.0
    vads["fewvoiced"][:] = 0.0
    vads["fewvoiced"][5:25] = 1.0                     # 20 voiced frames < min_chunk_size 25
    vads["badlen"] = vads["badlen"][:-3]
    del vads["novad"]
    return feats, vads


def _write(tmp_path, feats, vads):
    ark, scp = str(tmp_path / "feats.ark"), str(tmp_path / "feats.scp")
    with open(ark, "wb") as f, open(scp, "wt") as s:
        for k, m in feats.items():
            f.write((k + " ").encode())
            s.write("%s %s:%d\n" % (k, ark, f.tell()))
            kaldi_io.write_mat(f, m)
    vark, vscp = str(tmp_path / "vad.ark"), str(tmp_path / "vad.scp")
    with open(vark, "wb") as f, open(vscp, "wt") as s:
        for k in sorted(vads):
            f.write((k + " ").encode())
            s.write("%s %s:%d\n" % (k, vark, f.tell()))
            kaldi_io.write_vec_flt(f, vads[k])
    return ark, scp, vark, vscp


def test_scp_entries_and_vec_table(tmp_path):
    feats, vads = _raw_corpus()
    ark, scp, vark, vscp = _write(tmp_path, feats, vads)
    got = {}
    for e in kaldi_io.read_mat_scp_entries("scp:" + scp):
        assert (e.rows, e.cols) == feats[e.key].shape or e.rows == 0
        got[e.key] = e.read() if e.key != "b" else None          # an unread entry is skipped by the generator
    assert list(got) == list(feats)
    np.testing.assert_array_equal(got["c"], feats["c"])
    for spec in ("scp:" + vscp, "ark:" + vark, vark):
        table = kaldi_io.VecTable(spec)
        keys = sorted(feats)                                      # an ark table must be asked in sorted order
        for k in keys:
            v = table.get(k)
            if k in vads:
                np.testing.assert_array_equal(v, vads[k])
            else:
                assert v is None
        table.close()


@pytest.mark.parametrize("batch_frames", ["500", "400000"])
def test_make_embedding_with_device_frontend_matches_the_kaldi_pipe(model_dir, tmp_path, monkeypatch, caplog, batch_frames):
    d, engines = model_dir
    monkeypatch.setenv("XVEC_BATCH_FRAMES", batch_frames)
    feats, vads = _raw_corpus()
    feats = dict(sorted(feats.items()))                           # Kaldi tables are sorted; the ark VAD reader relies on it
    ark, scp, vark, vscp = _write(tmp_path, feats, vads)
    out = io.BytesIO()
    logger = logging.getLogger("test_frontend_host")
    table = kaldi_io.VecTable("ark:" + vark)
    with caplog.at_level(logging.INFO, logger="test_frontend_host"):
        models.Model().make_embedding(kaldi_io.read_mat_scp_entries("scp:" + scp), out, d, 25, 100, True, logger, vad_table=table)
    got = dict(kaldi_io.read_vec_flt_ark(io.BytesIO(out.getvalue())))
    assert list(got) == ["a", "b", "c", "d"]
    params = models.Model().get_models_weights(d)
    for k, v in got.items():
        piped = fe.frontend(feats[k], vads[k])                    # what the reference's loop would read from the pipe
        want = orc.make_embedding_one(piped, params, TOPOLOGY, 25, 100)
        assert orc.parity_metrics(v, want)["max_rel"] < 1e-5
    text = caplog.text
    assert "No features were judged as voiced for utterance 'allsil'" in text
    assert "No VAD decisions for utterance 'novad'" in text
    assert "Mismatch in number of frames for features and VAD of utterance 'badlen'" in text
    assert "Minimum chunk size of 25 is greater than the number of rows in utterance: fewvoiced" in text
    assert "Done 4 and failed 5" in text


def test_cmvn_only_frontend(model_dir, tmp_path):
    from xvector_b200._native import XvCmvnOpts
    d, engines = model_dir
    feats, _ = _raw_corpus()
    buf = io.BytesIO()
    for k in ("a", "c"):
        kaldi_io.write_mat(buf, feats[k], key=k)
    out = io.BytesIO()
    models.Model().make_embedding(io.BytesIO(buf.getvalue()), out, d, 25, -1, True, None, cmvn_opts=XvCmvnOpts(cmn_window=64))
    got = dict(kaldi_io.read_vec_flt_ark(io.BytesIO(out.getvalue())))
    params = models.Model().get_models_weights(d)
    for k in ("a", "c"):
        want = orc.make_embedding_one(fe.sliding_window_cmn(feats[k], 64), params, TOPOLOGY, 25, -1)
        assert orc.parity_metrics(got[k], want)["max_rel"] < 1e-5


def test_cli_with_device_frontend(model_dir, tmp_path):
    d, engines = model_dir
    feats, vads = _raw_corpus()
    feats = dict(sorted(feats.items()))
    ark, scp, vark, vscp = _write(tmp_path, feats, vads)
    out_ark, out_scp = str(tmp_path / "xvector.1.ark"), str(tmp_path / "xvector.1.scp")
    args = extract_embedding.get_args(["--use-gpu=yes", "--min-chunk-size=25", "--chunk-size=10000", "--model-dir=" + d,
                                       "--feature-rspecifier=scp:" + scp, "--apply-cmvn-sliding=yes", "--cmn-window=300",
                                       "--norm-vars=false", "--center=true", "--vad-rspecifier=scp,s,cs:" + vscp,
                                       "--vector-wspecifier=ark,scp:%s,%s" % (out_ark, out_scp)])
    extract_embedding.eval_dnn(args)
    got = dict(kaldi_io.read_vec_flt_scp(out_scp))
    assert list(got) == ["a", "b", "c", "d"]
    params = models.Model().get_models_weights(d)
    want = orc.make_embedding_one(fe.frontend(feats["d"], vads["d"]), params, TOPOLOGY, 25, 10000)
    assert orc.parity_metrics(got["d"], want)["max_rel"] < 1e-5
    with pytest.raises(Exception, match="apply-cmvn-sliding"):
        extract_embedding.eval_dnn(extract_embedding.get_args(
            ["--model-dir=" + d, "--feature-rspecifier=scp:" + scp, "--vad-rspecifier=scp:" + vscp,
             "--vector-wspecifier=ark,scp:%s,%s" % (str(tmp_path / "o.ark"), str(tmp_path / "o.scp"))]))


_WORKER = r"""
import io, os, sys, logging
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch.distributed as dist
from xvector_b200 import models, kaldi_io
import test_frontend_host as T
dist.init_process_group(backend="gloo")
rank = dist.get_rank()
models._create_engine = lambda meta, params, device: T.OracleFrontendEngine(meta, params)
os.environ["XVEC_BATCH_FRAMES"] = "500"
out = io.BytesIO() if rank == 0 else None
table = kaldi_io.VecTable("scp:" + %(vscp)r)
models.Model().make_embedding(kaldi_io.read_mat_scp_entries("scp:" + %(scp)r), out, %(model)r, 25, 100, True,
                              logging.getLogger("w"), vad_table=table)
if rank == 0:
    open(%(out)r, "wb").write(out.getvalue())
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gloo_raw_extraction_is_byte_identical_to_one_rank(model_dir, tmp_path, monkeypatch):
    """Utterance sharding with the front end on the device: every rank walks the whole feats.scp / vad table (so that the
    skip rules give all ranks the same numbering), stages only its own utterances, rank 0 gathers and writes."""
    import subprocess
    import sys
    d, _ = model_dir
    monkeypatch.setenv("XVEC_BATCH_FRAMES", "500")
    feats, vads = _raw_corpus()
    feats = dict(sorted(feats.items()))
    ark, scp, vark, vscp = _write(tmp_path, feats, vads)
    single = io.BytesIO()
    models.Model().make_embedding(kaldi_io.read_mat_scp_entries("scp:" + scp), single, d, 25, 100, True, None,
                                  vad_table=kaldi_io.VecTable("scp:" + vscp))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_path = str(tmp_path / "two_rank.ark")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % dict(root=root, model=d, out=out_path, scp=scp, vscp=vscp))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29741", str(script)],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, XVEC_SEED="11"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(single.getvalue()) > 4 * 512 * 4 and open(out_path, "rb").read() == single.getvalue()
