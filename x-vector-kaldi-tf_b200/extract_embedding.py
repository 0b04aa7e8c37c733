#!/usr/bin/env python
"""Extract x-vectors from Kaldi features: command-line twin of the reference's
``local/tf/extract_embedding.py`` (same flags, reference :50-70), so that
``local/tf/extract_xvectors.sh:75-88`` can call it unchanged.

Kept behaviours: ``ark,scp:`` / ``scp,ark:`` targets are written to ``*.tmp.*`` and renamed at
the end, the scp is rewritten to point at the final ark, and an existing final ark+scp pair
short-circuits the call (reference :94-108, :126-128, :135-148).

Multi-GPU: launch with ``python -m torch.distributed.run --nproc-per-node N ...``; every rank
evaluates its share of the utterances and rank 0 alone writes the output (the reference's
``--nj`` fan-out + ``cat`` of scp files, extract_xvectors.sh:63-95, collapsed into one job).
"""
import argparse
import logging
import os
import sys
import traceback

if __package__ in (None, ""):                       # run as a script: make the package importable
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from xvector_b200 import kaldi_io, sharding
    from xvector_b200 import ze_utils as utils
    from xvector_b200.models import Model
else:
    from . import kaldi_io, sharding
    from . import ze_utils as utils
    from .models import Model

LOG_FORMAT = "%(asctime)s [%(pathname)s:%(lineno)s - %(funcName)s - %(levelname)s ] %(message)s"   # the reference's layout


def _make_logger():
    log = logging.getLogger("extract_embedding")
    log.setLevel(logging.INFO)
    if not log.handlers:
        stream = logging.StreamHandler()
        stream.setLevel(logging.INFO)
        stream.setFormatter(logging.Formatter(LOG_FORMAT))
        log.addHandler(stream)
    return log


logger = _make_logger()

# (flag, attribute, type, default, required, help): names, types and defaults as in the reference CLI
_FLAGS = (
    ("--use-gpu", "use_gpu", str, "no", False, "Kept for compatibility ('yes'/'no'); this build always runs on the GPU."),
    ("--min-chunk-size", "min_chunk_size", int, 100, False, "Utterances (and trailing chunks) shorter than this many frames are dropped."),
    ("--chunk-size", "chunk_size", int, -1, False, "Frames per chunk; chunk x-vectors are averaged weighted by their length. "
                                                   "-1: one x-vector from all frames of the utterance."),
    ("--feature-rspecifier", "feature_rspecifier", str, None, True, "Kaldi rspecifier of the feature matrices (file, 'ark:file', or 'command |')."),
    ("--vector-wspecifier", "vector_wspecifier", str, None, True, "Kaldi wspecifier for the x-vectors (file, '| command', 'ark,scp:A,S')."),
    ("--model-dir", "model_dir", str, None, True, "Model directory holding model.meta, model.npz and done."),
    # ---- additive: the Kaldi pipe of local/tf/extract_xvectors.sh:68 on the device (include/xvec_frontend.h) ----
    ("--apply-cmvn-sliding", "apply_cmvn_sliding", str, "no", False, "'yes': --feature-rspecifier holds RAW features; sliding-window "
                                                                    "CMVN runs on the GPU (what apply-cmvn-sliding does in the recipe)."),
    ("--cmn-window", "cmn_window", int, 300, False, "apply-cmvn-sliding --cmn-window."),
    ("--norm-vars", "norm_vars", str, "false", False, "apply-cmvn-sliding --norm-vars (true/false)."),
    ("--center", "center", str, "true", False, "apply-cmvn-sliding --center (true/false)."),
    ("--min-cmn-window", "min_cmn_window", int, 100, False, "apply-cmvn-sliding --min-cmn-window (only used with --center=false)."),
    ("--vad-rspecifier", "vad_rspecifier", str, None, False, "Kaldi rspecifier of per-frame VAD decisions (e.g. 'scp:data/vad.scp'); "
                                                             "given: only voiced frames reach the network (select-voiced-frames), "
                                                             "selected on the GPU after the CMVN."),
)


def get_args(argv=None):
    parser = argparse.ArgumentParser(description="B200-native x-vector extraction from Kaldi features.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter, conflict_handler="resolve")
    for flag, dest, typ, default, required, text in _FLAGS:
        extra = dict(choices=["yes", "no"]) if dest in ("use_gpu", "apply_cmvn_sliding") else {}
        parser.add_argument(flag, dest=dest, type=typ, default=default, required=required, help=text, **extra)
    return process_args(parser.parse_args(argv))


def process_args(args):
    args.model_dir = args.model_dir.strip()
    meta = os.path.join(args.model_dir, "model.meta")
    if not args.model_dir or not os.path.exists(meta):
        raise Exception("This scripts expects the input model was exist in '{0}' directory.".format(args.model_dir))
    return args


def process_wspecifier(wspecifier):
    """(wspecifier to open, final ark, final scp): a trailing ``ark,scp:A,S`` / ``scp,ark:S,A`` token is redirected
    to ``A.tmp.ark`` / ``S.tmp.scp``; any other wspecifier is returned unchanged with (None, None)."""
    tokens = wspecifier.split()
    prefix = "".join(t + " " for t in tokens[:-1])
    last = tokens[-1]
    for order in ("ark,scp:", "scp,ark:"):
        if last.startswith(order):
            first, second = last[len(order):].split(",")
            ark, scp = (first, second) if order == "ark,scp:" else (second, first)
            if order == "ark,scp:":
                return prefix + "ark,scp:%s.tmp.ark,%s.tmp.scp" % (ark, scp), ark, scp
            return prefix + "scp,ark:%s.tmp.scp,%s.tmp.ark" % (scp, ark), ark, scp
    return wspecifier, None, None


def _init_distributed():
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend="nccl")
    else:
        dist.init_process_group(backend="gloo")


def _publish_outputs(ark, scp):
    """Temporary files -> final names; the scp is rewritten to point at the final ark and always ends with a
    newline (Kaldi rejects a last line without one).  The ``*.tmp.scp`` file stays, as in the reference."""
    if ark is not None:
        os.rename(ark + ".tmp.ark", ark)
    if scp is None:
        return
    with open(scp + ".tmp.scp", "rt") as src:
        table = src.read().replace("ark.tmp.ark", "ark")
    if table and not table.endswith("\n"):
        table += "\n"
    with open(scp + ".tmp", "wt") as dst:
        dst.write(table)
    os.rename(scp + ".tmp", scp)


def eval_dnn(args):
    wspecifier, ark, scp = process_wspecifier(args.vector_wspecifier)
    if ark is not None and scp is not None and os.path.exists(ark) and os.path.exists(scp):
        logger.info("Both output ark and scp files exist. Return from this call.")
        return
    _init_distributed()
    rank, _ = sharding.dist_info()
    want_gpu = args.use_gpu == "yes"
    extractor = Model()
    frontend = _frontend_arguments(args)
    with _open_features(args.feature_rspecifier) as features:
        try:
            if rank != 0:                               # other ranks compute their share; only rank 0 owns the output
                extractor.make_embedding(features, None, args.model_dir, args.min_chunk_size, args.chunk_size, want_gpu,
                                         logger, **frontend)
                return
            with kaldi_io.open_vector_writer(wspecifier) as vectors:
                extractor.make_embedding(features, vectors, args.model_dir, args.min_chunk_size, args.chunk_size, want_gpu,
                                         logger, **frontend)
        finally:
            if frontend.get("vad_table") is not None:
                frontend["vad_table"].close()
    # both streams are closed: a `| copy-vector ark:- ark,scp:A.tmp.ark,S.tmp.scp` writer (or the feature pipe) must have
    # exited -- successfully -- before its files are renamed; a non-zero exit fails the job here
    kaldi_io.wait_for_children()
    _publish_outputs(ark, scp)


def _kaldi_bool(text):
    return str(text).strip().lower() in ("true", "yes", "1", "t")


def _frontend_arguments(args):
    """Keyword arguments of make_embedding that put apply-cmvn-sliding / select-voiced-frames on the device."""
    want_cmvn = getattr(args, "apply_cmvn_sliding", "no") == "yes"
    vad_rspecifier = getattr(args, "vad_rspecifier", None)
    if not want_cmvn and not vad_rspecifier:
        return {}
    if vad_rspecifier and not want_cmvn:
        raise Exception("--vad-rspecifier needs --apply-cmvn-sliding=yes: the recipe selects voiced frames AFTER the sliding "
                        "CMVN (local/tf/extract_xvectors.sh:68), and both run in one pass on the device")
    from ._native import XvCmvnOpts
    out = dict(cmvn_opts=XvCmvnOpts(args.cmn_window, args.min_cmn_window, _kaldi_bool(args.center), _kaldi_bool(args.norm_vars)))
    if vad_rspecifier:
        out["vad_table"] = kaldi_io.VecTable(vad_rspecifier)
    return out


class _open_features(object):
    """``scp:`` rspecifiers are walked entry by entry (raw feats.scp, no Kaldi pipe in front); anything else is opened
    as the archive stream the reference passes to make_embedding (extract_embedding.py:121-125)."""

    def __init__(self, rspecifier):
        self.rspecifier = rspecifier
        self.obj = None

    def __enter__(self):
        if self.rspecifier.startswith("scp") and ":" in self.rspecifier.split()[0] and not self.rspecifier.rstrip().endswith("|"):
            self.obj = kaldi_io.read_mat_scp_entries(self.rspecifier)
        else:
            self.obj = kaldi_io.open_or_fd(self.rspecifier)
        return self.obj

    def __exit__(self, *exc):
        self.obj.close()
        return False


def main(argv=None):
    args = get_args(argv)
    logger.info("Start running on host: %s" % str(os.uname()[1]))
    logger.info("Extract embeddings from features (extract_embedding.py)")
    try:
        eval_dnn(args)
        utils.wait_for_background_commands()
    except BaseException as err:
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
