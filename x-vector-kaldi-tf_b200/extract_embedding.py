#!/usr/bin/env python
"""Extract x-vectors from Kaldi features: command-line twin of the reference's
``local/tf/extract_embedding.py`` (flags verbatim, reference :50-70), so that
``local/tf/extract_xvectors.sh:75-88`` can call it unchanged.

Kept behaviours: ``ark,scp:`` / ``scp,ark:`` targets are written to ``*.tmp.*`` and renamed at
the end, the scp is rewritten to point at the final ark, and an existing final ark+scp pair
short-circuits the call (reference :94-108, :126-128, :135-148).

Multi-GPU: launch with ``python -m torch.distributed.run --nproc-per-node N ...``; every rank
evaluates its share of the utterances and rank 0 alone writes the output (the reference's
``--nj`` fan-out + ``cat`` of scp files, extract_xvectors.sh:63-95, collapsed into one job).
"""
from __future__ import print_function

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

logger = logging.getLogger('extract_embedding')
logger.setLevel(logging.INFO)
formatter = logging.Formatter("%(asctime)s [%(pathname)s:%(lineno)s - "
                              "%(funcName)s - %(levelname)s ] %(message)s")
handler = logging.StreamHandler()
handler.setLevel(logging.INFO)
handler.setFormatter(formatter)
logger.addHandler(handler)


def get_args(argv=None):
    parser = argparse.ArgumentParser(
        description="Extract x-vector embeddings from Kaldi features with the B200-native forward pass.",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        conflict_handler='resolve')
    parser.add_argument("--use-gpu", type=str, dest='use_gpu', choices=["yes", "no"],
                        help="Kept for compatibility; this build always runs on the GPU.", default="no")
    parser.add_argument("--min-chunk-size", type=int, dest='min_chunk_size', default=100,
                        help="Minimum chunk-size allowed when extracting xvectors.")
    parser.add_argument("--chunk-size", type=int, dest='chunk_size', default=-1,
                        help="If set, extracts xvectors from specified chunk-size, and averages.  "
                             "If not set, extracts an xvector from all available features.")
    parser.add_argument("--feature-rspecifier", type=str, dest='feature_rspecifier', required=True,
                        help="Kaldi rspecifier producing the feature matrices (file, 'ark:', or 'cmd |').")
    parser.add_argument("--vector-wspecifier", type=str, dest='vector_wspecifier', required=True,
                        help="Kaldi wspecifier receiving the vectors (file, '| cmd', 'ark,scp:A,S').")
    parser.add_argument("--model-dir", type=str, dest='model_dir', required=True,
                        help="Model directory (model.meta, model.npz, done).")
    args = parser.parse_args(argv)
    return process_args(args)


def process_args(args):
    args.model_dir = args.model_dir.strip()
    if args.model_dir == '' or not os.path.exists(os.path.join(args.model_dir, 'model.meta')):
        raise Exception("This scripts expects the input model was exist in '{0}' directory.".format(args.model_dir))
    return args


def process_wspecifier(wspecifier):
    """Redirect a trailing ``ark,scp:A,S`` / ``scp,ark:S,A`` token to temporary files."""
    parts = wspecifier.split()
    head = ''.join(p + ' ' for p in parts[:-1])
    last = parts[-1]
    if last.startswith('ark,scp:'):
        ark, scp = last[8:].split(',')
        return head + 'ark,scp:%s.tmp.ark,%s.tmp.scp' % (ark, scp), ark, scp
    if last.startswith('scp,ark:'):
        scp, ark = last[8:].split(',')
        return head + 'scp,ark:%s.tmp.scp,%s.tmp.ark' % (scp, ark), ark, scp
    return wspecifier, None, None


def _init_distributed():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend="nccl")
        else:
            dist.init_process_group(backend="gloo")


def eval_dnn(args):
    model_dir = args.model_dir
    use_gpu = args.use_gpu == 'yes'
    wspecifier, ark, scp = process_wspecifier(args.vector_wspecifier)
    if ark is not None and os.path.exists(ark) and scp is not None and os.path.exists(scp):
        logger.info('Both output ark and scp files exist. Return from this call.')
        return
    _init_distributed()
    rank, world = sharding.dist_info()
    model = Model()
    with kaldi_io.open_or_fd(args.feature_rspecifier) as input_fid:
        if rank == 0:
            with kaldi_io.open_vector_writer(wspecifier) as output_fid:
                model.make_embedding(input_fid, output_fid, model_dir, args.min_chunk_size, args.chunk_size,
                                     use_gpu, logger)
        else:
            model.make_embedding(input_fid, None, model_dir, args.min_chunk_size, args.chunk_size, use_gpu, logger)
    if rank != 0:
        return
    if ark is not None:
        os.rename(ark + '.tmp.ark', ark)
    if scp is not None:
        with open(scp + '.tmp.scp', 'rt') as fid_in:
            text = fid_in.read()
        text = text.replace('ark.tmp.ark', 'ark')
        if text and text[-1] != '\n':
            text += '\n'
        with open(scp + '.tmp', 'wt') as fid_out:
            fid_out.write(text)
        os.rename(scp + '.tmp', scp)


def main(argv=None):
    args = get_args(argv)
    logger.info('Start running on host: %s' % str(os.uname()[1]))
    logger.info('Extract embeddings from features (extract_embedding.py)')
    try:
        eval_dnn(args)
        utils.wait_for_background_commands()
    except BaseException as e:
        if not isinstance(e, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
