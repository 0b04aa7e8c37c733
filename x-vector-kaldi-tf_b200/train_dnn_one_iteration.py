#!/usr/bin/env python
"""Mirror of the reference's ``local/tf/train_dnn_one_iteration.py``: one training job over one egs archive.

Same flags and exit behaviour (argument table below follows reference train_dnn_one_iteration.py:29-155; the
process exits 1 with a traceback on any error, :212-223).  The tar-file route (``--tar-file egs.N.tar`` plus
``egs.N.npy`` labels, TarFileDataLoader) is the one ``train_dnn.py`` uses (train_dnn.py:262-275); the scp/ranges
route needs the Kaldi feature archives of egs preparation and is rejected here with a clear message.

Data parallel: launched under ``torchrun --nproc-per-node N`` each rank reads ``--tar-file`` with ``{rank}``
substituted (or the same archive), gradients are all-reduced over NCCL every minibatch and rank 0 writes the model.
"""
from __future__ import annotations

import argparse
import logging
import os
import pprint
import sys
import traceback

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from xvector_b200 import models, sharding, ze_utils as utils
    from xvector_b200.examples_io import TarFileDataLoader
else:
    from . import models, sharding, ze_utils as utils
    from .examples_io import TarFileDataLoader

logger = logging.getLogger('train_dnn_one_iteration')
logger.setLevel(logging.INFO)
handler = logging.StreamHandler(sys.stdout)
handler.setLevel(logging.INFO)
handler.setFormatter(logging.Formatter("%(asctime)s [%(pathname)s:%(lineno)s - %(funcName)s - %(levelname)s ] %(message)s"))
logger.addHandler(handler)


def _str_to_bool(v):
    if v == "true":
        return True
    if v == "false":
        return False
    raise ValueError(v)


# (flag, dest, type, default, choices, required) -- reference train_dnn_one_iteration.py:39-149
FLAGS = [
    ("--use-gpu", "use_gpu", str, "yes", ["yes", "no", "wait"], False),
    ("--momentum", "momentum", float, 0.0, None, False),
    ("--shuffle", "shuffle", bool, False, None, False),
    ("--max-param-change", "max_param_change", float, 2.0, None, False),
    ("--l2-regularize-factor", "max_param_change", float, 1.0, None, False),
    ("--random-seed", "random_seed", int, 0, None, False),
    ("--print-interval", "print_interval", int, 10, None, False),
    ("--verbose", "verbose", int, 0, None, False),
    ("--feature-dim", "feature_dim", int, None, None, True),
    ("--minibatch-size", "minibatch_size", int, None, None, True),
    ("--minibatch-count", "minibatch_count", int, None, None, True),
    ("--learning-rate", "learning_rate", float, -1.0, None, False),
    ("--scale", "scale", float, 1.0, None, False),
    ("--dropout-proportion", "dropout_proportion", float, 0.0, None, False),
    ("--ranges-file", "ranges_file", str, None, None, False),
    ("--scp-file", "scp_file", str, None, None, False),
    ("--tar-file", "tar_file", str, None, None, False),
    ("--sequential-loading", "sequential_loading", _str_to_bool, True, None, False),
    ("--input-dir", "input_dir", str, None, None, True),
    ("--output-dir", "output_dir", str, None, None, True),
]


def get_args(argv=None):
    parser = argparse.ArgumentParser(description="One training iteration of the x-vector DNN on one egs archive.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter, conflict_handler='resolve')
    for flag, dest, typ, default, choices, required in FLAGS:
        kw = dict(dest=dest, type=typ, default=default, required=required)
        if choices is not None:
            kw["choices"] = choices
        parser.add_argument(flag, **kw)
    print(' '.join(sys.argv))
    return process_args(parser.parse_args(argv))


def process_args(args):
    args.input_dir = args.input_dir.strip()
    if args.input_dir == '' or not os.path.exists(os.path.join(args.input_dir, 'model.meta')):
        raise Exception("This scripts expects the input model was exist in '{0}' directory.".format(args.input_dir))
    if not args.tar_file:
        raise Exception("This build reads training examples from --tar-file archives only "
                        "(the --ranges-file/--scp-file route needs the Kaldi egs feature archives).")
    args.tar_file = args.tar_file.replace("{rank}", str(sharding.dist_info()[0]))
    if not os.path.exists(args.tar_file):
        raise Exception("The specified tar file '{0}' not exist.".format(args.tar_file))
    if not os.path.exists(args.tar_file.replace('.tar', '.npy')):
        raise Exception("There is no corresponding npy label file for tar file '{0}'.".format(args.tar_file))
    if args.dropout_proportion > 1.0 or args.dropout_proportion < 0.0:
        raise Exception("The value of dropout-proportion must be in range [0 - 1].")
    return args


def train(args):
    logger.info("Arguments for the experiment\n{0}".format(pprint.pformat(vars(args))))
    if args.random_seed != 0:
        np.random.seed(args.random_seed)
    data_loader = TarFileDataLoader(args.tar_file, logger=None, queue_size=16)
    with open(os.path.join(args.input_dir, "model.meta"), "rt") as fid:
        import json
        model_class = json.load(fid).get("model_class", "Model")
    model = getattr(models, model_class, models.Model)()
    return model.train_one_iteration(data_loader, args, logger)


def main():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(utils.pick_device())
        dist.init_process_group(backend="nccl")
    args = get_args()
    try:
        train(args)
        utils.wait_for_background_commands()
    except BaseException as e:
        if not isinstance(e, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    logger.info('Starting DNN trainer to do a training iteration (train_dnn_one_iteration.py)')
    main()
