#!/usr/bin/env python
"""Mirror of the reference's ``local/tf/eval_dnn.py``: loss / accuracy of a trained model over one egs archive with
``phase=False`` (``Model.eval``, reference models.py:307-354), which ``train_dnn.py`` runs on the ``valid_egs.1.tar`` and
``train_subset_egs.1.tar`` diagnostics archives of every iteration (train_dnn.py:429-460).

Same flags (``--use-gpu --tar-file --input-dir --log-file``, reference eval_dnn.py:39-55), same checks and messages
(:71-84), the summary lines go to ``--log-file`` (the reference separates them from TensorFlow's own logging that way),
exit status 1 with a traceback on any error (:102-112).
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys
import traceback

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from xvector_b200 import models, ze_utils as utils
    from xvector_b200.examples_io import TarFileDataLoader
else:
    from . import models, ze_utils as utils
    from .examples_io import TarFileDataLoader

logger = logging.getLogger('eval_dnn')
logger.setLevel(logging.INFO)
formatter = logging.Formatter("%(asctime)s [%(pathname)s:%(lineno)s - %(funcName)s - %(levelname)s ] %(message)s")


def process_args(args):
    args.input_dir = args.input_dir.strip()
    if args.input_dir == '' or not os.path.exists(os.path.join(args.input_dir, 'model.meta')):
        raise Exception("This scripts expects the input model was exist in '{0}' directory.".format(args.input_dir))
    if args.tar_file == '' or not os.path.exists(args.tar_file):
        raise Exception("The specified tar file '{0}' not exist.".format(args.tar_file))
    if not os.path.exists(args.tar_file.replace('.tar', '.npy')):
        raise Exception("There is no corresponding npy label file for tar file '{0}'.".format(args.tar_file))
    return args


def get_args(argv=None):
    parser = argparse.ArgumentParser(description="Evaluates a trained x-vector DNN on one egs archive (loss, accuracy).",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter, conflict_handler='resolve')
    parser.add_argument("--use-gpu", type=str, dest='use_gpu', choices=["yes", "no"], default="yes")
    parser.add_argument("--tar-file", type=str, dest='tar_file', required=True)
    parser.add_argument("--input-dir", type=str, dest='input_dir', required=True)
    parser.add_argument("--log-file", type=str, dest='log_file', required=True)
    print(' '.join(sys.argv))
    args = process_args(parser.parse_args(argv))
    handler = logging.StreamHandler(open(args.log_file, 'wt'))
    handler.setLevel(logging.INFO)
    handler.setFormatter(formatter)
    logger.addHandler(handler)
    logger.info('Starting DNN evaluation (eval_dnn.py)')
    return args


def eval_dnn(args):
    use_gpu = args.use_gpu == 'yes'
    data_loader = TarFileDataLoader(args.tar_file, logger=None, queue_size=16)
    model_class = "Model"
    try:
        with open(os.path.join(args.input_dir, "model.meta"), "rt") as fid:
            model_class = json.load(fid).get("model_class", "Model")
    except (ValueError, UnicodeDecodeError):
        pass                                   # a TensorFlow checkpoint directory: the base topology, as the reference's Model()
    model = getattr(models, model_class, models.Model)()
    return model.eval(data_loader, args.input_dir, use_gpu, logger)


def main():
    args = get_args()
    try:
        eval_dnn(args)
        utils.wait_for_background_commands()
    except BaseException as e:
        if not isinstance(e, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
