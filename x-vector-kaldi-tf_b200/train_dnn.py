#!/usr/bin/env python
"""Mirror of the reference's ``local/tf/train_dnn.py``: the multi-iteration training driver.

Same flags (the table below follows reference train_dnn.py:41-180; the recipe's call is run_xvector.sh:88-107), same
nnet-dir layout (``model_<iter>/``, transient ``model_<iter+1>.<job>/``, ``model_final`` symlink, ``model_name.txt``,
``random_seed``, ``accuracy.report``, ``log/train.<iter>.<job>.log``; train_dnn.py:285,324-338,495,583,591-593) and same
schedules: ``num_iters = (int(num_epochs * num_archives) * 2) / (jobs_initial + jobs_final)`` (:504-506), the job count
interpolated per iteration (:521-523), ``lr = jobs * lr0 * exp(processed * ln(lr1/lr0) / to_process)`` with the last
iteration at ``lr1`` (ze_utils.py:111-120), archive ``k % num_archives + 1`` for job k (:249-252).

What differs: the reference spawns each job as a queue command and leaves the averaging of the jobs' models unimplemented
(``get_average_nnet_model`` is a stub, ze_utils.py:164-183, so only ``--num-jobs-*=1`` -- the recipe's setting -- works
there).  Here the jobs of an iteration run in this process, one after another (or one per rank under ``torchrun``: rank r
takes jobs r+1, r+1+W, ...), each through ``Model.train_one_iteration`` on the sm_100a training step, and the accepted
jobs' variables are averaged (the nnet3-average the reference intended).  ``--cmd``, ``--momentum``,
``--max-param-change``, ``--proportional-shrink`` and the dropout schedule are accepted for command-line compatibility;
the ModelWithoutDropout family ignores them in the reference as well (its graph has no such ops).
"""
from __future__ import annotations

import argparse
import logging
import math
import os
import pprint
import re
import shutil
import sys
import traceback
from types import SimpleNamespace

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from xvector_b200 import models, sharding, ze_utils as utils
    from xvector_b200.examples_io import TarFileDataLoader
else:
    from . import models, sharding, ze_utils as utils
    from .examples_io import TarFileDataLoader

logger = logging.getLogger('train_dnn')
logger.setLevel(logging.INFO)
if not logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setFormatter(logging.Formatter("%(asctime)s [%(pathname)s:%(lineno)s - %(funcName)s - %(levelname)s ] %(message)s"))
    logger.addHandler(_h)

# (flag, dest, type, default, required)
FLAGS = [
    ("--use-gpu", "use_gpu", str, "yes", False), ("--momentum", "momentum", float, 0.0, False),
    ("--targets-scp", "targets_scp", str, None, False), ("--tf-model-class", "tf_model_class", str, None, True),
    ("--dir", "dir", str, None, True), ("--egs-dir", "egs_dir", str, None, True),
    ("--num-epochs", "num_epochs", float, 6.0, False), ("--num-targets", "num_targets", int, None, True),
    ("--initial-effective-lrate", "initial_effective_lrate", float, 0.0003, False),
    ("--final-effective-lrate", "final_effective_lrate", float, 0.00003, False),
    ("--num-jobs-initial", "num_jobs_initial", int, 1, False), ("--num-jobs-final", "num_jobs_final", int, 8, False),
    ("--minibatch-size", "minibatch_size", int, None, True), ("--do-final-combination", "do_final_combination", str, "false", False),
    ("--random-seed", "random_seed", int, 0, False), ("--dropout-schedule", "dropout_schedule", str, None, False),
    ("--max-objective-evaluations", "max_objective_evaluations", int, 30, False),
    ("--preserve-model-interval", "preserve_model_interval", int, 10, False), ("--cleanup", "cleanup", str, "true", False),
    ("--max-param-change", "max_param_change", float, 2.0, False), ("--proportional-shrink", "proportional_shrink", float, 0.0, False),
    ("--stage", "stage", int, -4, False), ("--cmd", "command", str, "queue.pl", False),
    ("--max-models-combine", "max_models_combine", int, 20, False), ("--print-interval", "print_interval", int, 10, False),
]


def get_args(argv=None):
    parser = argparse.ArgumentParser(description="Trains the x-vector DNN over the archives of an egs directory.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter, conflict_handler='resolve')
    for flag, dest, typ, default, required in FLAGS:
        parser.add_argument(flag, dest=dest, type=typ, default=default, required=required)
    args = parser.parse_args(argv)
    args.cleanup = args.cleanup == "true"
    args.do_final_combination = args.do_final_combination == "true"
    if args.do_final_combination:
        raise Exception('combine model using average not implemented yet.')          # as the reference, train_dnn.py:575
    if not hasattr(models, args.tf_model_class):
        raise Exception("unknown --tf-model-class %s" % args.tf_model_class)
    if args.num_jobs_initial < 1 or args.num_jobs_final < args.num_jobs_initial:
        raise Exception("need 1 <= num-jobs-initial <= num-jobs-final")
    return args


def verify_egs_dir(egs_dir):
    """[num_archives, feat_dim, {archive: minibatch count}] (reference ze_utils.py:56-73)."""
    try:
        egs_feat_dim = int(open('{0}/info/feat_dim'.format(egs_dir)).readline())
        num_archives = int(open('{0}/info/num_archives'.format(egs_dir)).readline())
        archives_minibatch_count = {}
        with open('{0}/temp/archive_minibatch_count'.format(egs_dir), 'rt') as fid:
            for line in fid:
                if len(line.strip()) == 0:
                    continue
                parts = line.split()
                archives_minibatch_count[int(parts[0])] = int(parts[1])
        return [num_archives, egs_feat_dim, archives_minibatch_count]
    except (IOError, ValueError):
        logger.error("The egs dir {0} has missing or malformed files.".format(egs_dir))
        raise


def get_learning_rate(_iter, num_jobs, num_iters, num_archives_processed, num_archives_to_process,
                      initial_effective_lrate, final_effective_lrate):
    """reference ze_utils.py:111-120"""
    if _iter + 1 >= num_iters:
        effective_learning_rate = final_effective_lrate
    else:
        effective_learning_rate = (initial_effective_lrate *
                                   math.exp(num_archives_processed * math.log(final_effective_lrate / initial_effective_lrate)
                                            / num_archives_to_process))
    return num_jobs * effective_learning_rate


def get_successful_models(objectives, difference_threshold=1.0):
    """[accepted job numbers, best job number] from the jobs' overall objectives (reference ze_utils.py:123-154, which
    parses the same numbers back out of the job logs)."""
    max_index = objectives.index(max(objectives))
    accepted = [i + 1 for i, o in enumerate(objectives) if (objectives[max_index] - o) <= difference_threshold]
    if len(accepted) != len(objectives):
        logger.warning("Only {0}/{1} of the models have been accepted for averaging.".format(len(accepted), len(objectives)))
    return [accepted, max_index + 1]


COUNTER_SUFFIXES = ("_power:0", "_step:0")      # Adam's step counters: taken from the first job, never averaged


def average_model_dirs(job_dirs, out_dir):
    """Element-wise mean of every array of the jobs' ``model.npz`` (what nnet3-average does for the reference's intended
    flow, ze_utils.py:176-183); Adam's step counters (beta*_power, global_adam_step) are taken from the first job."""
    acc, meta = None, None
    for d in job_dirs:
        with np.load(os.path.join(d, "model.npz")) as z:
            cur = {k: z[k].astype(np.float64) for k in z.files}
        if acc is None:
            acc = cur
            meta = open(os.path.join(d, "model.meta"), "rt").read()
        else:
            for k in acc:
                if not k.endswith(COUNTER_SUFFIXES):
                    acc[k] += cur[k]
    n = float(len(job_dirs))
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "model.meta"), "wt") as fid:
        fid.write(meta)
    with open(os.path.join(out_dir, "model.npz"), "wb") as fid:
        np.savez(fid, **{k: (np.rint(v).astype(np.int64) if k.endswith("_step:0") else
                             (v if k.endswith(COUNTER_SUFFIXES) else v / n).astype(np.float32)) for k, v in acc.items()})
    with open(os.path.join(out_dir, "done"), "wt") as fid:
        fid.write("done")


def eval_trained_dnn(args, _iter, egs_dir):
    """Diagnostics of model_<iter> on the validation and train-subset archives (reference train_dnn.py:429-460 starts two
    eval_dnn.py background commands; here Model.eval runs in this process, same log files and summary lines)."""
    input_model_dir = "{0}/model_{1}".format(args.dir, _iter)
    for name in ("valid", "train_subset"):
        tar_file = "{0}/{1}_egs.1.tar".format(egs_dir, name)
        if not os.path.exists(tar_file):
            continue                                         # the reference's command would fail; the archive is optional here
        log_path = "{0}/log/compute_prob_{1}.{2}.log".format(args.dir, name, _iter)
        ev_logger = logging.getLogger("train_dnn.eval.%s.%d" % (name, _iter))
        ev_logger.setLevel(logging.INFO)
        ev_logger.propagate = False
        fh = logging.FileHandler(log_path, mode="w")
        fh.setFormatter(logging.Formatter("%(asctime)s [%(pathname)s:%(lineno)s - %(funcName)s - %(levelname)s ] %(message)s"))
        ev_logger.addHandler(fh)
        try:
            ev_logger.info('Starting DNN evaluation (eval_dnn.py)')
            getattr(models, args.tf_model_class)().eval(TarFileDataLoader(tar_file, logger=None, queue_size=16), input_model_dir,
                                                        True, ev_logger)
        finally:
            ev_logger.removeHandler(fh)
            fh.close()


def train_one_iteration(args, _iter, egs_dir, num_jobs, num_archives_processed, num_archives, learning_rate,
                        archives_minibatch_count):
    model_dir = args.dir
    rank, world = sharding.dist_info()
    random_seed_file = '{0}/random_seed'.format(model_dir)
    if rank == 0 and not os.path.exists(random_seed_file):
        with open(random_seed_file, 'w') as fid:
            fid.write(str(args.random_seed))
    if rank == 0:
        eval_trained_dnn(args, _iter, egs_dir)               # computes train and validation set objectives
    if utils.is_correct_model_dir("{0}/model_{1}".format(model_dir, _iter + 1)):
        logger.info('The output model {0}/model_{1}/model.meta was exist and so I do not continue this iteration.'.format(model_dir, _iter + 1))
        return
    objectives = [None] * num_jobs
    for job in range(1 + rank, num_jobs + 1, world):
        k = num_archives_processed + job - 1
        archive_index = (k % num_archives) + 1
        tar_file = '{egs_dir}/egs.{archive_index}.tar'.format(egs_dir=egs_dir, archive_index=archive_index)
        job_args = SimpleNamespace(learning_rate=learning_rate, print_interval=args.print_interval, dropout_proportion=0.0,
                                   input_dir='{0}/model_{1}'.format(model_dir, _iter),
                                   output_dir='{0}/model_{1}.{2}'.format(model_dir, _iter + 1, job),
                                   random_seed=_iter + args.random_seed, data_parallel=False)
        log_path = '{0}/log/train.{1}.{2}.log'.format(model_dir, _iter, job)
        job_logger = logging.getLogger('train_dnn.job.%d.%d' % (_iter, job))
        job_logger.setLevel(logging.INFO)
        job_logger.propagate = False
        fh = logging.FileHandler(log_path, mode="w")
        fh.setFormatter(logging.Formatter("%(asctime)s [%(pathname)s:%(lineno)s - %(funcName)s - %(levelname)s ] %(message)s"))
        job_logger.addHandler(fh)
        try:
            model = getattr(models, args.tf_model_class)()
            st = model.train_one_iteration(TarFileDataLoader(tar_file, logger=None, queue_size=16), job_args, job_logger)
            assert st["minibatch_count"] == archives_minibatch_count.get(archive_index, st["minibatch_count"])
            objectives[job - 1] = -st["total_loss"] / max(st["minibatch_count"], 1)
        finally:
            job_logger.removeHandler(fh)
            fh.close()
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, objectives)
        objectives = [next(o[j] for o in gathered if o[j] is not None) for j in range(num_jobs)]
    if rank == 0:
        accepted, best = get_successful_models(objectives)
        out_dir = "{0}/model_{1}".format(model_dir, _iter + 1)
        if _iter > 0 and len(accepted) > 1:
            average_model_dirs(["{0}/model_{1}.{2}".format(model_dir, _iter + 1, n) for n in accepted], out_dir)
        else:
            shutil.copytree("{0}/model_{1}.{2}".format(model_dir, _iter + 1, best), out_dir)     # copy_best_nnet_dir
        for i in range(1, num_jobs + 1):
            shutil.rmtree("{0}/model_{1}.{2}".format(model_dir, _iter + 1, i))
        if not utils.is_correct_model_dir(out_dir):
            raise Exception("Could not find {0}/model.meta, at the end of iteration {1}".format(out_dir, _iter))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def remove_model(nnet_dir, _iter, preserve_model_interval=100):
    if _iter < 0 or _iter % preserve_model_interval == 0:
        return
    model_dir = '{0}/model_{1}'.format(nnet_dir, _iter)
    if os.path.exists(model_dir):
        shutil.rmtree(model_dir)


def generate_report(nnet_dir):
    """iteration, objective, accuracy per iteration from the job logs (the lines the reference parses, ze_utils.py:498-499)."""
    rx = re.compile(r"Overall average training loss is (-?[0-9]+(?:\.[0-9]+)?) over ([0-9]+) segments\. Also, the overall average "
                    r"training accuracy is ([0-9]+(?:\.[0-9]+)?)")
    rows = []
    for name in sorted(os.listdir(os.path.join(nnet_dir, "log"))):
        m = re.match(r"train\.(\d+)\.(\d+)\.log$", name)
        if not m:
            continue
        for line in open(os.path.join(nnet_dir, "log", name)):
            r = rx.search(line)
            if r:
                rows.append((int(m.group(1)), int(m.group(2)), float(r.group(1)), float(r.group(3))))
    rows.sort()
    report = ["%Iter\tjob\ttrain_loss\ttrain_objective\ttrain_accuracy"]
    report += ["%d\t%d\t%.4f\t%.4f\t%.4f" % (it, job, loss, -loss, acc) for it, job, loss, acc in rows]
    return "\n".join(report) + "\n"


def train(args):
    logger.info("Arguments for the experiment\n{0}".format(pprint.pformat(vars(args))))
    rank, _ = sharding.dist_info()
    egs_dir = args.egs_dir
    [num_archives, egs_feat_dim, archives_minibatch_count] = verify_egs_dir(egs_dir)
    if args.num_jobs_final > num_archives:
        raise Exception('num_jobs_final cannot exceed the number of archives in the egs directory')
    os.makedirs(os.path.join(args.dir, "log"), exist_ok=True)
    if args.stage <= -1 and rank == 0:
        if not os.path.exists('{0}/model_0/done'.format(args.dir)):
            logger.info("Preparing the initial network.")
            model = getattr(models, args.tf_model_class)()
            logger.info("Start calling build_model to initialize the model %s ..." % args.tf_model_class)
            model.build_model(args.num_targets, egs_feat_dim, '{0}/model_0'.format(args.dir), logger=logger)
            with open(os.path.join(args.dir, 'model_name.txt'), 'wt') as fid:
                fid.write(args.tf_model_class)
        else:
            logger.info("The initial network exist from before.")
    if sharding.dist_info()[1] > 1:
        import torch.distributed as dist
        dist.barrier()
    num_archives_to_process = int(args.num_epochs * num_archives)
    num_archives_processed = 0
    num_iters = (num_archives_to_process * 2) // (args.num_jobs_initial + args.num_jobs_final)
    logger.info("Training will run for {0} epochs = {1} iterations".format(args.num_epochs, num_iters))
    for _iter in range(num_iters):
        current_num_jobs = int(0.5 + args.num_jobs_initial + (args.num_jobs_final - args.num_jobs_initial) * float(_iter) / num_iters)
        if args.stage <= _iter:
            lrate = get_learning_rate(_iter, current_num_jobs, num_iters, num_archives_processed, num_archives_to_process,
                                      args.initial_effective_lrate, args.final_effective_lrate)
            percent = num_archives_processed * 100.0 / num_archives_to_process
            epoch = (num_archives_processed * args.num_epochs / num_archives_to_process)
            logger.info("Iter: {0}/{1}    Epoch: {2:0.2f}/{3:0.1f} ({4:0.1f}% complete)    lr: {5:0.6f}    ".format(
                _iter, num_iters - 1, epoch, args.num_epochs, percent, lrate))
            train_one_iteration(args, _iter, egs_dir, current_num_jobs, num_archives_processed, num_archives, lrate,
                                archives_minibatch_count)
            if args.cleanup and rank == 0:
                remove_model(args.dir, _iter - 2, args.preserve_model_interval)
        num_archives_processed = num_archives_processed + current_num_jobs
    if rank == 0:
        if args.stage <= num_iters:
            link = "{0}/model_final".format(args.dir)
            if os.path.islink(link) or os.path.exists(link):
                os.remove(link)
            os.symlink("model_{0}".format(num_iters), link)                             # utils.force_symlink, train_dnn.py:583
        if args.cleanup:
            logger.info("Cleaning up the experiment directory {0}".format(args.dir))
            for _iter in range(num_iters):
                remove_model(args.dir, _iter, args.preserve_model_interval)
        with open("{dir}/accuracy.report".format(dir=args.dir), "wt") as fid:
            fid.write(generate_report(args.dir))
    return num_iters


def main():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(utils.pick_device())
        dist.init_process_group(backend="nccl")
    try:
        train(get_args())
        utils.wait_for_background_commands()
    except BaseException as e:
        if not isinstance(e, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
