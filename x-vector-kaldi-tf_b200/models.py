"""Host-side mirror of the reference's ``local/tf/models.py`` for the extraction hot path.

Same surface (``Model().build_model / save_model / load_model / get_models_weights /
make_embedding`` and the "subclass overrides the topology" plugin rule, reference
models.py:25,131,143,180,356 and README.md:17), but the TensorFlow session is gone: the forward
pass is the C-ABI library ``libxvec_b200.so`` (hand-written sm_100a kernels) reached through
``_native.XvecEngine``.  There is no CPU fallback.

Model directory (what ``load_model`` needs, cf. reference models.py:131-162 and
ze_utils.py:561-567):

    model.meta   non-empty; here a small JSON header (topology + sizes) instead of a TF MetaGraph
    model.npz    every variable the reference's Saver would hold for inference, keyed by the
                 reference's TF variable names ("frame_level_info_layer-0/w:0", ...)
    done         marker written last

``make_embedding`` keeps the reference's per-utterance semantics (skip rules, chunking,
frame-weighted average, one ark entry per utterance, the three summary log lines;
models.py:373-432) but evaluates many segments per launch.
"""
from __future__ import annotations

import json
import os
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import kaldi_io
from . import sharding
from .ze_utils import set_cuda_visible_devices

VAR2STD_EPSILON = 0.00001       # reference models.py:16
BN_EPSILON = 1e-3               # reference tf_block.py:9
META_FORMAT = "xvec-b200-v1"
# checkpoint entries that are optimizer state, not network variables: Adam slots and beta powers under TF's names, plus the
# explicit step counter (TF's float32 beta1_power underflows to 0 after ~990 steps, so it cannot carry the counter)
ADAM_STEP_KEY = "global_adam_step:0"
OPTIMIZER_STATE_SUFFIXES = ("/Adam:0", "/Adam_1:0", "_power:0", "_step:0")


class _Session(object):
    """Stands in for the ``sess`` argument of save_model/load_model: the variable store."""

    def __init__(self, params=None, meta=None):
        self.params = params if params is not None else {}
        self.meta = meta if meta is not None else {}


def _create_engine(meta, params, device):
    """Build the device engine for a loaded model (patched by CPU tests of the host logic)."""
    from ._native import XvecEngine
    eng = XvecEngine(meta["kernel_sizes"], meta["dilation_rates"], meta["layer_sizes"],
                     meta["embedding_sizes"][0], meta["input_feature_dim"], device=device,
                     bn_eps=BN_EPSILON, var_eps=VAR2STD_EPSILON, activation=meta.get("activation", "relu"),
                     pooling=meta.get("pooling", "stats"))
    # optimizer state saved by train_one_iteration ("<var>/Adam:0", "<var>/Adam_1:0", "beta*_power:0") is not a network variable
    eng.set_params({k: v for k, v in params.items() if not k.endswith(OPTIMIZER_STATE_SUFFIXES)})
    return eng


def chunk_plan(num_rows, min_chunk_size, chunk_size):
    """(start, length) of the chunks the reference feeds to the network for one utterance, or
    None when it is skipped (reference models.py:378-409)."""
    if num_rows == 0 or num_rows < min_chunk_size:
        return None
    if chunk_size == -1 or num_rows <= chunk_size:               # the usual case, one chunk: same result as the loop below
        return [(0, num_rows)]
    this_chunk_size = chunk_size
    if num_rows < chunk_size:
        this_chunk_size = num_rows
    elif chunk_size == -1:
        this_chunk_size = num_rows
    num_chunks = int(np.ceil(num_rows / float(this_chunk_size)))
    plan = []
    for chunk_idx in range(num_chunks):
        offset = min(this_chunk_size, num_rows - chunk_idx * this_chunk_size)
        if offset < min_chunk_size:
            continue
        plan.append((chunk_idx * this_chunk_size, offset))
    return plan


def adam_step_from_checkpoint(params):
    """Adam's step counter t of a checkpoint.  Ours carry it explicitly (``global_adam_step:0``, int64).  A checkpoint
    that only has TensorFlow's ``beta1_power`` / ``beta2_power`` (b^t, float32) gives t back from beta2_power -- 0.999^t is
    a normal float32 up to t ~ 87 000, while 0.9^t underflows to 0 near t = 990 and loses digits long before; a power
    that reads 0 is a SATURATED counter, never step 0 (restarting the bias correction while keeping the m / v slots
    would run every following update at a fraction of the intended learning rate)."""
    if ADAM_STEP_KEY in params:
        return int(np.asarray(params[ADAM_STEP_KEY]).reshape(-1)[0])
    b2p = float(np.asarray(params["beta2_power:0"]).reshape(-1)[0]) if "beta2_power:0" in params else None
    b1p = float(np.asarray(params["beta1_power:0"]).reshape(-1)[0]) if "beta1_power:0" in params else None
    if b2p is not None and 0.0 < b2p < 1.0:
        return int(round(np.log(b2p) / np.log(0.999)))
    if b1p is not None and 0.0 < b1p < 1.0:
        return int(round(np.log(b1p) / np.log(0.9)))
    if (b2p is not None and b2p <= 0.0) or (b1p is not None and b1p <= 0.0):
        return 1 << 20                      # both powers underflowed: bias correction is 1 to float precision from here on
    return 0


class _Batch(object):
    __slots__ = ("slot", "n_frames", "seg_lens", "utts", "raw_lens", "keeps", "pending", "group")

    def __init__(self, slot):
        self.slot = slot
        self.n_frames = 0       # rows in the staging buffer (raw rows when the feature front end runs on the device)
        self.seg_lens = []
        self.utts = []          # (global_index, key, first_segment, [lengths])
        self.raw_lens = []      # device front end only: raw rows / selected rows the network sees, per utterance
        self.keeps = []
        self.pending = []       # payload reads of this batch running on the reader pool
        self.group = None       # payloads collected for the next pool job

    GROUP_BYTES = 4 << 20

    def fetch_later(self, pool, fileobj, offset, dst):
        if self.group is None:
            self.group = kaldi_io.PayloadGroup()
        self.group.add(fileobj, offset, dst)
        if self.group.nbytes >= self.GROUP_BYTES:
            self.flush(pool)

    def flush(self, pool):
        if self.group is not None:
            self.pending.append(pool.submit(self.group.fetch))
            self.group = None

    def wait_for_payloads(self, pool):
        if pool is not None:
            self.flush(pool)
        for job in self.pending:
            job.result()        # re-raises what a read raised (truncated file, ...)
        self.pending = []


# noinspection PyAttributeOutsideInit
class Model(object):
    # reference models.py:27-29 (same shapes as ModelWithoutDropout; the base class adds dropout,
    # which is the identity at extraction time: keep_prob is fed 1.0, models.py:411)
    layer_sizes = [512, 512, 512, 512, 3 * 512]
    kernel_sizes = [5, 5, 7, 1, 1]
    dilation_rates = [1, 1, 1, 1, 1]
    embedding_sizes = [512, 512]
    pooling = "stats"              # stats (mean | std over time) | attention (models.py:1037-1051)
    l2_beta = 0.0                  # ModelL2Loss*: weight of the L2 term of the training loss (models.py:876, 961)
    activation = "relu"            # frame-layer nonlinearity: relu | lrelu (0.2) | prelu (per-channel)
    init = "trunc_normal"          # model_0 initialisation: trunc_normal (sigma 0.1) | he

    def __init__(self):
        self.graph = None
        self.params = None
        self.meta = None
        self._engine = None
        self._loaded_stamp = None

    # ------------------------------------------------------------------ build / save / load
    def build_model(self, num_classes, input_feature_dim, output_dir, logger=None):
        """Initialise ``model_0`` the way the reference's build_model does (truncated normal
        sigma=0.1 weights, bias 0.1, identity BatchNorm, xavier output layer; models.py:473-474,
        492-493, 504-506; tf_block.py:10-15) and save it."""
        from .synthetic import make_params
        if logger is not None:
            logger.info("Start building the model ...")
        self.num_classes = num_classes
        seed = os.environ.get("XVEC_SEED")
        seed = int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (2 ** 31))
        params = make_params(self.kernel_sizes, self.layer_sizes, self.embedding_sizes,
                             feat_dim=input_feature_dim, num_classes=num_classes, weight_set="A", seed=seed,
                             activation=self.activation, init=self.init, pooling=self.pooling)
        meta = dict(pooling=self.pooling, format=META_FORMAT, model_class=type(self).__name__, num_classes=int(num_classes),
                    input_feature_dim=int(input_feature_dim), kernel_sizes=list(self.kernel_sizes),
                    dilation_rates=list(self.dilation_rates), layer_sizes=list(self.layer_sizes),
                    embedding_sizes=list(self.embedding_sizes), activation=self.activation)
        if logger is not None:
            logger.info("Start initializing the graph ...")
        Model.save_model(_Session(params, meta), output_dir, logger)
        self.params, self.meta = params, meta
        if logger is not None:
            logger.info("Building finished.")

    @staticmethod
    def save_model(sess, output_dir, logger):
        if logger is not None:
            logger.info("Start saving graph ...")
        if not os.path.exists(output_dir):
            os.makedirs(output_dir)
        save_path = os.path.join(output_dir, "model")
        with open(save_path + ".meta", "wt") as fid:
            json.dump(sess.meta, fid, indent=1, sort_keys=True)
        with open(save_path + ".npz", "wb") as fid:
            np.savez(fid, **{k: (np.asarray(v, dtype=np.int64) if k.endswith("_step:0") else np.asarray(v, dtype=np.float32))
                             for k, v in sess.params.items()})
        with open(os.path.join(output_dir, "done"), "wt") as fid:
            fid.write("done")
        if logger is not None:
            logger.info("Graph saved in path: %s" % save_path)

    def load_model(self, sess, input_dir, logger):
        """``sess`` keeps the reference's positional slot (a TF session there); pass None."""
        if logger is not None:
            logger.info("Start loading graph ...")
        from . import tf_bundle
        meta = None
        try:
            with open(os.path.join(input_dir, "model.meta"), "rt") as fid:
                meta = json.load(fid)
        except (ValueError, UnicodeDecodeError):
            meta = None                              # a TensorFlow MetaGraph: not ours
        if meta is None and tf_bundle.is_bundle_dir(input_dir):
            # a checkpoint written by the reference's tf.train.Saver (models.py:131-141): variables by name from the bundle,
            # topology from this Model subclass (the MetaGraph is not needed) and the variable shapes
            params = tf_bundle.read_bundle(os.path.join(input_dir, "model"))
            w0, wo = params["frame_level_info_layer-0/w:0"], params.get("output/w:0")
            n_layers = len([k for k in params if k.startswith("frame_level_info_layer-") and k.endswith("/w:0")])
            taps = [int(params["frame_level_info_layer-%d/w:0" % i].shape[0]) for i in range(n_layers)]
            if taps != list(self.kernel_sizes):
                raise RuntimeError("%s holds kernel sizes %s but %s declares %s: load it with the Model subclass it was "
                                   "trained with (model_name.txt, train_dnn.py:495)" % (input_dir, taps, type(self).__name__,
                                                                                       list(self.kernel_sizes)))
            meta = dict(format=META_FORMAT, model_class=type(self).__name__, source="tensorflow-bundle",
                        num_classes=int(wo.shape[1]) if wo is not None else 0, input_feature_dim=int(w0.shape[1]),
                        kernel_sizes=taps, dilation_rates=list(self.dilation_rates),
                        layer_sizes=[int(params["frame_level_info_layer-%d/w:0" % i].shape[2]) for i in range(n_layers)],
                        embedding_sizes=[int(params["embed_layer-%d/w:0" % i].shape[1]) for i in range(2)
                                         if "embed_layer-%d/w:0" % i in params],
                        activation=self.activation, pooling=self.pooling)
        elif meta is None:
            raise RuntimeError("%s/model.meta is not an %s header and there is no model.index checkpoint bundle beside it"
                               % (input_dir, META_FORMAT))
        else:
            if meta.get("format") != META_FORMAT:
                raise RuntimeError("unsupported model.meta format %r" % meta.get("format"))
            with np.load(os.path.join(input_dir, "model.npz")) as z:
                params = {k: z[k] for k in z.files}
        self.meta, self.params = meta, params
        self.num_classes = meta["num_classes"]
        self.kernel_sizes = list(meta["kernel_sizes"])
        self.dilation_rates = list(meta["dilation_rates"])
        self.layer_sizes = list(meta["layer_sizes"])
        self.embedding_sizes = list(meta["embedding_sizes"])
        self.activation = meta.get("activation", "relu")
        self.pooling = meta.get("pooling", "stats")
        self.graph = meta
        if sess is not None:
            sess.params, sess.meta = params, meta
        if logger is not None:
            logger.info("Graph restored from path: %s" % input_dir)

    def print_models_params(self, input_dir, logger=None):
        self.load_model(None, input_dir, logger)
        print("\n\nThe components are:\n")
        for name in sorted(self.params):
            if not name.endswith(("mean:0", "variance:0")):
                print(name)
        print("\n")

    def get_models_weights(self, input_dir, logger=None):
        """name -> array dict (reference models.py:180-214, minus the h5py cache)."""
        self.load_model(None, input_dir, logger)
        return dict(self.params)

    # ------------------------------------------------------------------ training / diagnostics
    ADAM_SLOTS = ("/Adam:0", "/Adam_1:0")        # tf.train.AdamOptimizer slot names: "<var>/Adam", "<var>/Adam_1"

    def _create_trainer(self, device):
        """Engine + trainer holding this model's variables, moving statistics and (if saved) Adam state."""
        from ._native import XvecTrainer, TRAIN_ADAM_M, TRAIN_ADAM_V
        if self.activation not in ("relu", "lrelu") or self.pooling != "stats":
            raise NotImplementedError("the training step covers the ReLU / leaky-ReLU topologies with statistics pooling; "
                                      "%s uses %s / %s" % (type(self).__name__, self.activation, self.pooling))
        eng = self._get_engine(device)
        tr = XvecTrainer(eng, self.num_classes, self.embedding_sizes[1])
        if self.l2_beta:
            tr.set_option("l2_beta", self.l2_beta)
        state = {k: v for k, v in self.params.items() if not k.endswith(OPTIMIZER_STATE_SUFFIXES)}
        tr.set_params(state)
        for name in state:
            base = name[:-2]
            for suffix, which in zip(self.ADAM_SLOTS, (TRAIN_ADAM_M, TRAIN_ADAM_V)):
                if base + suffix in self.params:
                    _, off, cnt = tr.span(name)
                    tr.upload(which, self.params[base + suffix], off)
        tr.step = adam_step_from_checkpoint(self.params)
        return eng, tr

    def _download_state(self, tr):
        """Variables, moving statistics and Adam state of the trainer, keyed as the reference's checkpoint is."""
        from ._native import TRAIN_ADAM_M, TRAIN_ADAM_V
        out = {}
        for name, arr in self.params.items():
            if name.endswith(OPTIMIZER_STATE_SUFFIXES):
                continue
            which, off, cnt = tr.span(name)
            shape = np.asarray(arr).shape
            out[name] = tr.download(which, off, cnt).reshape(shape)
            if which == 0:
                out[name[:-2] + self.ADAM_SLOTS[0]] = tr.download(TRAIN_ADAM_M, off, cnt).reshape(shape)
                out[name[:-2] + self.ADAM_SLOTS[1]] = tr.download(TRAIN_ADAM_V, off, cnt).reshape(shape)
        out["beta1_power:0"] = np.asarray([0.9 ** tr.step], dtype=np.float32)
        out["beta2_power:0"] = np.asarray([0.999 ** tr.step], dtype=np.float32)
        out[ADAM_STEP_KEY] = np.asarray([tr.step], dtype=np.int64)
        return out

    def _run_minibatches(self, data_loader, tr, eng, logger, training, learning_rate=0.0, print_interval=10, data_parallel=True):
        """The minibatch loop of train_one_iteration / eval (reference models.py:233-299 / :313-354): same skip rules,
        counters and log lines; the host->device copy of minibatch k+1 overlaps the kernels of minibatch k, and
        loss / accuracy are read back only when a log line needs them."""
        import torch
        dev = torch.device("cuda:%d" % eng.device)
        # data parallel (gradient all-reduce) unless the caller runs independent jobs per rank (train_dnn.py: args.data_parallel=False)
        world = sharding.dist_info()[1] if data_parallel else 1
        compute, copy, comm = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        minibatch_count = data_loader.count
        if training and world > 1:
            # every rank must step the same number of times (one gradient all-reduce per minibatch): check before the loop
            import torch.distributed as dist
            lo = torch.tensor([minibatch_count, -minibatch_count], dtype=torch.int64, device=dev)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            if int(lo[0]) != -int(lo[1]):
                raise ValueError("data-parallel training needs the same number of minibatches on every rank: this rank has %d, "
                                 "the ranks hold between %d and %d" % (minibatch_count, int(lo[0]), -int(lo[1])))
        results = torch.zeros((max(minibatch_count, 1), 2), dtype=torch.float32, device=dev)
        grad = torch.zeros(tr.n_grad, dtype=torch.float32, device=dev) if (training and world > 1) else None
        slots = [dict(feats=None, labels=None, ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
        start_minibatch = 1
        total_segments, minibatch_segments, total_segments_len = 0, 0, 0
        total_gpu_waiting, total_disk_waiting = 0.0, 0.0
        loss_scale = None                           # None: the library's automatic loss scale
        done = []                                   # indices into `results` of the minibatches that ran
        window = []                                 # ... since the last log line
        start_time = time.time()

        def stage(k, batch_data, labels):
            """Minibatch -> page-locked staging -> device.  float16 minibatches (the egs archives' storage type) travel as
            float16 and are widened on the device; anything else is converted to float32 here."""
            sl = slots[k % 2]
            half = batch_data.dtype == np.float16
            x = np.ascontiguousarray(batch_data) if half else np.ascontiguousarray(batch_data, dtype=np.float32)
            n = x.size
            if sl["feats"] is None or sl["feats"].numel() < n:
                sl["feats"] = torch.empty(n, dtype=torch.float32).pin_memory()
                sl["feats_dev"] = torch.empty(n, dtype=torch.float32, device=dev)
                sl["half"] = torch.empty(n, dtype=torch.float16).pin_memory()
                sl["half_dev"] = torch.empty(n, dtype=torch.float16, device=dev)
            if sl["labels"] is None or sl["labels"].numel() < x.shape[0]:
                sl["labels"] = torch.empty(x.shape[0], dtype=torch.int32).pin_memory()
                sl["labels_dev"] = torch.empty(x.shape[0], dtype=torch.int32, device=dev)
            sl["free"].synchronize()                # the kernels that read this slot two minibatches ago are done
            if half:
                sl["half"].numpy()[:n] = x.reshape(-1)
            else:
                sl["feats"].numpy()[:n] = x.reshape(-1)
            sl["labels"].numpy()[:x.shape[0]] = np.asarray(labels, dtype=np.int32)
            with torch.cuda.stream(copy):
                if half:
                    sl["half_dev"][:n].copy_(sl["half"][:n], non_blocking=True)
                else:
                    sl["feats_dev"][:n].copy_(sl["feats"][:n], non_blocking=True)
                sl["labels_dev"][:x.shape[0]].copy_(sl["labels"][:x.shape[0]], non_blocking=True)
                sl["ready"].record(copy)
            sl["is_half"] = half
            return sl, x.shape

        skip_state = dict(seen=tr.skipped_updates(compute, blocking=True) if training else 0, lowered=0)
        n_seg_last = [0, 0]

        def lower_loss_scale(skipped, n_seg, seg_len, at_minibatch):
            nonlocal loss_scale
            if skipped <= skip_state["seen"]:
                return
            new = skipped - skip_state["seen"]
            skip_state["seen"] = skipped
            skip_state["lowered"] += 1
            if skip_state["lowered"] > 8:
                raise RuntimeError("gradients keep leaving the fp16 range after %d reductions of the loss scale (now %g): "
                                   "giving up at minibatch %d" % (skip_state["lowered"] - 1, loss_scale, at_minibatch))
            loss_scale = (loss_scale or float(2 ** int(np.ceil(np.log2(8.0 * n_seg * seg_len))))) / 16.0
            tr.set_option("loss_scale", loss_scale)
            logger.warning("fp16 overflow in the gradients around minibatch %d: %d update(s) skipped on every rank; loss scale "
                           "lowered to %g" % (at_minibatch, new, loss_scale))

        def check_forward_overflow(at_minibatch):
            """An ACTIVATION beyond the fp16 range (the xv_model's flag) is not curable by the loss scale: fail, on every
            rank together (the flag is combined first, or the healthy ranks would hang in the next all-reduce)."""
            err = None
            try:
                eng.check_overflow(compute)
            except Exception as e:                                       # noqa: BLE001
                err = e
            bad = 1 if err is not None else 0
            if world > 1:
                import torch.distributed as dist
                flag = torch.tensor([bad], dtype=torch.int32, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                bad = int(flag.item())
            if err is not None:
                raise err
            if bad:
                raise RuntimeError("another rank reported an activation overflow before minibatch %d" % at_minibatch)

        for minibatch_idx in range(minibatch_count):
            try:
                disk_waiting = time.time()
                batch_data, labels = data_loader.pop()
                total_disk_waiting += time.time() - disk_waiting
            except queue.Empty:
                logger.warning('Timeout reach when reading the minibatch index %d' % minibatch_idx)
                continue
            if batch_data is None:
                logger.warning('batch_data is None for the minibatch index %d' % minibatch_idx)
                continue
            minibatch_segments += batch_data.shape[0]
            total_segments += batch_data.shape[0]
            total_segments_len += batch_data.shape[1]
            gpu_waiting = time.time()
            sl, shape = stage(minibatch_idx, batch_data, labels)
            n_seg, seg_len = int(shape[0]), int(shape[1])
            n_seg_last[0], n_seg_last[1] = n_seg, seg_len
            with torch.cuda.stream(compute):
                compute.wait_event(sl["ready"])
                if sl["is_half"]:
                    tr.convert_f16(sl["half_dev"], sl["feats_dev"], n_seg * seg_len * shape[2], stream=compute)
                feats_dev = sl["feats_dev"][:n_seg * seg_len * shape[2]].view(n_seg * seg_len, shape[2])
                if training:
                    if grad is not None:
                        # data parallel: sum the flat gradient (+ the overflow flag behind it) over the ranks (NCCL); the
                        # segment-level 60 % of it is reduced on `comm` under the frame-level backward
                        la = tr.forward_backward_allreduce(feats_dev, sl["labels_dev"][:n_seg], n_seg, seg_len, grad, compute, comm)
                    else:
                        la = tr.forward_backward(feats_dev, sl["labels_dev"][:n_seg], n_seg, seg_len, grad_dev=grad, stream=compute)
                    tr.apply(learning_rate, grad_dev=grad, grad_scale=1.0 / world, stream=compute)
                else:
                    la = tr.evaluate(feats_dev, sl["labels_dev"][:n_seg], n_seg, seg_len, stream=compute)
                results[minibatch_idx].copy_(la, non_blocking=True)
                sl["free"].record(compute)
            done.append(minibatch_idx)
            window.append(minibatch_idx)
            total_gpu_waiting += time.time() - gpu_waiting
            end_minibatch = minibatch_idx + 1
            if training:
                # gradient overflow: the skip itself is decided on the device from the flag the all-reduce combined over the
                # ranks (every replica skips alike); here the count of skipped updates is polled WITHOUT waiting, every
                # step, so the loss scale drops a step or two after the first overflow instead of at the next log line
                lower_loss_scale(tr.skipped_updates(compute), n_seg, seg_len, end_minibatch)
            if training and end_minibatch % print_interval == 0:
                compute.synchronize()
                lower_loss_scale(tr.skipped_updates(compute, blocking=True), n_seg, seg_len, end_minibatch)
                check_forward_overflow(end_minibatch)
                res = results[window].cpu().numpy().astype(np.float64)
                cnt = end_minibatch - start_minibatch + 1
                minibatch_loss, minibatch_accuracy = res[:, 0].sum(), res[:, 1].sum()
                logger.info("Average training loss for minibatches %d-%d is %.4f over %d segments. Also, the "
                            "average training accuracy for these minibatches is %.4f and the average "
                            "objective function for these minibatches is %.4f. Average DISK waiting: %.1f "
                            "secs and average GPU waiting: %.1f secs for each minibatch." %
                            (start_minibatch, end_minibatch, minibatch_loss / cnt,
                             minibatch_segments, minibatch_accuracy / cnt, -minibatch_loss / cnt,
                             total_disk_waiting / cnt, total_gpu_waiting / cnt))
                start_minibatch = end_minibatch + 1
                minibatch_segments = 0
                window = []
                total_gpu_waiting = 0.0
                total_disk_waiting = 0.0
        compute.synchronize()
        if training:
            # the trailing (partial) window is handled like any other: skipped updates are logged, the iteration is kept
            lower_loss_scale(tr.skipped_updates(compute, blocking=True), max(n_seg_last[0], 1), max(n_seg_last[1], 1), minibatch_count)
        check_forward_overflow(minibatch_count)
        res = results[done].cpu().numpy().astype(np.float64) if done else np.zeros((0, 2))
        return dict(minibatch_count=minibatch_count, total_segments=total_segments, total_segments_len=total_segments_len,
                    total_loss=float(res[:, 0].sum()), total_accuracy=float(res[:, 1].sum()), losses=res[:, 0],
                    elapsed=time.time() - start_time)

    def train_one_iteration(self, data_loader, args, logger):
        """One pass over ``data_loader``'s minibatches with Adam (reference models.py:216-305): load the model of
        ``args.input_dir``, per minibatch run the training step (``sess.run([optimizer, loss, accuracy])``,
        models.py:263) at ``args.learning_rate``, log as the reference does, save to ``args.output_dir``.
        ``args.dropout_proportion`` is accepted and ignored by the topologies without dropout (the default
        ModelWithoutDropout family, run_xvector.sh:90); ``args.random_seed`` has nothing left to seed."""
        learning_rate = args.learning_rate
        print_interval = args.print_interval
        device = set_cuda_visible_devices(use_gpu=True, logger=logger)
        self.load_model(None, args.input_dir, logger)
        eng, tr = self._create_trainer(device)
        try:
            data_parallel = getattr(args, "data_parallel", True)
            st = self._run_minibatches(data_loader, tr, eng, logger, True, learning_rate, print_interval, data_parallel)
            minibatch_count = max(st["minibatch_count"], 1)
            logger.info("Processed %d segments of average size %d into %d minibatches. Avg minibatch size was %d." %
                        (st["total_segments"], st["total_segments_len"] / minibatch_count, minibatch_count,
                         st["total_segments"] / minibatch_count))
            logger.info("Overall average training loss is %.4f over %d segments. Also, the overall "
                        "average training accuracy is %.4f." % (st["total_loss"] / minibatch_count,
                                                                st["total_segments"], st["total_accuracy"] / minibatch_count))
            logger.info("Overall average objective function is %.4f over %d segments." %
                        (-st["total_loss"] / minibatch_count, st["total_segments"]))
            self.params = self._download_state(tr)
            if sharding.dist_info()[0] == 0 or not data_parallel:
                Model.save_model(_Session(self.params, self.meta), args.output_dir, logger)
            logger.info("Elapsed time for processing whole training minibatches is %.2f minutes." % (st["elapsed"] / 60.0))
            return st
        finally:
            tr.close()
            eng.close()
            self._engine = None

    def eval(self, data_loader, input_dir, use_gpu, logger):
        """Loss / accuracy over ``data_loader`` with phase=False (reference models.py:307-354)."""
        device = set_cuda_visible_devices(use_gpu=use_gpu, logger=logger)
        self.load_model(None, input_dir, logger)
        eng, tr = self._create_trainer(device)
        try:
            st = self._run_minibatches(data_loader, tr, eng, logger, False, data_parallel=False)
            minibatch_count = max(st["minibatch_count"], 1)
            logger.info("Processed %d segments of average size %d into %d minibatches. Avg minibatch size was %d." %
                        (st["total_segments"], st["total_segments_len"] / minibatch_count, minibatch_count,
                         st["total_segments"] / minibatch_count))
            logger.info("Overall average loss is %.4f over %d segments. Also, the overall "
                        "average accuracy is %.4f." % (st["total_loss"] / minibatch_count, st["total_segments"],
                                                       st["total_accuracy"] / minibatch_count))
            logger.info("Elapsed time for processing whole training minibatches is %.2f minutes." % (st["elapsed"] / 60.0))
            return st
        finally:
            tr.close()
            eng.close()
            self._engine = None

    # ------------------------------------------------------------------ extraction
    def _get_engine(self, device):
        if self._engine is None:
            self._engine = _create_engine(self.meta, self.params, device)
        return self._engine

    def make_embedding(self, input_stream, output_stream, model_dir, min_chunk_size, chunk_size, use_gpu, logger,
                       vad_table=None, cmvn_opts=None):
        """The reference's signature (models.py:356) plus two optional arguments that move the Kaldi pipe of
        local/tf/extract_xvectors.sh:68 onto the device: ``cmvn_opts`` (an ``_native.XvCmvnOpts``: apply-cmvn-sliding)
        and ``vad_table`` (a ``kaldi_io.VecTable`` of VAD decisions: select-voiced-frames).  With either one,
        ``input_stream`` carries RAW features and the skip / chunk rules apply to the voiced frames, as they do in the
        reference where the pipe runs first.  ``input_stream`` may also be an iterator of ``kaldi_io.MatArkEntry``."""
        start_time = time.time()
        device_frontend = vad_table is not None or cmvn_opts is not None
        if device_frontend and cmvn_opts is None:
            from ._native import XvCmvnOpts
            cmvn_opts = XvCmvnOpts()
        device = set_cuda_visible_devices(use_gpu=use_gpu, logger=logger)
        # a Model that extracts repeatedly from the same, unchanged model directory keeps its variables and its device
        # engine (the reference pays a full graph restore per call, models.py:365-366)
        stamp = _model_dir_stamp(model_dir)
        if self._engine is None or self._loaded_stamp != stamp:
            self.load_model(None, model_dir, logger)
            if self._engine is not None:
                getattr(self._engine, "close", lambda: None)()
                self._engine = None
            self._loaded_stamp = stamp
        engine = self._get_engine(device)
        if os.environ.get("XVEC_BLOCKING_COLLECT") is not None and hasattr(engine, "set_option"):
            engine.set_option("blocking_collect", int(os.environ["XVEC_BLOCKING_COLLECT"]))
        feat_dim = self.meta["input_feature_dim"]
        emb_dim = self.embedding_sizes[0]
        batch_frames = int(os.environ.get("XVEC_BATCH_FRAMES", "400000"))
        rank, world = sharding.dist_info()

        # A feature ark in a regular file takes the native job path: no per-utterance Python at all (headers indexed and
        # batches filled by libxvec_b200.so's reader, chunk averages finished on the device, vector ark formatted natively).
        # Pipes (the recipe's Kaldi feature pipe), in-memory streams, scp tables, text / compressed matrices, the device
        # front end and DEBUG logging (one line per utterance) keep the general path below.
        if not device_frontend and os.environ.get("XVEC_NATIVE_READER", "1") != "0" and \
                not (logger is not None and logger.isEnabledFor(10)):
            source = _regular_ark_file(input_stream)
            if source is not None and self._make_embedding_from_ark_file(source, output_stream, engine, device, min_chunk_size,
                                                                         chunk_size, batch_frames, logger, start_time):
                return

        # page-locked staging buffers: the reader thread fills one while two batches are in flight on the GPU
        staging = _Staging(feat_dim, batch_frames, with_vad=vad_table is not None)
        counters = dict(total_segments=0, total_segments_len=0, num_fail=0, num_success=0)
        work = queue.Queue(maxsize=_Staging.SLOTS)
        failure = []

        def reader():
            try:
                self._read_batches(input_stream, staging, work, counters, min_chunk_size, chunk_size,
                                   batch_frames, rank, world, logger, device_frontend, vad_table)
            except BaseException as e:      # surfaced on the main thread
                failure.append(e)
            finally:
                work.put(None)

        thread = threading.Thread(target=reader, name="xvec-ark-reader", daemon=True)
        thread.start()

        total_gpu_waiting = 0.0
        local_index, local_emb, local_keys = [], [], []
        pending = None                       # (ticket, batch, pinned embedding buffer) of the submission in flight
        while True:
            batch = work.get()
            submitted = None
            if batch is not None:
                feats = staging.view(batch.slot, batch.n_frames)
                emb_buf = staging.emb_view(batch.slot, len(batch.seg_lens), emb_dim)
                gpu_waiting = time.time()
                if device_frontend:
                    vad = staging.vad_view(batch.slot, batch.n_frames) if vad_table is not None else None
                    ticket = engine.submit_host_raw(feats, vad, batch.raw_lens, batch.keeps if vad is not None else None,
                                                    batch.seg_lens, emb_buf, cmvn_opts)
                else:
                    ticket = engine.submit_host(feats, np.asarray(batch.seg_lens, dtype=np.int32), emb_buf)
                total_gpu_waiting += time.time() - gpu_waiting
                submitted = (ticket, batch, emb_buf)
            if pending is not None:          # batch k-1 finishes while batch k copies in / computes
                ticket, done, seg_emb = pending
                gpu_waiting = time.time()
                engine.collect(ticket)
                total_gpu_waiting += time.time() - gpu_waiting
                out = _average_chunks(seg_emb, done.utts)
                staging.release(done.slot)
                if world == 1:
                    _write_vectors(output_stream, [u[1] for u in done.utts], out)
                else:
                    local_index.extend(u[0] for u in done.utts)
                    local_keys.extend(u[1] for u in done.utts)
                    local_emb.append(out)
            pending = submitted
            if batch is None:
                break
        thread.join()
        if failure:
            raise failure[0]

        if world > 1:
            n_ok = counters["num_success"]                       # every rank parsed the whole stream
            emb = np.concatenate(local_emb, axis=0) if local_emb else np.zeros((0, emb_dim), np.float32)
            full = sharding.gather_to_rank0(np.asarray(local_index, dtype=np.int64), emb, n_ok, emb_dim,
                                            device="cuda:%d" % device)
            if rank == 0:
                keys = counters["keys_in_order"]
                _write_vectors(output_stream, keys[:n_ok], full)

        if logger is not None:
            total_segments = max(counters["total_segments"], 1)
            logger.info("Processed %d features of average size %d frames. Done %d and failed %d" %
                        (counters["total_segments"], counters["total_segments_len"] / total_segments,
                         counters["num_success"], counters["num_fail"]))
            logger.info("Total time for neural network computations is %.2f minutes." % (total_gpu_waiting / 60.0))
            logger.info("Elapsed time for extracting whole embeddings is %.2f minutes." %
                        ((time.time() - start_time) / 60.0))

    def _make_embedding_from_ark_file(self, source, output_stream, engine, device, min_chunk_size, chunk_size, batch_frames,
                                      logger, start_time):
        """make_embedding over an ark in a regular file (reference models.py:373-432, same skip rules, chunk plan, average,
        output bytes and log lines).  Returns False -- having touched nothing -- when the archive holds entries the native
        scanner does not parse, so that the caller can take the general path.

        How the x-vectors of a multi-GPU job reach the ONE output (the reference has ``nj`` jobs write ``nj`` arks and
        concatenates their scp files, extract_xvectors.sh:83-95):
          * output in regular files (``ark:file``, ``ark,scp:A,S``): entry sizes are known once the keys are, so every rank
            formats its own rows and ``pwrite``s them into ITS byte range of the one ark while it computes; rank 0 only
            appends the scp lines the ranks hand it.  No x-vector crosses to rank 0 at all.
          * output is a stream on rank 0 (pipe, in-memory): every rank's utt_average_kernel stores its rows into rank 0's
            result table over NVLink peer memory, rank 0 reads the table after one barrier and writes the stream.
          * no peer memory (CPU process groups of the host-logic tests): one gather of the host rows."""
        from . import ark_job
        path, start = source
        feat_dim = self.meta["input_feature_dim"]
        on_gpu = getattr(engine, "handle", None) is not None       # a device engine (stand-ins of the host-logic tests have none)
        dev_name = "cuda:%d" % device if on_gpu else "cpu"
        t_job = time.time()
        # the reader rounds the rows to float16 while it copies them into the page-locked batches (the device's pack kernel
        # does that to a feature first anyway: same bits out, half the bytes through the host's memory and over PCIe);
        # not for the split-precision models, which need the float32 values
        feats_f16 = on_gpu and self.pooling != "attention" and os.environ.get("XVEC_FEED_F16", "1") != "0"
        reader, counts = ark_job.open_striped_reader(path, start, feat_dim, min_chunk_size, chunk_size, batch_frames,
                                                     device=dev_name, pinned=on_gpu, feats_f16=feats_f16)
        if reader is None:
            return False
        self._run_extraction_job(reader, counts, output_stream, engine, device, min_chunk_size, logger, start_time, t_job)
        return True

    def _run_extraction_job(self, reader, counts, output_stream, engine, device, min_chunk_size, logger, start_time, t_job):
        """The batch loop of an extraction job over a batch source (``_native.ArkReader``, or ``ark_job.SyntheticSource``
        whose features are generated on the device): submissions two deep, outputs as described above."""
        from . import ark_job
        rank, world = sharding.dist_info()
        emb_dim = self.embedding_sizes[0]
        on_gpu = getattr(engine, "handle", None) is not None
        dev_name = "cuda:%d" % device if on_gpu else "cpu"
        feats_on_device = bool(getattr(reader, "feats_on_device", False))
        stats = dict(pre_s=t_job - start_time, index_s=time.time() - t_job, reader_wait_s=0.0, submit_s=0.0, collect_s=0.0, write_s=0.0,
                     tail_s=0.0, batches=0)
        self.last_job_stats = stats                  # where the wall time of the last native job went (bench / diagnostics)
        peer, shared_fd, shared_map, shared_view, populate = None, None, None, None, None
        scp_fd, scp_map, scp_view, scp_cum = None, None, None, None
        try:
            if logger is not None:
                for key, reason, rows in reader.failures():
                    if reason == 1:
                        logger.warning("Zero-length utterance: '%s'" % key)
                    else:
                        logger.warning("Minimum chunk size of %d is greater than the number of rows in utterance: %s" %
                                       (min_chunk_size, key))
            n_ok_total = int(counts[:, 1].sum())
            base = int(counts[:rank, 1].sum())
            key_blob, key_off = reader.keys()
            entry_bytes = 11 + 4 * emb_dim               # an entry is its key + this (kaldi_io.write_vec_flt)
            mode = "local"
            shared = None
            if world > 1:
                import torch.distributed as dist
                total_out = int((counts[:, 5] + counts[:, 1] * entry_bytes).sum())
                box = [ark_job.shared_output_spec(output_stream, total_out) if rank == 0 else None]
                dist.broadcast_object_list(box, src=0)
                shared = box[0]
                mode = "shared_file" if shared is not None else ("peer" if on_gpu else "host_gather")
            stats["output"] = mode
            sink = ark_job.VectorSink(output_stream) if rank == 0 else None
            if mode == "peer":
                # rank 0's result table mapped into every rank (CUDA IPC).  Where the mapping cannot be made (processes in
                # different IPC namespaces, a driver without peer access between the devices) every rank learns it here and
                # the job gathers the host rows instead -- slower at the end, same bytes out.
                from ._native import PeerTable
                ok = 1
                try:
                    if rank == 0:
                        peer = PeerTable.create(device, max(n_ok_total, 1), emb_dim)
                except Exception as err:                                 # noqa: BLE001
                    ok, peer = 0, None
                    if logger is not None:
                        logger.warning("no peer-memory result table (%s): gathering host rows instead" % err)
                box = [peer.handle if (rank == 0 and peer is not None) else None]
                dist.broadcast_object_list(box, src=0)
                if rank != 0 and box[0] is not None:
                    try:
                        peer = PeerTable.open(device, max(n_ok_total, 1), emb_dim, box[0])
                    except Exception as err:                             # noqa: BLE001
                        ok, peer = 0, None
                        if logger is not None:
                            logger.warning("cannot map rank 0's result table (%s): gathering host rows instead" % err)
                elif rank != 0:
                    ok = 0
                if int(ark_job._all_gather_i64([ok], dev_name).min()) == 0:
                    if peer is not None:
                        peer.close()
                        peer = None
                    mode = "host_gather"
                    stats["output"] = mode
            ark_base = 0
            if mode == "shared_file":
                # this rank's byte range of the one ark, mapped: entries are formatted straight into the file's pages.
                # (write / pwrite would serialise the ranks on the file's inode lock: measured 98 ms for 26 MB per rank)
                import mmap
                ark_base = shared["base"] + int((counts[:rank, 5] + counts[:rank, 1] * entry_bytes).sum())
                my_bytes = int(counts[rank, 5] + counts[rank, 1] * entry_bytes)
                shared_fd = os.open(shared["ark"], os.O_RDWR)
                if my_bytes > 0:
                    map_from = ark_base // mmap.ALLOCATIONGRANULARITY * mmap.ALLOCATIONGRANULARITY
                    shared_map = mmap.mmap(shared_fd, ark_base - map_from + my_bytes, offset=map_from)
                    shared_view = np.frombuffer(shared_map, dtype=np.uint8)[ark_base - map_from:]
                    populate = ark_job.populate_pages(shared_map, ark_base - map_from, my_bytes)      # page faults off the critical path
                if shared["scp_name"] is not None:
                    # ... and of the one scp: a line's length follows from its key and the decimal digits of its offset, so the
                    # ranks exchange their scp sizes before a line exists and each writes its own range (rank 0 gathering and
                    # writing a million lines was the tail of the job: 57 of 443 ms at 8 GPUs)
                    scp_cum = ark_job.scp_line_offsets(key_off, shared["scp_name"], ark_base, entry_bytes)
                    scp_sizes = ark_job._all_gather_i64([int(scp_cum[-1])], dev_name)[:, 0]
                    if rank == 0:
                        ark_job.extend_shared_scp(output_stream, shared, int(scp_sizes.sum()))
                    dist.barrier()                       # the scp has its final size
                    my_scp, scp_base = int(scp_sizes[rank]), int(shared["scp_base"]) + int(scp_sizes[:rank].sum())
                    if my_scp > 0:
                        scp_fd = os.open(shared["scp"], os.O_RDWR)
                        map_from = scp_base // mmap.ALLOCATIONGRANULARITY * mmap.ALLOCATIONGRANULARITY
                        scp_map = mmap.mmap(scp_fd, scp_base - map_from + my_scp, offset=map_from)
                        scp_view = np.frombuffer(scp_map, dtype=np.uint8)[scp_base - map_from:]
            reader.start(base if mode == "peer" else 0)
            host_rows = [None, None]                     # page-locked [n_utt, emb_dim] per submission slot
            local = [] if mode == "host_gather" else None
            total_gpu_waiting = 0.0
            pending = None
            k = 0
            while True:
                t0 = time.time()
                b = reader.next()
                stats["reader_wait_s"] += time.time() - t0
                submitted = None
                if b is not None:
                    stats["batches"] += 1
                    out_host = None
                    if mode != "peer":
                        if host_rows[k & 1] is None or host_rows[k & 1].shape[0] < b.n_utt:
                            host_rows[k & 1] = _pinned_rows(max(2 * b.n_utt, 1024), emb_dim, on_gpu)
                        out_host = host_rows[k & 1][:b.n_utt]
                    t0 = time.time()
                    kw = dict(utt_first_seg=b.utt_first_seg, dst_rows=b.utt_dst_row if mode == "peer" else None,
                              out_dev=peer if mode == "peer" else None, out_host=out_host)
                    if feats_on_device:
                        ticket = engine.submit_dev_utts(b.feats, b.seg_len, ready_event=getattr(b, "ready_event", None), **kw)
                    else:
                        ticket = engine.submit_host_utts(b.feats, b.seg_len, **kw)
                    total_gpu_waiting += time.time() - t0
                    stats["submit_s"] += time.time() - t0
                    submitted = (ticket, b, out_host)
                    k += 1
                if pending is not None:
                    ticket, done, rows = pending
                    t0 = time.time()
                    engine.collect(ticket)
                    total_gpu_waiting += time.time() - t0
                    stats["collect_s"] += time.time() - t0
                    t0 = time.time()
                    window = key_off[done.first_ok_index:done.first_ok_index + done.n_utt + 1]
                    if mode == "local":
                        sink.write(key_blob, window, rows)
                    elif mode == "shared_file":
                        from ._native import scp_format, vec_ark_format
                        rel = int(window[0]) + done.first_ok_index * entry_bytes      # offset inside this rank's range
                        _, markers = vec_ark_format(key_blob, window, rows, with_markers=True, out=shared_view[rel:])
                        if shared["scp_name"] is not None:
                            lines = scp_format(key_blob, window, shared["scp_name"], ark_base + rel, markers)
                            lo = int(scp_cum[done.first_ok_index])
                            if lo + lines.shape[0] != int(scp_cum[done.first_ok_index + done.n_utt]):
                                raise RuntimeError("scp lines of a batch do not fill their planned byte range")
                            scp_view[lo:lo + lines.shape[0]] = lines
                    elif mode == "host_gather":
                        local.append(rows.copy())
                    reader.release(done.slot)
                    stats["write_s"] += time.time() - t0
                pending = submitted
                if b is None:
                    break
            t_tail = time.time()
            if mode == "shared_file":
                dist.barrier()                           # every rank's byte ranges of the ark and of the scp are written
                if rank == 0:
                    total = int((counts[:, 5] + counts[:, 1] * entry_bytes).sum())
                    ark_job.finish_shared_output(output_stream, shared, total)
            elif world > 1:
                blobs = ark_job.gather_bytes_to_rank0(key_blob, dev_name)
                lens = ark_job.gather_bytes_to_rank0(np.diff(key_off).astype(np.int32).view(np.uint8), dev_name)
                full = None
                if mode == "peer":
                    import torch
                    torch.cuda.synchronize(device)
                    dist.barrier()                       # every rank's rows are in rank 0's table
                else:
                    emb = np.concatenate(local, axis=0) if local else np.zeros((0, emb_dim), np.float32)
                    full = sharding.gather_to_rank0(np.arange(base, base + emb.shape[0], dtype=np.int64), emb, n_ok_total,
                                                    emb_dim, device=dev_name)
                if rank == 0:
                    all_blob = np.concatenate(blobs) if blobs else np.zeros(0, np.uint8)
                    all_len = np.concatenate([l.view(np.int32) for l in lens]) if lens else np.zeros(0, np.int32)
                    all_off = np.concatenate([[0], np.cumsum(all_len, dtype=np.int64)])
                    block = 1 << 14
                    stage = _pinned_rows(block, emb_dim, True) if mode == "peer" else None     # page-locked: D2H at full PCIe rate
                    for r0 in range(0, n_ok_total, block):
                        n = min(block, n_ok_total - r0)
                        rows = peer.read(r0, n, out=stage[:n]) if mode == "peer" else full[r0:r0 + n]
                        sink.write(all_blob, all_off[r0:r0 + n + 1], rows)
                if mode == "peer":
                    dist.barrier()                       # rank 0 has read the table: the mappings may go
            stats["tail_s"] = time.time() - t_tail
            stats["total_s"] = time.time() - t_job
            if logger is not None:
                n_entries, n_ok, n_fail, rows_used = (int(counts[:, i].sum()) for i in range(4))
                logger.info("Processed %d features of average size %d frames. Done %d and failed %d" %
                            (n_entries, rows_used / max(n_entries, 1), n_ok, n_fail))
                logger.info("Total time for neural network computations is %.2f minutes." % (total_gpu_waiting / 60.0))
                logger.info("Elapsed time for extracting whole embeddings is %.2f minutes." %
                            ((time.time() - start_time) / 60.0))
        finally:
            t0 = time.time()
            reader.close()
            if peer is not None:
                peer.close()
            shared_view = scp_view = None
            if scp_map is not None:
                try:
                    scp_map.close()
                except BufferError:
                    pass
            if scp_fd is not None:
                os.close(scp_fd)
            if populate is not None:
                populate.join()
            if shared_map is not None:
                try:
                    shared_map.close()
                except BufferError:                  # a view is still alive somewhere: the mapping goes with it
                    pass
            if shared_fd is not None:
                os.close(shared_fd)
            stats["close_s"] = time.time() - t0

    def _read_batches(self, input_stream, staging, work, counters, min_chunk_size, chunk_size, batch_frames,
                      rank, world, logger, device_frontend=False, vad_table=None):
        """Reader thread: parse the ark, apply the reference's skip/chunk rules, copy the rows of
        this rank's utterances into a pinned buffer and hand full batches to the GPU loop.

        With ``device_frontend`` the stream holds raw features: ``raw_rows`` of them are staged (plus the VAD track),
        while ``num_rows`` -- what the reference's loop sees after select-voiced-frames -- is the voiced count."""
        keys_in_order = []
        counters["keys_in_order"] = keys_in_order
        # matrices that sit in regular files (an ark on disk, the arks behind a feats.scp) are fetched by a few threads
        # with pread straight into the page-locked buffer, 4 MB per job: one thread's header parsing + read() gives
        # 45 M frames/s on the B200 box's host, the indexed / pooled path 63 M (profiles/r01_reader_throughput.txt), both
        # well below what the GPU consumes.  Pipes and in-memory streams keep the sequential path.
        n_readers = int(os.environ.get("XVEC_READER_THREADS", str(min(4, max(1, (os.cpu_count() or 2) // 2)))))
        pool = ThreadPoolExecutor(max_workers=n_readers, thread_name_prefix="xvec-pread") if n_readers > 1 else None
        try:
            self._read_batches_loop(input_stream, staging, work, counters, min_chunk_size, chunk_size, batch_frames,
                                    rank, world, logger, device_frontend, vad_table, pool, keys_in_order)
        finally:
            if pool is not None:
                pool.shutdown(wait=True)

    def _read_batches_loop(self, input_stream, staging, work, counters, min_chunk_size, chunk_size, batch_frames,
                           rank, world, logger, device_frontend, vad_table, pool, keys_in_order):
        batch = None
        ok_index = 0
        is_stream = isinstance(input_stream, str) or hasattr(input_stream, "read")
        if not is_stream:
            entries = input_stream                                   # an iterator of MatArkEntry (e.g. over a feats.scp)
        elif pool is not None and not isinstance(input_stream, str) and kaldi_io._plain_file(input_stream):
            from ._native import ark_scan                            # an ark on disk: native header index + parallel pread
            entries = kaldi_io.read_mat_ark_entries_indexed(input_stream, ark_scan)
        else:
            entries = kaldi_io.read_mat_ark_entries(input_stream)
        debug_lines = logger is not None and logger.isEnabledFor(10)      # logging.DEBUG
        for entry in entries:
            key, num_rows = entry.key, entry.rows
            raw_rows, vad = entry.rows, None
            if logger is not None:
                if debug_lines:
                    logger.debug("Processing features with key '%s' which have shape '%s'" % (key, str((entry.rows, entry.cols))))
            counters["total_segments"] += 1
            if vad_table is not None:
                # select-voiced-frames: no VAD / length mismatch / no voiced frame -> the utterance never reaches the network
                vad = vad_table.get(key)
                problem = None
                if vad is None:
                    problem = "No VAD decisions for utterance '%s'" % key
                elif vad.shape[0] != raw_rows:
                    problem = "Mismatch in number of frames for features and VAD of utterance '%s': %d vs %d" % (
                        key, raw_rows, vad.shape[0])
                else:
                    num_rows = int(np.count_nonzero(vad))
                    if num_rows == 0 and raw_rows > 0:
                        problem = "No features were judged as voiced for utterance '%s'" % key
                if problem is not None:
                    if logger is not None:
                        logger.warning(problem)
                    counters["num_fail"] += 1
                    continue
            if num_rows == 0:
                if logger is not None:
                    logger.warning("Zero-length utterance: '%s'" % key)
                counters["num_fail"] += 1
                continue
            if num_rows < min_chunk_size:
                if logger is not None:
                    logger.warning("Minimum chunk size of %d is greater than the number of rows in utterance: %s" %
                                   (min_chunk_size, key))
                counters["num_fail"] += 1
                continue
            plan = chunk_plan(num_rows, min_chunk_size, chunk_size)
            this_index = ok_index
            ok_index += 1
            counters["num_success"] += 1
            used = sum(n for _, n in plan)                   # chunks are contiguous from row 0; only a short tail is dropped
            counters["total_segments_len"] += used
            if world > 1:
                keys_in_order.append(key)
                if sharding.block_cyclic_rank(this_index, world) != rank:
                    continue                                 # another rank's utterance: payload skipped unread
            if entry.cols != staging.feat_dim:
                raise ValueError("utterance %s has feature dim %d, model expects %d" % (key, entry.cols, staging.feat_dim))
            if batch is not None and batch.n_frames + raw_rows > max(batch_frames, raw_rows):
                batch.wait_for_payloads(pool)
                work.put(batch)
                batch = None
            if batch is None:
                batch = _Batch(staging.acquire(max(batch_frames, raw_rows)))
            # the payload goes straight from the stream into the page-locked buffer (no intermediate copy)
            dst = staging.view(batch.slot, batch.n_frames + raw_rows)[batch.n_frames:]
            span = entry.detach_payload() if pool is not None else None
            if span is not None:
                # pooled pread jobs run concurrently and in any order: every job must own its destination rows.  Without
                # the device front end only the `used` rows are kept (a tail shorter than min_chunk_size is dropped,
                # models.py:404-405) and the next utterance starts right behind them, so only those rows are fetched -- a
                # full-payload read would land on the neighbour's head after the neighbour's own read
                batch.fetch_later(pool, span[0], span[1], dst if device_frontend else dst[:used])
            else:
                entry.read_into(dst)      # sequential stream: the dropped tail is overwritten by the next utterance's read
            first_seg = len(batch.seg_lens)
            for _, length in plan:
                batch.seg_lens.append(length)
            if device_frontend:                              # every raw row stays: the CMVN windows need them
                if vad is not None:
                    staging.vad_view(batch.slot, batch.n_frames + raw_rows)[batch.n_frames:] = vad
                batch.raw_lens.append(raw_rows)
                batch.keeps.append(used)                     # the front end drops a short tail chunk itself
                batch.n_frames += raw_rows
            else:
                batch.n_frames += used
            batch.utts.append((this_index, key, first_seg, [n for _, n in plan]))
        if batch is not None and batch.utts:
            batch.wait_for_payloads(pool)
            work.put(batch)


def _model_dir_stamp(model_dir):
    """Identity of a model directory's contents: path + size / mtime of the files load_model reads."""
    out = [os.path.realpath(model_dir)]
    for name in ("model.meta", "model.npz", "model.index", "done"):
        try:
            st = os.stat(os.path.join(model_dir, name))
            out.append((name, st.st_size, st.st_mtime_ns))
        except OSError:
            out.append((name, None))
    return tuple(out)


def _regular_ark_file(input_stream):
    """(path, start offset) when ``input_stream`` is a binary archive in a regular file the native reader can open: a
    buffered reader (its position is the start) or a plain / ``ark:``-prefixed file name; None otherwise."""
    import io
    import stat
    if isinstance(input_stream, io.BufferedReader):
        try:
            if stat.S_ISREG(os.fstat(input_stream.fileno()).st_mode):
                return "/proc/self/fd/%d" % input_stream.fileno(), input_stream.tell()
        except (OSError, ValueError):
            return None
        return None
    if isinstance(input_stream, str):
        name = input_stream.strip()
        if name.startswith("ark:"):
            name = name[4:]
        if "|" in name or name.endswith(".gz") or ":" in name or not os.path.isfile(name):
            return None
        return name, 0
    return None


def _pinned_rows(rows, cols, pinned):
    try:
        import torch
        return torch.empty((rows, cols), dtype=torch.float32, pin_memory=bool(pinned and torch.cuda.is_available())).numpy()
    except ImportError:
        return np.empty((rows, cols), dtype=np.float32)


def _write_vectors(output_stream, keys, vectors):
    """One ``write_vec_flt`` entry per utterance (reference models.py:422), batched per write."""
    if hasattr(output_stream, "write_vec_entries"):
        output_stream.write_vec_entries(keys, vectors)
    else:
        output_stream.write(b"".join(kaldi_io.vec_flt_entry_bytes(v, k) for k, v in zip(keys, vectors)))


def _average_chunks(seg_emb, utts):
    """Frame-weighted average of the chunk x-vectors of each utterance, in the reference's own
    float32 arithmetic (models.py:398-421): ``avg = sum(offset * xvector) / sum(offset)``."""
    out = np.empty((len(utts), seg_emb.shape[1]), dtype=np.float32)
    single = [i for i, u in enumerate(utts) if len(u[3]) == 1]
    if single:                                   # one chunk per utterance (every utterance <= chunk_size): same float32
        idx = np.asarray(single)                 # arithmetic as the loop below -- (offset * x) / offset -- vectorised
        first = np.asarray([utts[i][2] for i in single])
        w = np.asarray([utts[i][3][0] for i in single], dtype=np.float32)[:, None]
        out[idx] = (w * seg_emb[first]) / w
    for i, (_, _, first, lengths) in enumerate(utts):
        if len(lengths) == 1:
            continue
        xvector_avg = 0
        tot_weight = 0.0
        for c, offset in enumerate(lengths):
            tot_weight += offset
            xvector_avg = xvector_avg + offset * seg_emb[first + c]
        xvector_avg /= tot_weight
        out[i] = xvector_avg
    return out


class _Staging(object):
    """Three page-locked [cap, feat_dim] float32 buffers (+ one pinned embedding buffer each) cycled
    between the reader thread and the GPU loop: one being filled, two in flight."""
    SLOTS = 3

    def __init__(self, feat_dim, cap_frames, with_vad=False):
        self.feat_dim = feat_dim
        self.with_vad = with_vad
        self.vads = [None] * self.SLOTS      # [cap] float32 VAD track next to the raw rows (device front end)
        self.free = queue.Queue()
        self.bufs = [None] * self.SLOTS
        self.caps = [0] * self.SLOTS
        self.embs = [None] * self.SLOTS
        for s in range(self.SLOTS):
            self.free.put(s)
        self._cap0 = cap_frames

    def _alloc(self, rows, cols):
        try:
            import torch
            return torch.empty((rows, cols), dtype=torch.float32, pin_memory=torch.cuda.is_available()).numpy()
        except ImportError:
            return np.empty((rows, cols), dtype=np.float32)

    def emb_view(self, slot, n_seg, emb_dim):
        if self.embs[slot] is None or self.embs[slot].shape[0] < n_seg or self.embs[slot].shape[1] != emb_dim:
            self.embs[slot] = self._alloc(max(n_seg * 2, 1024), emb_dim)
        return self.embs[slot][:n_seg]

    def acquire(self, need_frames):
        slot = self.free.get()
        if self.caps[slot] < need_frames:
            cap = max(need_frames, self._cap0)
            self.bufs[slot] = self._alloc(cap, self.feat_dim)
            if self.with_vad:
                self.vads[slot] = self._alloc(cap, 1).reshape(-1)
            self.caps[slot] = cap
        return slot

    def vad_view(self, slot, n_frames):
        return self.vads[slot][:n_frames]

    def view(self, slot, n_frames):
        return self.bufs[slot][:n_frames]

    def release(self, slot):
        self.free.put(slot)


# noinspection PyAttributeOutsideInit
class ModelWithoutDropout(Model):
    """Recipe default (run_xvector.sh:90): dense kernels 5,5,7,1,1 (reference models.py:443-445)."""
    layer_sizes = [512, 512, 512, 512, 3 * 512]
    kernel_sizes = [5, 5, 7, 1, 1]
    dilation_rates = [1, 1, 1, 1, 1]
    embedding_sizes = [512, 512]


# noinspection PyAttributeOutsideInit
class ModelWithoutDropoutTdnn(Model):
    """Kaldi-style splice [-2..2],[-2,0,2],[-3,0,3],{0},{0} (reference models.py:545-548)."""
    layer_sizes = [512, 512, 512, 512, 3 * 512]
    kernel_sizes = [5, 3, 3, 1, 1]
    dilation_rates = [1, 2, 3, 1, 1]
    embedding_sizes = [512, 512]


# The remaining topologies of the reference differ from ModelWithoutDropout only in the frame-layer
# nonlinearity, the initialisation and (for training, out of scope here) an L2 term in the loss; their
# extraction forward is the same fused kernel with another activation.


# noinspection PyAttributeOutsideInit
class ModelWithoutDropoutPRelu(ModelWithoutDropout):
    """prelu(h, shared=False) after bias_add (reference models.py:643-744, tf_block.py:38-47)."""
    activation = "prelu"


# noinspection PyAttributeOutsideInit
class ModelL2LossWithoutDropoutPRelu(ModelWithoutDropoutPRelu):
    """Same forward as ModelWithoutDropoutPRelu; adds an L2 loss term for training (reference models.py:746-864)."""


# noinspection PyAttributeOutsideInit
class ModelL2LossWithoutDropoutLRelu(ModelWithoutDropout):
    """tf.nn.leaky_relu(h, alpha=0.2) (reference models.py:866-983)."""
    activation = "lrelu"
    l2_beta = 0.0002


# noinspection PyAttributeOutsideInit
class ModelL2LossWithoutDropoutReluHeInit(ModelWithoutDropout):
    """ReLU with He-normal weights / He-uniform biases at model_0 (reference models.py:1118-1244)."""
    init = "he"
    l2_beta = 0.0002


class ModelL2LossWithoutDropoutLReluAttention(ModelL2LossWithoutDropoutLRelu):
    """Self-attention statistics pooling (reference models.py:983-1100): the last frame layer is 6*512 wide and split
    into the score input h1 and the pooled half h2; ``attention/{w,b,v}``; leaky ReLU 0.2 frame layers."""
    layer_sizes = [512, 512, 512, 512, 6 * 512]
    pooling = "attention"
