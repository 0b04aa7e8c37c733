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

import numpy as np

from . import kaldi_io
from . import sharding
from .ze_utils import set_cuda_visible_devices

VAR2STD_EPSILON = 0.00001       # reference models.py:16
BN_EPSILON = 1e-3               # reference tf_block.py:9
META_FORMAT = "xvec-b200-v1"


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
                     bn_eps=BN_EPSILON, var_eps=VAR2STD_EPSILON, activation=meta.get("activation", "relu"))
    eng.set_params(params)
    return eng


def chunk_plan(num_rows, min_chunk_size, chunk_size):
    """(start, length) of the chunks the reference feeds to the network for one utterance, or
    None when it is skipped (reference models.py:378-409)."""
    if num_rows == 0 or num_rows < min_chunk_size:
        return None
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


class _Batch(object):
    __slots__ = ("slot", "n_frames", "seg_lens", "utts")

    def __init__(self, slot):
        self.slot = slot
        self.n_frames = 0
        self.seg_lens = []
        self.utts = []          # (global_index, key, first_segment, [lengths])


# noinspection PyAttributeOutsideInit
class Model(object):
    # reference models.py:27-29 (same shapes as ModelWithoutDropout; the base class adds dropout,
    # which is the identity at extraction time: keep_prob is fed 1.0, models.py:411)
    layer_sizes = [512, 512, 512, 512, 3 * 512]
    kernel_sizes = [5, 5, 7, 1, 1]
    dilation_rates = [1, 1, 1, 1, 1]
    embedding_sizes = [512, 512]
    activation = "relu"            # frame-layer nonlinearity: relu | lrelu (0.2) | prelu (per-channel)
    init = "trunc_normal"          # model_0 initialisation: trunc_normal (sigma 0.1) | he

    def __init__(self):
        self.graph = None
        self.params = None
        self.meta = None
        self._engine = None

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
                             activation=self.activation, init=self.init)
        meta = dict(format=META_FORMAT, model_class=type(self).__name__, num_classes=int(num_classes),
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
            np.savez(fid, **{k: np.asarray(v, dtype=np.float32) for k, v in sess.params.items()})
        with open(os.path.join(output_dir, "done"), "wt") as fid:
            fid.write("done")
        if logger is not None:
            logger.info("Graph saved in path: %s" % save_path)

    def load_model(self, sess, input_dir, logger):
        """``sess`` keeps the reference's positional slot (a TF session there); pass None."""
        if logger is not None:
            logger.info("Start loading graph ...")
        with open(os.path.join(input_dir, "model.meta"), "rt") as fid:
            try:
                meta = json.load(fid)
            except ValueError:
                raise RuntimeError("%s/model.meta is not an %s header (TensorFlow checkpoints must be converted "
                                   "first: see INTEGRATION.md)" % (input_dir, META_FORMAT))
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

    # ------------------------------------------------------------------ out of scope here
    def train_one_iteration(self, data_loader, args, logger):
        raise NotImplementedError("training (reference models.py:216-305) is outside the extraction hot path "
                                  "this build covers; see DESIGN.md 'out of scope'")

    def eval(self, data_loader, input_dir, use_gpu, logger):
        raise NotImplementedError("diagnostic eval (reference models.py:307-354) is outside the extraction hot "
                                  "path this build covers; see DESIGN.md 'out of scope'")

    # ------------------------------------------------------------------ extraction
    def _get_engine(self, device):
        if self._engine is None:
            self._engine = _create_engine(self.meta, self.params, device)
        return self._engine

    def make_embedding(self, input_stream, output_stream, model_dir, min_chunk_size, chunk_size, use_gpu, logger):
        start_time = time.time()
        device = set_cuda_visible_devices(use_gpu=use_gpu, logger=logger)
        self.load_model(None, model_dir, logger)
        engine = self._get_engine(device)
        feat_dim = self.meta["input_feature_dim"]
        emb_dim = self.embedding_sizes[0]
        batch_frames = int(os.environ.get("XVEC_BATCH_FRAMES", "400000"))
        rank, world = sharding.dist_info()

        # page-locked staging buffers: the reader thread fills one while two batches are in flight on the GPU
        staging = _Staging(feat_dim, batch_frames)
        counters = dict(total_segments=0, total_segments_len=0, num_fail=0, num_success=0)
        work = queue.Queue(maxsize=_Staging.SLOTS)
        failure = []

        def reader():
            try:
                self._read_batches(input_stream, staging, work, counters, min_chunk_size, chunk_size,
                                   batch_frames, rank, world, logger)
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

    def _read_batches(self, input_stream, staging, work, counters, min_chunk_size, chunk_size, batch_frames,
                      rank, world, logger):
        """Reader thread: parse the ark, apply the reference's skip/chunk rules, copy the rows of
        this rank's utterances into a pinned buffer and hand full batches to the GPU loop."""
        keys_in_order = []
        counters["keys_in_order"] = keys_in_order
        batch = None
        ok_index = 0
        for entry in kaldi_io.read_mat_ark_entries(input_stream):
            key, num_rows = entry.key, entry.rows
            if logger is not None:
                logger.debug("Processing features with key '%s' which have shape '%s'" % (key, str((entry.rows, entry.cols))))
            counters["total_segments"] += 1
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
            if batch is not None and batch.n_frames + num_rows > max(batch_frames, num_rows):
                work.put(batch)
                batch = None
            if batch is None:
                batch = _Batch(staging.acquire(max(batch_frames, num_rows)))
            # the payload goes straight from the stream into the page-locked buffer (no intermediate copy); rows of a
            # dropped tail are overwritten by the next utterance
            entry.read_into(staging.view(batch.slot, batch.n_frames + num_rows)[batch.n_frames:])
            first_seg = len(batch.seg_lens)
            for _, length in plan:
                batch.seg_lens.append(length)
            batch.n_frames += used
            batch.utts.append((this_index, key, first_seg, [n for _, n in plan]))
        if batch is not None and batch.utts:
            work.put(batch)


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

    def __init__(self, feat_dim, cap_frames):
        self.feat_dim = feat_dim
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
            self.caps[slot] = cap
        return slot

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


# noinspection PyAttributeOutsideInit
class ModelL2LossWithoutDropoutReluHeInit(ModelWithoutDropout):
    """ReLU with He-normal weights / He-uniform biases at model_0 (reference models.py:1118-1244)."""
    init = "he"
