"""ctypes binding of libxvec_b200.so (C ABI in include/xvec.h).

There is deliberately NO fallback: if the shared library is missing or no B200 is present the
calls raise -- the product path never routes through a CPU implementation.  PyTorch is used
only as the device/pinned memory container (``tensor.data_ptr()``) and for stream handles.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libxvec_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)
CSRC = os.path.join(_HERE, "csrc")
REPO_ROOT = os.path.dirname(_HERE)

XV_MAX_FRAME_LAYERS = 8
XV_ABI_VERSION = 2
XV_OK, XV_EINVAL, XV_ECUDA, XV_ENOMEM, XV_ESTATE, XV_EOVERFLOW = 0, -1, -2, -3, -4, -5

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class XvTopology(ctypes.Structure):
    _fields_ = [("feat_dim", ctypes.c_int32), ("n_frame_layers", ctypes.c_int32),
                ("taps", ctypes.c_int32 * XV_MAX_FRAME_LAYERS),
                ("dilation", ctypes.c_int32 * XV_MAX_FRAME_LAYERS),
                ("width", ctypes.c_int32 * XV_MAX_FRAME_LAYERS),
                ("emb_dim", ctypes.c_int32), ("act", ctypes.c_int32),
                ("bn_eps", ctypes.c_float), ("var_eps", ctypes.c_float), ("pooling", ctypes.c_int32)]


class XvCmvnOpts(ctypes.Structure):
    """xv_cmvn_opts (include/xvec_frontend.h): the options of Kaldi's apply-cmvn-sliding; defaults are what the
    reference passes (local/tf/extract_xvectors.sh:68)."""
    _fields_ = [("cmn_window", ctypes.c_int32), ("min_window", ctypes.c_int32), ("center", ctypes.c_int32),
                ("normalize_variance", ctypes.c_int32)]

    def __init__(self, cmn_window=300, min_window=100, center=True, normalize_variance=False):
        super().__init__(int(cmn_window), int(min(min_window, cmn_window)), int(bool(center)), int(bool(normalize_variance)))


class XvArkReaderOpts(ctypes.Structure):
    """xv_ark_reader_opts (include/xvec_job.h)."""
    _fields_ = [("feat_dim", ctypes.c_int32), ("min_chunk_size", ctypes.c_int32), ("chunk_size", ctypes.c_int32),
                ("n_threads", ctypes.c_int32), ("n_slots", ctypes.c_int32), ("pinned", ctypes.c_int32),
                ("batch_frames", ctypes.c_int64), ("byte_begin", ctypes.c_int64), ("byte_end", ctypes.c_int64),
                ("begin_is_boundary", ctypes.c_int32), ("feats_f16", ctypes.c_int32)]


class XvArkIndexInfo(ctypes.Structure):
    """xv_ark_index_info (include/xvec_job.h)."""
    _fields_ = [(n, ctypes.c_int64) for n in ("n_entries", "n_ok", "n_fail", "n_segments", "rows_used", "n_batches",
                                              "first_marker_off", "next_marker_off", "next_key_off", "stopped_at", "key_bytes")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class XvArkBatch(ctypes.Structure):
    """xv_ark_batch (include/xvec_job.h)."""
    _fields_ = [("slot", ctypes.c_int32), ("n_seg", ctypes.c_int32), ("n_utt", ctypes.c_int32), ("feats_f16", ctypes.c_int32),
                ("n_rows", ctypes.c_int64), ("feats", ctypes.c_void_p), ("seg_len", ctypes.c_void_p),
                ("utt_first_seg", ctypes.c_void_p), ("utt_dst_row", ctypes.c_void_p), ("first_ok_index", ctypes.c_int64)]


class XvecError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("xvec_b200 error %d: %s" % (code, message))
        self.code = code


def build_library(verbose=False):
    """Compile csrc/xvec_api.cu for sm_100a into the in-tree shared library (nvcc cross-compiles
    without a GPU).  Skips the compile when the library is newer than every source."""
    sources = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp"))]
    sources.append(os.path.join(REPO_ROOT, "include", "xvec.h"))
    sources.append(os.path.join(REPO_ROOT, "include", "xvec_train.h"))
    sources.append(os.path.join(REPO_ROOT, "include", "xvec_frontend.h"))
    sources.append(os.path.join(REPO_ROOT, "include", "xvec_job.h"))
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in sources):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    # f16_convert.cpp is plain C++ (AVX2 / F16C intrinsics behind a runtime check): nvcc hands it to the host compiler as it is
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "xvec_api.cu"), os.path.join(CSRC, "f16_convert.cpp")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load_library():
    """Load libxvec_b200.so; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XvecError(XV_ESTATE, "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    P, I32, I64, SZ = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
    lib.xv_create.argtypes = [ctypes.POINTER(P), ctypes.c_int, ctypes.POINTER(XvTopology)]
    lib.xv_create.restype = ctypes.c_int
    lib.xv_destroy.argtypes = [P]
    lib.xv_destroy.restype = None
    lib.xv_set_param.argtypes = [P, ctypes.c_char_p, P, ctypes.POINTER(I64), I32]
    lib.xv_set_param.restype = ctypes.c_int
    lib.xv_workspace_bytes.argtypes = [P, I64, I32]
    lib.xv_workspace_bytes.restype = SZ
    lib.xv_forward.argtypes = [P, P, P, I32, P, P, SZ, P]
    lib.xv_forward.restype = ctypes.c_int
    lib.xv_forward_layers.argtypes = [P, P, P, I32, P, P, SZ, P, ctypes.POINTER(P), P]
    lib.xv_forward_layers.restype = ctypes.c_int
    lib.xv_extract_host.argtypes = [P, P, P, I32, P]
    lib.xv_extract_host.restype = ctypes.c_int
    lib.xv_submit_host.argtypes = [P, P, P, I32, P, ctypes.POINTER(I32)]
    lib.xv_submit_host.restype = ctypes.c_int
    lib.xv_collect.argtypes = [P, I32]
    lib.xv_collect.restype = ctypes.c_int
    lib.xv_forward_utts.argtypes = [P, P, P, I32, P, P, I32, P, P, SZ, P]
    lib.xv_forward_utts.restype = ctypes.c_int
    lib.xv_submit_host_utts.argtypes = [P, P, P, I32, P, P, I32, P, P, ctypes.POINTER(I32)]
    lib.xv_submit_host_utts.restype = ctypes.c_int
    lib.xv_submit_host_utts_f16.argtypes = [P, P, P, I32, P, P, I32, P, P, ctypes.POINTER(I32)]
    lib.xv_submit_host_utts_f16.restype = ctypes.c_int
    lib.xv_convert_f32_to_f16_host.argtypes = [P, P, SZ]
    lib.xv_convert_f32_to_f16_host.restype = None
    lib.xv_convert_f32_to_f16_host_scalar.argtypes = [P, P, SZ]
    lib.xv_convert_f32_to_f16_host_scalar.restype = None
    lib.xv_peer_alloc.argtypes = [ctypes.c_int, SZ, ctypes.POINTER(P), P]
    lib.xv_peer_alloc.restype = ctypes.c_int
    lib.xv_peer_open.argtypes = [ctypes.c_int, P, ctypes.POINTER(P)]
    lib.xv_peer_open.restype = ctypes.c_int
    lib.xv_peer_close.argtypes = [ctypes.c_int, P]
    lib.xv_peer_close.restype = ctypes.c_int
    lib.xv_peer_free.argtypes = [ctypes.c_int, P]
    lib.xv_peer_free.restype = ctypes.c_int
    lib.xv_peer_read.argtypes = [ctypes.c_int, P, P, SZ]
    lib.xv_peer_read.restype = ctypes.c_int
    lib.xv_abi_version.argtypes = []
    lib.xv_abi_version.restype = I32
    lib.xv_topology_size.argtypes = []
    lib.xv_topology_size.restype = SZ
    if lib.xv_abi_version() != XV_ABI_VERSION or lib.xv_topology_size() != ctypes.sizeof(XvTopology):
        raise XvecError(XV_ESTATE, "%s is ABI version %d with a %d-byte xv_topology; this binding is version %d / %d bytes: "
                                    "rebuild the library" % (LIB_PATH, lib.xv_abi_version(), lib.xv_topology_size(),
                                                             XV_ABI_VERSION, ctypes.sizeof(XvTopology)))
    lib.xv_check_overflow.argtypes = [P, P]
    lib.xv_check_overflow.restype = ctypes.c_int
    lib.xv_rescue_overflow.argtypes = [P, P]
    lib.xv_rescue_overflow.restype = ctypes.c_int
    lib.xv_last_launch_count.argtypes = [P]
    lib.xv_last_launch_count.restype = I32
    lib.xv_last_kernel_ms.argtypes = [P, ctypes.POINTER(ctypes.c_float), I32]
    lib.xv_last_kernel_ms.restype = I32
    lib.xv_set_option.argtypes = [P, ctypes.c_char_p, I64]
    lib.xv_set_option.restype = ctypes.c_int
    lib.xv_last_error.argtypes = []
    lib.xv_last_error.restype = ctypes.c_char_p
    lib.xv_version.argtypes = []
    lib.xv_version.restype = ctypes.c_char_p
    lib.xv_ark_scan.argtypes = [P, I64, I64, P, P, P, P, P, P, ctypes.POINTER(I64)]
    lib.xv_ark_scan.restype = I64
    # extraction job: striped ark reader + vector-ark formatter (include/xvec_job.h)
    lib.xv_ark_reader_open.argtypes = [ctypes.POINTER(P), ctypes.c_char_p, ctypes.POINTER(XvArkReaderOpts)]
    lib.xv_ark_reader_open.restype = ctypes.c_int
    lib.xv_ark_reader_index.argtypes = [P, ctypes.POINTER(XvArkIndexInfo)]
    lib.xv_ark_reader_index.restype = ctypes.c_int
    lib.xv_ark_reader_set_first.argtypes = [P, I64, I64, ctypes.POINTER(XvArkIndexInfo)]
    lib.xv_ark_reader_set_first.restype = ctypes.c_int
    lib.xv_ark_reader_keys.argtypes = [P, P, I64, P]
    lib.xv_ark_reader_keys.restype = ctypes.c_int
    lib.xv_ark_reader_failures.argtypes = [P, P, P, P, I64, P]
    lib.xv_ark_reader_failures.restype = ctypes.c_int
    lib.xv_ark_reader_start.argtypes = [P, I64]
    lib.xv_ark_reader_start.restype = ctypes.c_int
    lib.xv_ark_reader_next.argtypes = [P, ctypes.POINTER(XvArkBatch)]
    lib.xv_ark_reader_next.restype = ctypes.c_int
    lib.xv_ark_reader_release.argtypes = [P, I32]
    lib.xv_ark_reader_release.restype = ctypes.c_int
    lib.xv_ark_reader_close.argtypes = [P]
    lib.xv_ark_reader_close.restype = None
    lib.xv_vec_ark_bytes.argtypes = [P, I64, I32]
    lib.xv_vec_ark_bytes.restype = I64
    lib.xv_vec_ark_format.argtypes = [P, P, I64, P, I32, P, I64, P, I32]
    lib.xv_vec_ark_format.restype = I64
    lib.xv_scp_format.argtypes = [P, P, I64, ctypes.c_char_p, I64, P, P, I64]
    lib.xv_scp_format.restype = I64
    lib.xv_synth_mfcc.argtypes = [ctypes.c_int, P, P, P, I32, I32, ctypes.c_uint64, P]
    lib.xv_synth_mfcc.restype = ctypes.c_int
    lib.xv_submit_dev_utts.argtypes = [P, P, P, I32, P, P, I32, P, P, P, ctypes.POINTER(I32)]
    lib.xv_submit_dev_utts.restype = ctypes.c_int
    # training step (include/xvec_train.h)
    F32, F64 = ctypes.c_float, ctypes.c_double
    lib.xv_train_create.argtypes = [ctypes.POINTER(P), P, I32, I32]
    lib.xv_train_create.restype = ctypes.c_int
    lib.xv_train_destroy.argtypes = [P]
    lib.xv_train_destroy.restype = None
    lib.xv_train_size.argtypes = [P, I32]
    lib.xv_train_size.restype = I64
    lib.xv_train_span.argtypes = [P, ctypes.c_char_p, ctypes.POINTER(I32), ctypes.POINTER(I64), ctypes.POINTER(I64)]
    lib.xv_train_span.restype = ctypes.c_int
    lib.xv_train_upload.argtypes = [P, I32, P, I64, I64]
    lib.xv_train_upload.restype = ctypes.c_int
    lib.xv_train_download.argtypes = [P, I32, P, I64, I64]
    lib.xv_train_download.restype = ctypes.c_int
    lib.xv_train_set_step.argtypes = [P, I64]
    lib.xv_train_set_step.restype = ctypes.c_int
    lib.xv_train_get_step.argtypes = [P]
    lib.xv_train_get_step.restype = I64
    lib.xv_train_forward_backward.argtypes = [P, P, P, I32, I32, P, P, P]
    lib.xv_train_forward_backward.restype = ctypes.c_int
    lib.xv_train_forward_backward_part.argtypes = [P, P, P, I32, I32, P, P, P, I32]
    lib.xv_train_forward_backward_part.restype = ctypes.c_int
    lib.xv_train_segment_grad_offset.argtypes = [P]
    lib.xv_train_segment_grad_offset.restype = I64
    lib.xv_train_frame_grad_span.argtypes = [P, I32, ctypes.POINTER(I64), ctypes.POINTER(I64)]
    lib.xv_train_frame_grad_span.restype = ctypes.c_int
    lib.xv_train_eval.argtypes = [P, P, P, I32, I32, P, P]
    lib.xv_train_eval.restype = ctypes.c_int
    lib.xv_train_apply.argtypes = [P, P, F32, F32, P]
    lib.xv_train_apply.restype = ctypes.c_int
    lib.xv_train_skipped_updates.argtypes = [P, P, I32]
    lib.xv_train_skipped_updates.restype = I64
    lib.xv_train_sync_model.argtypes = [P]
    lib.xv_train_sync_model.restype = ctypes.c_int
    lib.xv_train_debug_tensor.argtypes = [P, ctypes.c_char_p, P, I64]
    lib.xv_train_debug_tensor.restype = I64
    lib.xv_train_set_option.argtypes = [P, ctypes.c_char_p, F64]
    lib.xv_train_set_option.restype = ctypes.c_int
    lib.xv_train_last_launch_count.argtypes = [P]
    lib.xv_train_last_launch_count.restype = I32
    lib.xv_convert_f16_to_f32.argtypes = [P, P, I64, P]
    lib.xv_convert_f16_to_f32.restype = ctypes.c_int
    lib.xv_train_last_kernel_names.argtypes = [P, ctypes.c_char_p, I64]
    lib.xv_train_last_kernel_names.restype = I64
    # feature front end (include/xvec_frontend.h)
    lib.xv_frontend_workspace_bytes.argtypes = [P, I64, I32]
    lib.xv_frontend_workspace_bytes.restype = SZ
    lib.xv_frontend.argtypes = [P, P, P, P, P, I32, ctypes.POINTER(XvCmvnOpts), P, P, SZ, P]
    lib.xv_frontend.restype = ctypes.c_int
    lib.xv_submit_host_raw.argtypes = [P, P, P, P, P, I32, ctypes.POINTER(XvCmvnOpts), P, I32, P, ctypes.POINTER(I32)]
    lib.xv_submit_host_raw.restype = ctypes.c_int
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ["xv_create", "xv_destroy", "xv_set_param", "xv_workspace_bytes", "xv_forward",
                    "xv_forward_layers", "xv_extract_host", "xv_submit_host", "xv_collect", "xv_check_overflow", "xv_rescue_overflow", "xv_last_launch_count",
                    "xv_last_kernel_ms", "xv_set_option", "xv_last_error", "xv_version", "xv_ark_scan",
                    "xv_forward_utts", "xv_submit_host_utts", "xv_submit_host_utts_f16", "xv_peer_alloc", "xv_peer_open", "xv_peer_close", "xv_peer_free",
                    "xv_peer_read", "xv_abi_version", "xv_topology_size",
                    # include/xvec_job.h
                    "xv_ark_reader_open", "xv_ark_reader_index", "xv_ark_reader_set_first", "xv_ark_reader_keys",
                    "xv_ark_reader_failures", "xv_ark_reader_start", "xv_ark_reader_next", "xv_ark_reader_release",
                    "xv_ark_reader_close", "xv_vec_ark_bytes", "xv_vec_ark_format", "xv_scp_format", "xv_synth_mfcc",
                    "xv_submit_dev_utts", "xv_convert_f32_to_f16_host", "xv_convert_f32_to_f16_host_scalar",
                    # include/xvec_train.h
                    "xv_train_create", "xv_train_destroy", "xv_train_size", "xv_train_span", "xv_train_upload",
                    "xv_train_download", "xv_train_set_step", "xv_train_get_step", "xv_train_forward_backward", "xv_train_forward_backward_part",
                    "xv_train_segment_grad_offset", "xv_train_frame_grad_span", "xv_train_eval",
                    "xv_train_apply", "xv_train_skipped_updates", "xv_train_sync_model", "xv_train_debug_tensor", "xv_train_set_option",
                    "xv_train_last_launch_count", "xv_train_last_kernel_names", "xv_convert_f16_to_f32",
                    # include/xvec_frontend.h
                    "xv_frontend_workspace_bytes", "xv_frontend", "xv_submit_host_raw"]


def _check(lib, rc):
    if rc != XV_OK:
        raise XvecError(rc, lib.xv_last_error().decode(errors="replace"))


def ark_scan(buffer, start=0, max_entries=1 << 20):
    """Index of the binary float matrices of a Kaldi ark held in ``buffer`` (bytes-like, e.g. an mmap) from byte
    ``start``: (key_off, key_len, rows, cols, elem_bytes, payload_off) numpy arrays -- offsets relative to the start of
    ``buffer`` -- and the offset of the first byte that was not parsed.  Host-only (xv_ark_scan); needs no GPU."""
    lib = load_library()
    view = np.frombuffer(buffer, dtype=np.uint8)
    n_max = int(min(max_entries, max(1, (view.shape[0] - start) // 16 + 1)))
    key_off = np.empty(n_max, np.int64)
    key_len = np.empty(n_max, np.int32)
    rows = np.empty(n_max, np.int32)
    cols = np.empty(n_max, np.int32)
    elem = np.empty(n_max, np.int32)
    pay = np.empty(n_max, np.int64)
    consumed = ctypes.c_int64(0)
    n = int(lib.xv_ark_scan(view.ctypes.data + start, view.shape[0] - start, n_max, key_off.ctypes.data, key_len.ctypes.data,
                            rows.ctypes.data, cols.ctypes.data, elem.ctypes.data, pay.ctypes.data, ctypes.byref(consumed)))
    if n < 0:
        raise XvecError(XV_EINVAL, "xv_ark_scan: bad argument")
    return (key_off[:n] + start, key_len[:n], rows[:n], cols[:n], elem[:n], pay[:n] + start), start + int(consumed.value)


class ArkBatch(object):
    """One batch of an ArkReader: numpy views over the reader's own memory, valid until ``release``."""
    __slots__ = ("slot", "n_seg", "n_utt", "n_rows", "feats", "seg_len", "utt_first_seg", "utt_dst_row", "first_ok_index")


def _view(ptr, ctype, n, dtype):
    if n == 0:
        return np.empty(0, dtype)
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,))


class ArkReader(object):
    """xv_ark_reader (include/xvec_job.h): the host side of an extraction job over a feature ark in a regular file --
    header index of one byte stripe, make_embedding's skip / chunk rules, batches filled by a pread pool.  Needs no GPU
    with ``pinned=False``."""

    def __init__(self, path, feat_dim, min_chunk_size, chunk_size, batch_frames, byte_begin=0, byte_end=-1,
                 begin_is_boundary=True, n_threads=4, n_slots=3, pinned=True, feats_f16=False):
        self.lib = load_library()
        opts = XvArkReaderOpts(int(feat_dim), int(min_chunk_size), int(chunk_size), int(n_threads), int(n_slots),
                               1 if pinned else 0, int(batch_frames), int(byte_begin), int(byte_end),
                               1 if begin_is_boundary else 0, 1 if feats_f16 else 0)
        self.feat_dim = int(feat_dim)
        self.handle = ctypes.c_void_p()
        _check(self.lib, self.lib.xv_ark_reader_open(ctypes.byref(self.handle), os.fsencode(path), ctypes.byref(opts)))
        self.info = None

    def index(self):
        info = XvArkIndexInfo()
        _check(self.lib, self.lib.xv_ark_reader_index(self.handle, ctypes.byref(info)))
        self.info = info.as_dict()
        return self.info

    def set_first(self, marker_off, key_off):
        info = XvArkIndexInfo()
        _check(self.lib, self.lib.xv_ark_reader_set_first(self.handle, int(marker_off), int(key_off), ctypes.byref(info)))
        self.info = info.as_dict()
        return self.info

    def keys(self):
        """(blob, key_off): keys of the ok utterances, in order, as one bytes-like blob + int64 offsets [n_ok + 1]."""
        n = self.info["n_ok"]
        blob = np.empty(max(self.info["key_bytes"], 1), np.uint8)
        off = np.empty(n + 1, np.int64)
        _check(self.lib, self.lib.xv_ark_reader_keys(self.handle, blob.ctypes.data, self.info["key_bytes"], off.ctypes.data))
        return blob[:self.info["key_bytes"]], off

    def failures(self):
        """[(key, reason, rows)] of the skipped utterances, in order (reason: XV_UTT_ZERO_LENGTH = 1, XV_UTT_TOO_SHORT = 2)."""
        n = self.info["n_fail"]
        if n == 0:
            return []
        reason, rows, off = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n + 1, np.int64)
        blob = np.empty(n * 4097, np.uint8)
        _check(self.lib, self.lib.xv_ark_reader_failures(self.handle, reason.ctypes.data, rows.ctypes.data, blob.ctypes.data,
                                                         blob.shape[0], off.ctypes.data))
        raw = blob.tobytes()
        return [(raw[off[i]:off[i + 1]].decode(), int(reason[i]), int(rows[i])) for i in range(n)]

    def start(self, dst_row_base=0):
        _check(self.lib, self.lib.xv_ark_reader_start(self.handle, int(dst_row_base)))

    def next(self):
        """The next batch (blocks until its payloads are in memory), or None after the last one."""
        b = XvArkBatch()
        _check(self.lib, self.lib.xv_ark_reader_next(self.handle, ctypes.byref(b)))
        if b.n_utt == 0:
            return None
        out = ArkBatch()
        out.slot, out.n_seg, out.n_utt, out.n_rows, out.first_ok_index = b.slot, b.n_seg, b.n_utt, int(b.n_rows), int(b.first_ok_index)
        if b.feats_f16:                                # rows already rounded to float16 by the reader (feats_f16)
            out.feats = _view(b.feats, ctypes.c_uint16, out.n_rows * self.feat_dim, np.uint16).view(np.float16).reshape(out.n_rows, self.feat_dim)
        else:
            out.feats = _view(b.feats, ctypes.c_float, out.n_rows * self.feat_dim, np.float32).reshape(out.n_rows, self.feat_dim)
        out.seg_len = _view(b.seg_len, ctypes.c_int32, b.n_seg, np.int32)
        out.utt_first_seg = _view(b.utt_first_seg, ctypes.c_int32, b.n_utt + 1, np.int32)
        out.utt_dst_row = _view(b.utt_dst_row, ctypes.c_int64, b.n_utt, np.int64)
        return out

    def release(self, slot):
        _check(self.lib, self.lib.xv_ark_reader_release(self.handle, int(slot)))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.xv_ark_reader_close(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_mfcc(device, out_dev, utt_ids, lens, seed, stream=None):
    """Rows of the synthetic utterances ``utt_ids`` (lengths ``lens``) generated on the device into ``out_dev`` (float32 CUDA
    [>= sum(lens), feat_dim]); ``synthetic.counter_mfcc`` is the same recipe in numpy."""
    import torch
    lib = load_library()
    ids = np.ascontiguousarray(utt_ids, dtype=np.int64)
    ln = np.ascontiguousarray(lens, dtype=np.int32)
    assert out_dev.is_cuda and out_dev.dtype == torch.float32 and out_dev.is_contiguous() and out_dev.shape[0] >= int(ln.sum())
    s = torch.cuda.current_stream(device) if stream is None else stream
    _check(lib, lib.xv_synth_mfcc(int(device), out_dev.data_ptr(), ids.ctypes.data, ln.ctypes.data, int(ids.shape[0]),
                                  int(out_dev.shape[1]), int(seed) & (2 ** 64 - 1), s.cuda_stream))


def vec_ark_format(key_blob, key_off, vecs, with_markers=False, n_threads=4, out=None):
    """The bytes write_vec_flt (reference kaldi_io.py:309-343) emits for the float32 rows of ``vecs``, keyed by
    ``key_blob[key_off[i]:key_off[i+1]]`` -- one uint8 array -- and, with ``with_markers``, each entry's marker offset.
    ``out``: a writable uint8 array to format into (e.g. a window of the mmap'ed output file) instead of a fresh one."""
    lib = load_library()
    vecs = np.ascontiguousarray(vecs, dtype=np.float32)
    key_off = np.ascontiguousarray(key_off, dtype=np.int64)
    key_blob = np.ascontiguousarray(key_blob, dtype=np.uint8)
    n, dim = int(vecs.shape[0]), int(vecs.shape[1])
    assert key_off.shape[0] == n + 1
    total = int(lib.xv_vec_ark_bytes(key_off.ctypes.data, n, dim))
    if out is None:
        out = np.empty(max(total, 1), np.uint8)
    assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape[0] >= total
    markers = np.empty(max(n, 1), np.int64) if with_markers else None
    # key_off may be a window of a larger table: the formatter indexes the blob with the absolute offsets
    got = int(lib.xv_vec_ark_format(key_blob.ctypes.data, key_off.ctypes.data, n, vecs.ctypes.data, dim, out.ctypes.data, total,
                                    None if markers is None else markers.ctypes.data, int(n_threads)))
    if got < 0:
        _check(lib, got)
    return (out[:total], markers[:n]) if with_markers else out[:total]


def scp_format(key_blob, key_off, ark_name, base, markers):
    """The scp lines beside such an ark: ``key ark_name:offset`` (offset = base + marker)."""
    lib = load_library()
    key_off = np.ascontiguousarray(key_off, dtype=np.int64)
    key_blob = np.ascontiguousarray(key_blob, dtype=np.uint8)
    markers = np.ascontiguousarray(markers, dtype=np.int64)
    n = int(markers.shape[0])
    name = os.fsencode(ark_name)
    cap = int(key_off[n] - key_off[0]) + n * (len(name) + 24) + 1
    out = np.empty(cap, np.uint8)
    got = int(lib.xv_scp_format(key_blob.ctypes.data, key_off.ctypes.data, n, name, int(base), markers.ctypes.data, out.ctypes.data, cap))
    if got < 0:
        _check(lib, got)
    return out[:got]


class XvecEngine:
    """One xv_model on one CUDA device: the replacement for the reference's TF session +
    restored graph (models.py:365-366) on the extraction path."""

    ACTIVATIONS = {"relu": 0, "lrelu": 1, "prelu": 2}      # XV_ACT_* (include/xvec.h)
    POOLINGS = {"stats": 0, "attention": 1}                # XV_POOL_*

    def __init__(self, kernel_sizes, dilations, layer_sizes, emb_dim, feat_dim, device=0,
                 bn_eps=1e-3, var_eps=1e-5, activation="relu", pooling="stats"):
        self.lib = load_library()
        topo = XvTopology()
        topo.feat_dim = feat_dim
        topo.n_frame_layers = len(kernel_sizes)
        for i, (k, d, w) in enumerate(zip(kernel_sizes, dilations, layer_sizes)):
            topo.taps[i], topo.dilation[i], topo.width[i] = k, d, w
        topo.emb_dim = emb_dim
        topo.act = self.ACTIVATIONS[activation]
        topo.bn_eps = bn_eps
        topo.var_eps = var_eps
        topo.pooling = self.POOLINGS[pooling]
        self.pooling = pooling
        self.handle = ctypes.c_void_p()
        self.device = device
        self.feat_dim, self.emb_dim = feat_dim, emb_dim
        self.layer_sizes = list(layer_sizes)
        _check(self.lib, self.lib.xv_create(ctypes.byref(self.handle), device, ctypes.byref(topo)))
        self._ws = None

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.xv_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _check(self.lib, self.lib.xv_set_option(self.handle, name.encode(), int(value)))

    def set_params(self, params):
        """params: dict TF-variable-name -> array (the contract of load_model, models.py:199-210)."""
        for name, arr in params.items():
            a = np.ascontiguousarray(np.asarray(arr), dtype=np.float32)
            shape = (ctypes.c_int64 * a.ndim)(*a.shape)
            _check(self.lib, self.lib.xv_set_param(self.handle, name.encode(), a.ctypes.data_as(ctypes.c_void_p),
                                                   shape, a.ndim))

    def workspace_bytes(self, total_frames, n_seg):
        return int(self.lib.xv_workspace_bytes(self.handle, int(total_frames), int(n_seg)))

    def _workspace(self, nbytes):
        import torch
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device="cuda:%d" % self.device)
        return self._ws

    def forward(self, feats_dev, seg_lens, emb_dev=None, stream=None, return_layers=False):
        """feats_dev: torch float32 CUDA tensor [total_frames, feat_dim]; seg_lens: int32 host array.
        Enqueues on ``stream`` (default: torch's current stream); returns emb_dev [n_seg, emb_dim]."""
        import torch
        lens = np.ascontiguousarray(seg_lens, dtype=np.int32)
        n_seg, total = int(lens.shape[0]), int(lens.sum())
        assert feats_dev.is_cuda and feats_dev.dtype == torch.float32 and feats_dev.is_contiguous()
        assert feats_dev.shape[0] == total and feats_dev.shape[1] == self.feat_dim
        dev = feats_dev.device
        if emb_dev is None:
            emb_dev = torch.empty((n_seg, self.emb_dim), dtype=torch.float32, device=dev)
        ws = self._workspace(self.workspace_bytes(total, n_seg))
        s = torch.cuda.current_stream(dev) if stream is None else stream
        lens_p = lens.ctypes.data_as(ctypes.c_void_p)
        if not return_layers:
            _check(self.lib, self.lib.xv_forward(self.handle, feats_dev.data_ptr(), lens_p, n_seg, emb_dev.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), s.cuda_stream))
            return emb_dev
        layers = [torch.empty((total, w), dtype=torch.float32, device=dev) for w in self.layer_sizes]
        c_pool = self.layer_sizes[-1] // 2 if self.pooling == "attention" else self.layer_sizes[-1]
        stats = torch.empty((n_seg, 2 * c_pool), dtype=torch.float32, device=dev)
        ptrs = (ctypes.c_void_p * len(layers))(*[t.data_ptr() for t in layers])
        _check(self.lib, self.lib.xv_forward_layers(self.handle, feats_dev.data_ptr(), lens_p, n_seg,
                                                    emb_dev.data_ptr(), ws.data_ptr(), ws.numel(), s.cuda_stream,
                                                    ptrs, stats.data_ptr()))
        return emb_dev, layers, stats

    def check_overflow(self, stream=None):
        import torch
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        _check(self.lib, self.lib.xv_check_overflow(self.handle, s.cuda_stream))

    def rescue_overflow(self, stream=None):
        """For the enqueue-only calls (``forward`` / ``forward_utts``): 0 if nothing overflowed the fp16 range since the
        last check, else the number of per-layer scales that were raised -- run the forward again."""
        import torch
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        rc = int(self.lib.xv_rescue_overflow(self.handle, s.cuda_stream))
        if rc < 0:
            _check(self.lib, rc)
        return rc

    def extract_host(self, feats_host, seg_lens, emb_host=None):
        """Host in / host out (the reference's sess.run boundary).  feats_host: float32
        [total_frames, feat_dim] numpy array or CPU torch tensor (pinned for full PCIe speed)."""
        lens = np.ascontiguousarray(seg_lens, dtype=np.int32)
        n_seg = int(lens.shape[0])
        if hasattr(feats_host, "data_ptr"):
            fptr = feats_host.data_ptr()
            assert feats_host.is_contiguous() and feats_host.shape[0] == int(lens.sum())
        else:
            feats_host = np.ascontiguousarray(feats_host, dtype=np.float32)
            assert feats_host.shape[0] == int(lens.sum())
            fptr = feats_host.ctypes.data
        if emb_host is None:
            emb_host = np.empty((n_seg, self.emb_dim), dtype=np.float32)
        eptr = emb_host.data_ptr() if hasattr(emb_host, "data_ptr") else emb_host.ctypes.data
        _check(self.lib, self.lib.xv_extract_host(self.handle, fptr, lens.ctypes.data_as(ctypes.c_void_p), n_seg, eptr))
        return emb_host

    def submit_host(self, feats_host, seg_lens, emb_host):
        """Asynchronous extract_host: returns a ticket at once; ``collect(ticket)`` waits for it.  Both buffers
        must stay alive (and should be pinned) until collected; at most two submissions may be in flight."""
        lens = np.ascontiguousarray(seg_lens, dtype=np.int32)
        n_seg = int(lens.shape[0])
        if hasattr(feats_host, "data_ptr"):
            assert feats_host.is_contiguous() and feats_host.shape[0] == int(lens.sum())
            fptr = feats_host.data_ptr()
        else:
            assert feats_host.dtype == np.float32 and feats_host.flags.c_contiguous and feats_host.shape[0] == int(lens.sum())
            fptr = feats_host.ctypes.data
        eptr = emb_host.data_ptr() if hasattr(emb_host, "data_ptr") else emb_host.ctypes.data
        ticket = ctypes.c_int32(-1)
        _check(self.lib, self.lib.xv_submit_host(self.handle, fptr, lens.ctypes.data_as(ctypes.c_void_p), n_seg, eptr,
                                                 ctypes.byref(ticket)))
        return int(ticket.value)

    def collect(self, ticket):
        _check(self.lib, self.lib.xv_collect(self.handle, int(ticket)))

    # ---- utterance-level output (chunk average on the device; destination rows may be peer memory) ----
    @staticmethod
    def _utt_plan(n_seg, utt_first_seg, dst_rows):
        first = None if utt_first_seg is None else np.ascontiguousarray(utt_first_seg, dtype=np.int32)
        n_utt = n_seg if first is None else int(first.shape[0]) - 1
        dst = None if dst_rows is None else np.ascontiguousarray(dst_rows, dtype=np.int64)
        assert dst is None or dst.shape[0] == n_utt
        return first, dst, n_utt

    def forward_utts(self, feats_dev, seg_lens, out, utt_first_seg=None, dst_rows=None, stream=None):
        """``forward`` + the frame-weighted chunk average of make_embedding on the device.  ``out``: a float32 CUDA
        tensor, a ``PeerTable`` or a raw device address; row ``dst_rows[u]`` (default ``u``) receives utterance ``u``."""
        import torch
        lens = np.ascontiguousarray(seg_lens, dtype=np.int32)
        n_seg, total = int(lens.shape[0]), int(lens.sum())
        assert feats_dev.is_cuda and feats_dev.dtype == torch.float32 and feats_dev.is_contiguous()
        assert feats_dev.shape[0] == total and feats_dev.shape[1] == self.feat_dim
        first, dst, n_utt = self._utt_plan(n_seg, utt_first_seg, dst_rows)
        ws = self._workspace(self.workspace_bytes(total, n_seg))
        s = torch.cuda.current_stream(feats_dev.device) if stream is None else stream
        optr = out if isinstance(out, int) else out.data_ptr()
        _check(self.lib, self.lib.xv_forward_utts(self.handle, feats_dev.data_ptr(), lens.ctypes.data_as(ctypes.c_void_p), n_seg,
                                                  None if first is None else first.ctypes.data_as(ctypes.c_void_p),
                                                  None if dst is None else dst.ctypes.data_as(ctypes.c_void_p), n_utt,
                                                  optr, ws.data_ptr(), ws.numel(), s.cuda_stream))
        return n_utt

    def submit_dev_utts(self, feats_dev, seg_lens, utt_first_seg=None, dst_rows=None, out_dev=None, out_host=None, ready_event=None):
        """``submit_host_utts`` for features already on the device (a float32 CUDA tensor that stays untouched until collected).
        ``ready_event``: a ``torch.cuda.Event`` recorded behind the work that produces them on another stream."""
        lens = np.ascontiguousarray(seg_lens, dtype=np.int32)
        n_seg = int(lens.shape[0])
        assert feats_dev.is_cuda and feats_dev.is_contiguous() and feats_dev.shape[0] >= int(lens.sum())
        first, dst, n_utt = self._utt_plan(n_seg, utt_first_seg, dst_rows)
        optr = None if out_dev is None else (out_dev if isinstance(out_dev, int) else out_dev.data_ptr())
        hptr = None
        if out_host is not None:
            assert out_host.shape[0] >= n_utt
            hptr = out_host.data_ptr() if hasattr(out_host, "data_ptr") else out_host.ctypes.data
        ticket = ctypes.c_int32(-1)
        _check(self.lib, self.lib.xv_submit_dev_utts(self.handle, feats_dev.data_ptr(), lens.ctypes.data_as(ctypes.c_void_p), n_seg,
                                                     None if first is None else first.ctypes.data_as(ctypes.c_void_p),
                                                     None if dst is None else dst.ctypes.data_as(ctypes.c_void_p), n_utt,
                                                     optr, hptr, None if ready_event is None else ctypes.c_void_p(ready_event.cuda_event),
                                                     ctypes.byref(ticket)))
        return int(ticket.value)

    def submit_host_utts(self, feats_host, seg_lens, utt_first_seg=None, dst_rows=None, out_dev=None, out_host=None):
        """``submit_host`` with utterance-level output: averaged rows to ``out_dev[dst_rows[u]]`` (CUDA tensor, PeerTable
        or raw address; may be peer memory) and / or to the host array ``out_host[u]``."""
        lens = np.ascontiguousarray(seg_lens, dtype=np.int32)
        n_seg = int(lens.shape[0])
        half = False
        if hasattr(feats_host, "data_ptr"):
            assert feats_host.is_contiguous() and feats_host.shape[0] == int(lens.sum())
            fptr = feats_host.data_ptr()
        else:
            half = feats_host.dtype == np.float16      # rows the reader already rounded to float16 (what the device does first anyway)
            assert (half or feats_host.dtype == np.float32) and feats_host.flags.c_contiguous and feats_host.shape[0] == int(lens.sum())
            fptr = feats_host.ctypes.data
        first, dst, n_utt = self._utt_plan(n_seg, utt_first_seg, dst_rows)
        optr = None if out_dev is None else (out_dev if isinstance(out_dev, int) else out_dev.data_ptr())
        hptr = None
        if out_host is not None:
            assert out_host.shape[0] >= n_utt
            hptr = out_host.data_ptr() if hasattr(out_host, "data_ptr") else out_host.ctypes.data
        ticket = ctypes.c_int32(-1)
        fn = self.lib.xv_submit_host_utts_f16 if half else self.lib.xv_submit_host_utts
        _check(self.lib, fn(self.handle, fptr, lens.ctypes.data_as(ctypes.c_void_p), n_seg,
                                                      None if first is None else first.ctypes.data_as(ctypes.c_void_p),
                                                      None if dst is None else dst.ctypes.data_as(ctypes.c_void_p), n_utt,
                                                      optr, hptr, ctypes.byref(ticket)))
        return int(ticket.value)

    # ---- feature front end: apply-cmvn-sliding | select-voiced-frames on the device (include/xvec_frontend.h) ----
    def frontend(self, feats_dev, vad_dev, utt_lens, out_keep=None, opts=None, out_dev=None, stream=None):
        """feats_dev: float32 CUDA [sum(utt_lens), feat_dim] raw rows; vad_dev: float32 CUDA [sum(utt_lens)] or None;
        out_keep: selected rows to write per utterance (int32 host; required with a VAD track).  Enqueues on ``stream``;
        returns out_dev [sum(out_keep), feat_dim], laid out as the ``feats_dev`` of ``forward``."""
        import torch
        lens = np.ascontiguousarray(utt_lens, dtype=np.int32)
        n_utt, total = int(lens.shape[0]), int(lens.sum())
        assert feats_dev.is_cuda and feats_dev.dtype == torch.float32 and feats_dev.is_contiguous()
        assert feats_dev.shape[0] == total and feats_dev.shape[1] == self.feat_dim
        if vad_dev is not None:
            assert vad_dev.is_cuda and vad_dev.dtype == torch.float32 and vad_dev.is_contiguous() and vad_dev.numel() == total
            assert out_keep is not None, "out_keep (voiced rows to write per utterance) is required with a VAD track"
        keep = lens if out_keep is None else np.ascontiguousarray(out_keep, dtype=np.int32)
        assert keep.shape == lens.shape
        dev = feats_dev.device
        if out_dev is None:
            out_dev = torch.empty((int(keep.sum()), self.feat_dim), dtype=torch.float32, device=dev)
        assert out_dev.is_contiguous() and out_dev.shape[0] >= int(keep.sum())
        opts = XvCmvnOpts() if opts is None else opts
        need = int(self.lib.xv_frontend_workspace_bytes(self.handle, total, n_utt))
        if getattr(self, "_fe_ws", None) is None or self._fe_ws.numel() < need:
            self._fe_ws = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=dev)
        s = torch.cuda.current_stream(dev) if stream is None else stream
        _check(self.lib, self.lib.xv_frontend(self.handle, feats_dev.data_ptr(), None if vad_dev is None else vad_dev.data_ptr(),
                                              lens.ctypes.data_as(ctypes.c_void_p),
                                              None if out_keep is None else keep.ctypes.data_as(ctypes.c_void_p), n_utt,
                                              ctypes.byref(opts), out_dev.data_ptr(), self._fe_ws.data_ptr(),
                                              self._fe_ws.numel(), s.cuda_stream))
        return out_dev

    def submit_host_raw(self, feats_host, vad_host, utt_lens, out_keep, seg_lens, emb_host, opts=None):
        """``submit_host`` with the front end in front: raw rows + VAD track (host, should be pinned) in, embeddings of
        the ``seg_lens`` segments that tile the selected rows out.  Collect with ``collect(ticket)``."""
        lens = np.ascontiguousarray(utt_lens, dtype=np.int32)
        segs = np.ascontiguousarray(seg_lens, dtype=np.int32)
        total = int(lens.sum())

        def host_ptr(a, rows):
            if a is None:
                return None
            if hasattr(a, "data_ptr"):
                assert a.is_contiguous() and a.shape[0] == rows
                return a.data_ptr()
            assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape[0] == rows
            return a.ctypes.data

        keep = None if out_keep is None else np.ascontiguousarray(out_keep, dtype=np.int32)
        opts = XvCmvnOpts() if opts is None else opts
        eptr = emb_host.data_ptr() if hasattr(emb_host, "data_ptr") else emb_host.ctypes.data
        ticket = ctypes.c_int32(-1)
        _check(self.lib, self.lib.xv_submit_host_raw(self.handle, host_ptr(feats_host, total), host_ptr(vad_host, total),
                                                     lens.ctypes.data_as(ctypes.c_void_p),
                                                     None if keep is None else keep.ctypes.data_as(ctypes.c_void_p),
                                                     int(lens.shape[0]), ctypes.byref(opts),
                                                     segs.ctypes.data_as(ctypes.c_void_p), int(segs.shape[0]), eptr,
                                                     ctypes.byref(ticket)))
        return int(ticket.value)

    def last_kernel_ms(self):
        """Device duration (ms) of every launch of the last forward (option ``profile`` must be 1);
        order: pack, frame layers 0..n-1, pool+embed."""
        buf = (ctypes.c_float * 256)()
        n = int(self.lib.xv_last_kernel_ms(self.handle, buf, 256))
        if n < 0:
            _check(self.lib, n)
        return [float(buf[i]) for i in range(min(n, 256))]

    @property
    def last_launch_count(self):
        return int(self.lib.xv_last_launch_count(self.handle))


class PeerTable:
    """Rank 0's [rows, cols] float32 result table in device memory, mapped into every rank of a one-node job
    (xv_peer_alloc / xv_peer_open, include/xvec.h).  The owner creates it with ``PeerTable.create`` and hands ``handle``
    (64 bytes) to the other processes, which map it with ``PeerTable.open``; every rank then passes the table as the
    ``out`` of ``forward_utts`` / ``submit_host_utts`` and its rows land on rank 0 over NVLink."""

    def __init__(self, device, rows, cols, ptr, handle, owner):
        self.device, self.rows, self.cols, self.ptr, self.handle, self.owner = device, int(rows), int(cols), ptr, handle, owner
        self.lib = load_library()

    @classmethod
    def create(cls, device, rows, cols):
        lib = load_library()
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        _check(lib, lib.xv_peer_alloc(int(device), max(int(rows) * int(cols) * 4, 4), ctypes.byref(ptr), handle))
        return cls(device, rows, cols, int(ptr.value), bytes(handle), True)

    @classmethod
    def open(cls, device, rows, cols, handle):
        lib = load_library()
        ptr = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
        _check(lib, lib.xv_peer_open(int(device), buf, ctypes.byref(ptr)))
        return cls(device, rows, cols, int(ptr.value), bytes(handle), False)

    def data_ptr(self, row=0):
        return self.ptr + int(row) * self.cols * 4

    def read(self, row0=0, n_rows=None, out=None):
        """Blocking device -> host copy of rows [row0, row0 + n_rows) (the caller has synchronised with the writers)."""
        n_rows = self.rows - row0 if n_rows is None else int(n_rows)
        if out is None:
            out = np.empty((n_rows, self.cols), dtype=np.float32)
        if n_rows > 0:
            _check(self.lib, self.lib.xv_peer_read(self.device, out.ctypes.data_as(ctypes.c_void_p),
                                                   ctypes.c_void_p(self.data_ptr(row0)), n_rows * self.cols * 4))
        return out

    def close(self):
        if self.ptr:
            if self.owner:
                self.lib.xv_peer_free(self.device, ctypes.c_void_p(self.ptr))
            else:
                self.lib.xv_peer_close(self.device, ctypes.c_void_p(self.ptr))
            self.ptr = 0


TRAIN_PARAMS, TRAIN_ADAM_M, TRAIN_ADAM_V, TRAIN_MOVING, TRAIN_GRAD = 0, 1, 2, 3, 4


class XvecTrainer:
    """One xv_trainer bound to an XvecEngine: the replacement for the reference's TF session on the training path
    (``sess.run([optimizer, loss, accuracy])``, models.py:263).  State is addressed by TF variable name."""

    def __init__(self, engine, num_classes, emb1_dim=512):
        self.engine = engine
        self.lib = engine.lib
        self.num_classes = int(num_classes)
        self.handle = ctypes.c_void_p()
        _check(self.lib, self.lib.xv_train_create(ctypes.byref(self.handle), engine.handle, int(num_classes), int(emb1_dim)))
        self.n_params = int(self.lib.xv_train_size(self.handle, TRAIN_PARAMS))
        self.n_grad = int(self.lib.xv_train_size(self.handle, TRAIN_GRAD))     # gradient + the combined-overflow-flag tail
        self.seg_grad_offset = int(self.lib.xv_train_segment_grad_offset(self.handle))   # [this, n_params): segment-level gradients
        self.frame_grad_spans = []                       # (offset, count) of every frame layer's gradients (w | b | gamma | beta)
        for i in range(len(engine.layer_sizes)):
            off, cnt = ctypes.c_int64(0), ctypes.c_int64(0)
            _check(self.lib, self.lib.xv_train_frame_grad_span(self.handle, i, ctypes.byref(off), ctypes.byref(cnt)))
            self.frame_grad_spans.append((int(off.value), int(cnt.value)))
        self.n_moving = int(self.lib.xv_train_size(self.handle, TRAIN_MOVING))
        self._loss_acc = None
        self._geom = None

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.xv_train_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def span(self, name):
        which, off, cnt = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
        _check(self.lib, self.lib.xv_train_span(self.handle, name.encode(), ctypes.byref(which), ctypes.byref(off), ctypes.byref(cnt)))
        return int(which.value), int(off.value), int(cnt.value)

    def upload(self, which, arr, offset=0):
        a = np.ascontiguousarray(np.asarray(arr).reshape(-1), dtype=np.float32)
        _check(self.lib, self.lib.xv_train_upload(self.handle, int(which), a.ctypes.data_as(ctypes.c_void_p), int(offset), a.size))

    def download(self, which, offset=0, count=None):
        if count is None:
            count = int(self.lib.xv_train_size(self.handle, int(which))) - offset
        out = np.empty(int(count), dtype=np.float32)
        _check(self.lib, self.lib.xv_train_download(self.handle, int(which), out.ctypes.data_as(ctypes.c_void_p), int(offset), int(count)))
        return out

    def set_params(self, params, slots=None):
        """params: dict TF-variable-name -> array (trainable variables and moving statistics)."""
        for name, arr in params.items():
            which, off, cnt = self.span(name)
            a = np.asarray(arr)
            if a.size != cnt:
                raise XvecError(XV_EINVAL, "shape mismatch for %s: %d values, expected %d" % (name, a.size, cnt))
            self.upload(which, a, off)

    def get_param(self, name, which=None, shape=None):
        w, off, cnt = self.span(name)
        out = self.download(w if which is None else which, off, cnt)
        return out.reshape(shape) if shape is not None else out

    def set_option(self, name, value):
        _check(self.lib, self.lib.xv_train_set_option(self.handle, name.encode(), float(value)))

    @property
    def step(self):
        return int(self.lib.xv_train_get_step(self.handle))

    @step.setter
    def step(self, value):
        _check(self.lib, self.lib.xv_train_set_step(self.handle, int(value)))

    def forward_backward(self, feats_dev, labels_dev, n_seg, seg_len, grad_dev=None, stream=None, part=0):
        """feats_dev: torch float32 CUDA [n_seg*seg_len, feat_dim]; labels_dev: torch int32 CUDA [n_seg].
        Enqueues on ``stream``; returns the device tensor [loss, accuracy] (read it after a synchronize).
        ``part``: 0 the whole step; 1 / 2 its halves; ``PART_FRAME + i`` the slice of 2 that ends with frame layer i's
        gradients final (see ``forward_backward_allreduce``)."""
        import torch
        assert feats_dev.is_cuda and feats_dev.dtype == torch.float32 and feats_dev.is_contiguous()
        assert feats_dev.numel() == n_seg * seg_len * self.engine.feat_dim
        assert labels_dev.is_cuda and labels_dev.dtype == torch.int32 and labels_dev.numel() == n_seg
        if self._loss_acc is None:
            self._loss_acc = torch.zeros(2, dtype=torch.float32, device=feats_dev.device)
        s = torch.cuda.current_stream(feats_dev.device) if stream is None else stream
        gptr = None if grad_dev is None else grad_dev.data_ptr()
        _check(self.lib, self.lib.xv_train_forward_backward_part(self.handle, feats_dev.data_ptr(), labels_dev.data_ptr(), int(n_seg),
                                                                 int(seg_len), gptr, self._loss_acc.data_ptr(), s.cuda_stream, int(part)))
        self._geom = (int(n_seg), int(seg_len))
        return self._loss_acc

    PART_FRAME = 16                                      # XV_TRAIN_PART_FRAME (include/xvec_train.h)

    def forward_backward_allreduce(self, feats_dev, labels_dev, n_seg, seg_len, grad_dev, stream, comm_stream, fine=False):
        """Data-parallel step: gradient buckets are all-reduced on ``comm_stream`` as soon as they are final -- the
        segment-level gradients (60 % of the bytes) after the first half of the step; the frame-level ones behind the step, or
        with ``fine`` frame layer by frame layer from the top down, each under the backward of the layers below it.  (Measured
        on 8 B200: 1.094 ms/step with two buckets, 1.111 with the fine ones, 0.915 without any all-reduce: a collective's CTAs
        and the persistent all-SM grids of the backward do not share the SMs well whatever the bucketing)
        On return ``stream`` waits for all of them: ``apply`` may be enqueued."""
        import torch
        import torch.distributed as dist
        la = self.forward_backward(feats_dev, labels_dev, n_seg, seg_len, grad_dev=grad_dev, stream=stream, part=1)
        comm_stream.wait_stream(stream)
        with torch.cuda.stream(comm_stream):
            dist.all_reduce(grad_dev[self.seg_grad_offset:self.n_params])
        if fine:
            for i in reversed(range(len(self.frame_grad_spans))):
                self.forward_backward(feats_dev, labels_dev, n_seg, seg_len, grad_dev=grad_dev, stream=stream, part=self.PART_FRAME + i)
                off, cnt = self.frame_grad_spans[i]
                comm_stream.wait_stream(stream)
                with torch.cuda.stream(comm_stream):
                    dist.all_reduce(grad_dev[off:off + cnt])
        else:                                            # the frame level as ONE bucket behind the step (diagnostics)
            self.forward_backward(feats_dev, labels_dev, n_seg, seg_len, grad_dev=grad_dev, stream=stream, part=2)
            comm_stream.wait_stream(stream)
            with torch.cuda.stream(comm_stream):
                dist.all_reduce(grad_dev[:self.seg_grad_offset])
        with torch.cuda.stream(comm_stream):
            dist.all_reduce(grad_dev[self.n_params:])
        stream.wait_stream(comm_stream)
        return la

    def evaluate(self, feats_dev, labels_dev, n_seg, seg_len, stream=None):
        """Loss / accuracy with phase=False (moving statistics); nothing is updated."""
        import torch
        assert feats_dev.is_cuda and feats_dev.dtype == torch.float32 and feats_dev.is_contiguous()
        assert labels_dev.is_cuda and labels_dev.dtype == torch.int32 and labels_dev.numel() == n_seg
        if self._loss_acc is None:
            self._loss_acc = torch.zeros(2, dtype=torch.float32, device=feats_dev.device)
        s = torch.cuda.current_stream(feats_dev.device) if stream is None else stream
        _check(self.lib, self.lib.xv_train_eval(self.handle, feats_dev.data_ptr(), labels_dev.data_ptr(), int(n_seg), int(seg_len),
                                                self._loss_acc.data_ptr(), s.cuda_stream))
        return self._loss_acc

    def apply(self, learning_rate, grad_dev=None, grad_scale=1.0, stream=None):
        import torch
        s = torch.cuda.current_stream(self.engine.device) if stream is None else stream
        gptr = None if grad_dev is None else grad_dev.data_ptr()
        _check(self.lib, self.lib.xv_train_apply(self.handle, gptr, float(learning_rate), float(grad_scale), s.cuda_stream))

    def skipped_updates(self, stream=None, blocking=False):
        """Updates skipped so far because a loss-scaled fp16 gradient overflowed on any rank (never waits unless blocking)."""
        import torch
        s = torch.cuda.current_stream(self.engine.device) if stream is None else stream
        n = int(self.lib.xv_train_skipped_updates(self.handle, s.cuda_stream, 1 if blocking else 0))
        if n < 0:
            _check(self.lib, n)
        return n

    def convert_f16(self, src_dev, dst_dev, n, stream=None):
        """float16 CUDA tensor -> float32 CUDA tensor (first n values), on ``stream``."""
        import torch
        s = torch.cuda.current_stream(self.engine.device) if stream is None else stream
        _check(self.lib, self.lib.xv_convert_f16_to_f32(src_dev.data_ptr(), dst_dev.data_ptr(), int(n), s.cuda_stream))

    def sync_model(self):
        _check(self.lib, self.lib.xv_train_sync_model(self.handle))

    def debug_tensor(self, name, cols=None):
        n_seg, seg_len = self._geom
        cap = max(n_seg * seg_len * 1536, n_seg * max(self.num_classes, 3072)) + 16
        out = np.empty(cap, dtype=np.float32)
        n = int(self.lib.xv_train_debug_tensor(self.handle, name.encode(), out.ctypes.data_as(ctypes.c_void_p), cap))
        if n < 0:
            _check(self.lib, n)
        return out[:n].copy()

    def last_kernel_names(self):
        buf = ctypes.create_string_buffer(16384)
        n = int(self.lib.xv_train_last_kernel_names(self.handle, buf, 16384))
        if n < 0:
            _check(self.lib, n)
        return [x for x in buf.value.decode().split(";") if x]

    @property
    def last_launch_count(self):
        return int(self.lib.xv_train_last_launch_count(self.handle))
