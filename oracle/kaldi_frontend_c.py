"""ctypes binding of oracle/kaldi_frontend_oracle.c (plain-C restatement of the Kaldi feature pipe).  TEST INFRASTRUCTURE
ONLY, like everything under oracle/.  The shared object is built with gcc into oracle/_build/ (git-ignored; it travels to
the GPU box with the snapshot, and is rebuilt there on demand if absent)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kaldi_frontend_oracle.c")
LIB = os.path.join(HERE, "_build", "libkaldi_frontend_oracle.so")
_lib = None


def build(verbose=False):
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        P, I = ctypes.c_void_p, ctypes.c_int32
        lib.kfo_sliding_window_cmn.argtypes = [P, I, I, I, I, I, I, P]
        lib.kfo_sliding_window_cmn.restype = ctypes.c_int
        lib.kfo_select_voiced_frames.argtypes = [P, P, I, I, P]
        lib.kfo_select_voiced_frames.restype = I
        _lib = lib
    return _lib


def sliding_window_cmn(feats, cmn_window=300, center=True, normalize_variance=False, min_window=100):
    x = np.ascontiguousarray(feats, dtype=np.float32)
    out = np.empty_like(x)
    rc = _load().kfo_sliding_window_cmn(x.ctypes.data, x.shape[0], x.shape[1], int(cmn_window), int(min_window), int(bool(center)),
                                        int(bool(normalize_variance)), out.ctypes.data)
    assert rc == 0
    return out


def frontend(feats, vad, cmn_window=300, center=True, normalize_variance=False, min_window=100):
    """apply-cmvn-sliding | select-voiced-frames for one utterance (None = the utterance is skipped)."""
    cm = sliding_window_cmn(feats, cmn_window, center, normalize_variance, min_window)
    if vad is None:
        return cm
    vad = np.ascontiguousarray(vad, dtype=np.float32)
    if vad.shape[0] != cm.shape[0] or not np.any(vad != 0):
        return None
    out = np.empty_like(cm)
    kept = _load().kfo_select_voiced_frames(cm.ctypes.data, vad.ctypes.data, cm.shape[0], cm.shape[1], out.ctypes.data)
    return out[:kept]
