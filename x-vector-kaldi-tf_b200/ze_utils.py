"""The slice of the reference's ``local/tf/ze_utils.py`` that the extraction path touches.

  * ``set_cuda_visible_devices`` (reference ze_utils.py:25-48) -- the reference picks a free GPU
    by parsing ``nvidia-smi`` or *hides* every GPU when ``use_gpu`` is false so TensorFlow runs
    on the CPU.  This build has no CPU path: the device is chosen per process from
    ``LOCAL_RANK`` (torchrun) / ``XVEC_DEVICE`` and ``use_gpu=False`` only logs a note.
  * ``is_correct_model_dir`` (ze_utils.py:561-567) -- unchanged contract: non-empty
    ``model.meta`` and ``done``.
  * ``wait_for_background_commands`` (reference ze_utils.py:240-247, called by extract_embedding.py:157) -- joins the
    clean-up threads of the shell pipes kaldi_io.popen started and, unlike the reference, fails the caller when a child
    exited non-zero (an exception inside a thread cannot).
"""
from __future__ import annotations

import os


def pick_device():
    """CUDA device index for this process: XVEC_DEVICE, else LOCAL_RANK, else 0."""
    for var in ("XVEC_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var, "").strip()
        if v.isdigit():
            return int(v)
    return 0


def set_cuda_visible_devices(use_gpu=True, logger=None):
    if not use_gpu and logger is not None:
        # the reference's own wrapper passes --use-gpu=no (extract_xvectors.sh:72-79): a drop-in run changes device here
        logger.warning("--use-gpu=no was requested, but this build has no CPU path: "
                       "running the sm_100a kernels on CUDA device %d." % pick_device())
    elif logger is not None:
        logger.info("Using CUDA device %d" % pick_device())
    return pick_device()


def is_correct_model_dir(model_dir):
    model_file = os.path.join(model_dir, "model.meta")
    done_file = os.path.join(model_dir, "done")
    return (os.path.isfile(model_file) and os.stat(model_file).st_size > 0 and
            os.path.isfile(done_file) and os.stat(done_file).st_size > 0)


def wait_for_background_commands():
    from . import kaldi_io
    kaldi_io.wait_for_children()
