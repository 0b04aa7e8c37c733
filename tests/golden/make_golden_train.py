#!/usr/bin/env python
"""Regenerates tests/golden/train_golden.npz: loss, accuracy, per-variable gradient norms and first values, and the updated
moving statistics of ONE minibatch (B=4, T=40, 30 classes, seeds fixed) as computed by oracle/xvector_train_oracle.py (fp64).

TensorFlow is not available here and the reference ships no fixtures, so this file does not pin the oracle
against the reference; it pins the oracle against ITSELF across refactorings (tests/test_train_oracle.py::test_golden_minibatch),
next to the finite-difference check of its gradients.
usage: python tests/golden/make_golden_train.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import xvector_oracle as orc                      # noqa: E402
from oracle import xvector_train_oracle as tro               # noqa: E402
from xvector_b200 import synthetic                           # noqa: E402


def compute(topology):
    topo = orc.TOPOLOGIES[topology]
    P = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], num_classes=30, weight_set="B",
                              activation=topo.get("act", "relu"))
    x = synthetic.mfcc(7, 4 * 40).reshape(4, 40, 23)
    labels = np.random.default_rng(7).integers(0, 30, 4)
    out = tro.forward_backward(x, labels, P, topology)
    rec = {"loss": np.float64(out["loss"]), "accuracy": np.float64(out["accuracy"])}
    for name, g in out["grads"].items():
        rec["gnorm/" + name] = np.float64(np.linalg.norm(g))
        rec["ghead/" + name] = np.asarray(g, np.float64).ravel()[:4].copy()
    for name, v in out["moving"].items():
        rec["moving/" + name] = np.asarray(v, np.float64)[:4].copy()
    return rec


if __name__ == "__main__":
    blob = {}
    for topology in ("ModelWithoutDropoutTdnn", "ModelWithoutDropout", "ModelL2LossWithoutDropoutLRelu"):
        for k, v in compute(topology).items():
            blob[topology + "|" + k] = v
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "train_golden.npz"), **blob)
    print("wrote %d entries" % len(blob))
