#!/usr/bin/env python
"""GPU diagnostic of the training step: every stage of xv_train_forward_backward against the fp64 oracle.
usage: diag_train.py [topology] [weight_set] [B] [T] [classes]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xvector_oracle as orc                    # noqa: E402
from oracle import xvector_train_oracle as tro             # noqa: E402
from xvector_b200 import _native, synthetic                 # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)), float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    topology = sys.argv[1] if len(sys.argv) > 1 else "ModelWithoutDropoutTdnn"
    ws = sys.argv[2] if len(sys.argv) > 2 else "B"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 120
    NC = int(sys.argv[5]) if len(sys.argv) > 5 else 700
    topo = orc.TOPOLOGIES[topology]
    P = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], num_classes=NC, weight_set=ws)
    x = synthetic.mfcc(5, B * T).reshape(B, T, 23)
    labels = np.random.default_rng(5).integers(0, NC, B).astype(np.int32)
    ref = tro.forward_backward(x, labels, P, topology, return_intermediates=True, fp16_storage=bool(os.environ.get("FP16_ORACLE")),
                               bn_eps=float(os.environ.get("BN_EPS", "1e-3")))
    print("diag_train: %s set %s B=%d T=%d classes=%d  oracle loss %.6f acc %.3f" % (topology, ws, B, T, NC, ref["loss"], ref["accuracy"]))

    eng = _native.XvecEngine(topo["kernel_sizes"], topo["dilations"], topo["layer_sizes"], 512, 23, device=0, bn_eps=float(os.environ.get("BN_EPS", "1e-3")))
    tr = _native.XvecTrainer(eng, NC, 512)
    tr.set_params(P)
    feats = torch.from_numpy(x.reshape(B * T, 23)).cuda()
    lab = torch.from_numpy(labels).cuda()
    variants = [(None, None)]
    if os.environ.get("WGRAD_SWEEP"):
        variants += [(16384, 1024), (1024, 16384), (128, 1024), (1024, 128), (16384, 128)]
    for lbo, sbo in variants:
        if lbo is not None:
            tr.set_option("wgrad_lbo", lbo); tr.set_option("wgrad_sbo", sbo)
            print("  --- wgrad descriptor lbo=%d sbo=%d" % (lbo, sbo))
        la = tr.forward_backward(feats, lab, B, T)
        torch.cuda.synchronize()
        try:
            eng.check_overflow()
            print("  overflow: none")
        except Exception as e:                                     # noqa: BLE001
            print("  overflow:", e)
        la = la.cpu().numpy()
        print("  loss %.6f (oracle %.6f, rel %.2e)  accuracy %.4f (oracle %.4f)  launches %d" %
              (la[0], ref["loss"], abs(la[0] - ref["loss"]) / abs(ref["loss"]), la[1], ref["accuracy"], tr.last_launch_count))
        S = 1.0
        while S < 8.0 * B * T:
            S *= 2.0
        inter, ig = ref["intermediates"], ref["intermediate_grads"]
        if lbo is None:
            for i in range(5):
                s = "frame_level_info_layer-%d/" % i
                C = topo["layer_sizes"][i]
                r = tr.debug_tensor("r%d" % i).reshape(B, T, C)
                print("  layer %d  relu out   l2 %.2e max %.2e" % ((i,) + rel(r, inter[s + "relu"])))
                if i < 4:
                    y = tr.debug_tensor("y%d" % i).reshape(B, T, C)
                    print("  layer %d  bn out     l2 %.2e max %.2e" % ((i,) + rel(y, inter[s + "bn"])))
            print("  h0 (stats)         l2 %.2e max %.2e" % rel(tr.debug_tensor("h0").reshape(B, -1), inter["stats"]))
            print("  embed-0 scores     l2 %.2e max %.2e" % rel(tr.debug_tensor("z5").reshape(B, -1), inter["embed_layer-0/scores"]))
            print("  embed-1 bn         l2 %.2e max %.2e" % rel(tr.debug_tensor("y6").reshape(B, -1), inter["embed_layer-1/bn"]))
            print("  logits             l2 %.2e max %.2e" % rel(tr.debug_tensor("logits").reshape(B, -1), ref["logits"]))
            print("  dh0                l2 %.2e max %.2e" % rel(tr.debug_tensor("dh0").reshape(B, -1), ig["stats"]))
            for i in range(4, -1, -1):
                s = "frame_level_info_layer-%d/" % i
                C = topo["layer_sizes"][i]
                dz = tr.debug_tensor("dz%d" % i).reshape(B, T, C) / S
                want = ig[s + "relu"] * (inter[s + "relu"] > 0)
                print("  layer %d  dz         l2 %.2e max %.2e   (|dz*S| max %.3g)" % ((i,) + rel(dz, want) + (np.abs(dz).max() * S,)))
                if i < 4:
                    dy = tr.debug_tensor("dy%d" % i).reshape(B, T, C) / S
                    print("  layer %d  dy         l2 %.2e max %.2e" % ((i,) + rel(dy, ig[s + "bn"])))
        worst = 0.0
        for name in tro.trainable_names(topo, P):
            g = tr.get_param(name, which=_native.TRAIN_GRAD)
            e = rel(g, ref["grads"][name])
            worst = max(worst, e[0])
            print("  grad %-36s l2 %.2e max %.2e" % ((name,) + e))
        print("  worst gradient l2 error %.3e" % worst)
        for s, (mean, var) in ref["batch_stats"].items():
            mm = tr.get_param(s + "mean:0"); mv = tr.get_param(s + "variance:0")
            e1 = rel(mm, ref["moving"][s + "mean:0"]); e2 = rel(mv, ref["moving"][s + "variance:0"])
            print("  moving %-28s mean l2 %.2e  var l2 %.2e" % (s, e1[0], e2[0]))
        tr.set_params({k: v for k, v in P.items() if k.endswith("mean:0") or k.endswith("variance:0")})
    # Adam: 3 steps against the oracle's fp64 Adam on the same minibatch
    names = tro.trainable_names(topo, P)
    Pn = {k: np.asarray(v, np.float64) for k, v in P.items()}
    slots = tro.adam_init(Pn, names)
    tr.set_params(P)
    tr.set_option("wgrad_lbo", 16384); tr.set_option("wgrad_sbo", 1024)
    for it in range(3):
        o = tro.forward_backward(x, labels, Pn, topology)
        tro.adam_step(Pn, o["grads"], slots, 1e-3)
        for k, v in o["moving"].items():
            Pn[k] = v
        la = tr.forward_backward(feats, lab, B, T)
        tr.apply(1e-3)
        torch.cuda.synchronize()
        print("  step %d loss gpu %.6f oracle %.6f" % (it + 1, float(la[0]), o["loss"]))
    worst = 0.0
    for name in names:
        e = rel(tr.get_param(name) - np.asarray(P[name]).ravel(), (Pn[name] - np.asarray(P[name], np.float64)).ravel())
        worst = max(worst, e[0])
    print("  after 3 Adam steps: worst relative error of the parameter UPDATE %.3e" % worst)
    tr.close(); eng.close()


if __name__ == "__main__":
    main()
