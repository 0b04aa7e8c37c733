#!/usr/bin/env python
"""Generate the Kaldi-ark golden fixtures with the REFERENCE's own kaldi_io module.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):  ``python tests/golden/make_golden_ark.py``.  The reference module
(local/tf/kaldi_io.py) is pure numpy, so it imports under Python 3.  Outputs (committed):

    mat_f32.ark, mat_f64.ark   matrices written by reference write_mat      (kaldi_io.py:506-542)
    vec_f32.ark, vec_f64.ark   vectors  written by reference write_vec_flt  (kaldi_io.py:309-343)
    cm.ark + cm_expected.npy   hand-assembled 'CM ' compressed matrix and what the reference's
                               reader decodes it to                          (kaldi_io.py:455-502)
    text.ark + text_expected.npy   text-form matrix and the reference's parse (kaldi_io.py:440-452)
    expected.npz               the source arrays, keyed "<file>/<key>"
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/local/tf")
import kaldi_io as ref_io  # noqa: E402  (the reference module)


def main():
    rng = np.random.Generator(np.random.PCG64(7))
    expected = {}

    mats = {"utt-a_01": rng.standard_normal((7, 23)).astype(np.float32),
            "spk1/utt.b": rng.standard_normal((1, 23)).astype(np.float32),
            "empty_utt": np.zeros((0, 23), np.float32),
            "Z9": rng.standard_normal((30, 5)).astype(np.float32)}
    with open(os.path.join(HERE, "mat_f32.ark"), "wb") as f:
        for k, m in mats.items():
            ref_io.write_mat(f, m, key=k)
            expected["mat_f32.ark/" + k] = m
    with open(os.path.join(HERE, "mat_f64.ark"), "wb") as f:
        for k, m in mats.items():
            ref_io.write_mat(f, m.astype(np.float64), key=k)
            expected["mat_f64.ark/" + k] = m.astype(np.float64)

    vecs = {"utt-a_01": rng.standard_normal(512).astype(np.float32),
            "x": rng.standard_normal(3).astype(np.float32)}
    with open(os.path.join(HERE, "vec_f32.ark"), "wb") as f:
        for k, v in vecs.items():
            ref_io.write_vec_flt(f, v, key=k)
            expected["vec_f32.ark/" + k] = v
    with open(os.path.join(HERE, "vec_f64.ark"), "wb") as f:
        for k, v in vecs.items():
            ref_io.write_vec_flt(f, v.astype(np.float64), key=k)
            expected["vec_f64.ark/" + k] = v.astype(np.float64)

    # compressed matrix: assemble bytes by hand (the reference has no CM writer), decode with the reference
    rows, cols = 9, 4
    body = struct.pack("<ffii", -3.5, 11.25, rows, cols)
    pct = np.sort(rng.integers(0, 65536, size=(cols, 4)), axis=1).astype("<u2")
    data = rng.integers(0, 256, size=(cols, rows)).astype(np.uint8)
    data[0, :4] = [0, 64, 192, 255]                      # segment boundaries
    data[1, :4] = [65, 193, 1, 254]
    blob = b"cm_utt \0BCM " + body + pct.tobytes() + data.tobytes()
    with open(os.path.join(HERE, "cm.ark"), "wb") as f:
        f.write(blob)
    got = {k: m for k, m in ref_io.read_mat_ark(os.path.join(HERE, "cm.ark"))}
    np.save(os.path.join(HERE, "cm_expected.npy"), np.ascontiguousarray(got["cm_utt"]))

    # text matrix
    with open(os.path.join(HERE, "text.ark"), "wb") as f:
        f.write(b"t1  [\n  1.5 -2 3e-2\n  4 5 6.25 ]\n")
    got = {k: m for k, m in ref_io.read_mat_ark(os.path.join(HERE, "text.ark"))}
    np.save(os.path.join(HERE, "text_expected.npy"), got["t1"])

    np.savez(os.path.join(HERE, "expected.npz"), **expected)

    # self-check: the reference reads back what it wrote
    for k, m in ref_io.read_mat_ark(os.path.join(HERE, "mat_f32.ark")):
        assert np.array_equal(m, mats[k])
    print("golden ark fixtures written to", HERE)


if __name__ == "__main__":
    main()
