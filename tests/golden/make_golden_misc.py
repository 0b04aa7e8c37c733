#!/usr/bin/env python
"""Golden fixtures for the integer-vector / posterior / confusion-network-time / segments readers, made with the
REFERENCE's own kaldi_io module (local/tf/kaldi_io.py, pure numpy: imports under Python 3).  Build container only
(/root/reference does not exist on the GPU box):  ``python tests/golden/make_golden_misc.py``.

    vec_int.ark            written by reference write_vec_int                     (kaldi_io.py:189-222)
    post.ark, cntime.ark   hand-assembled binary entries (the reference has no writer for them)
    segments.txt           a Kaldi segments file of one recording
    misc_expected.npz      what the reference's readers return for each of them   (kaldi_io.py:163-187, :581-617, :643-673, :680-700)
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/local/tf")
import kaldi_io as ref_io  # noqa: E402  (the reference module)


def main():
    rng = np.random.Generator(np.random.PCG64(11))
    expected = {}
    ints = {"utt-a": rng.integers(-5, 5000, 17).astype(np.int32), "utt.b": np.array([7], np.int32),
            "spk/utt_c": rng.integers(0, 2 ** 31 - 1, 300).astype(np.int32)}
    with open(os.path.join(HERE, "vec_int.ark"), "wb") as f:
        for k, v in ints.items():
            ref_io.write_vec_int(f, v, key=k)
    for k, v in ref_io.read_vec_int_ark(os.path.join(HERE, "vec_int.ark")):
        expected["vec_int/" + k] = np.asarray(v)
        assert np.array_equal(v, ints[k])

    posts = {"p1": [[(3, 0.25), (17, 0.75)], [(0, 1.0)], [(5, 0.5), (6, 0.25), (7, 0.25)]], "p2": [[(42, 1.0)]]}
    with open(os.path.join(HERE, "post.ark"), "wb") as f:
        for k, post in posts.items():
            f.write((k + " ").encode() + b"\0B\4" + struct.pack("<i", len(post)))
            for frame in post:
                f.write(b"\4" + struct.pack("<i", len(frame)))
                for idx, val in frame:
                    f.write(b"\4" + struct.pack("<i", idx) + b"\4" + struct.pack("<f", val))
    for k, post in ref_io.read_post_ark(os.path.join(HERE, "post.ark")):
        expected["post/" + k + "/lens"] = np.array([len(fr) for fr in post], np.int64)
        expected["post/" + k + "/idx"] = np.array([i for fr in post for i, _ in fr], np.int64)
        expected["post/" + k + "/val"] = np.array([v for fr in post for _, v in fr], np.float64)

    times = {"c1": [(0.0, 0.31), (0.31, 0.5), (0.5, 1.25)], "c2": [(2.5, 2.75)]}
    with open(os.path.join(HERE, "cntime.ark"), "wb") as f:
        for k, bins in times.items():
            f.write((k + " ").encode() + b"\0B\4" + struct.pack("<i", len(bins)))
            for b, e in bins:
                f.write(b"\4" + struct.pack("<f", b) + b"\4" + struct.pack("<f", e))
    for k, bins in ref_io.read_cntime_ark(os.path.join(HERE, "cntime.ark")):
        expected["cntime/" + k] = np.array(bins, np.float64)

    with open(os.path.join(HERE, "segments.txt"), "wt") as f:
        f.write("rec1-0001 rec1 0.10 0.55\nrec1-0002 rec1 0.80 1.20\nrec1-0003 rec1 1.20 1.37\n")
    expected["segments"] = np.asarray(ref_io.read_segments_as_bool_vec(os.path.join(HERE, "segments.txt")))
    np.savez(os.path.join(HERE, "misc_expected.npz"), **expected)
    print("wrote", sorted(expected))


if __name__ == "__main__":
    main()
