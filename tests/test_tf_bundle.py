"""TensorFlow checkpoint-bundle reader (x-vector-kaldi-tf_b200/tf_bundle.py): primitives against published known answers,
round trips through the writer, corruption detection, and Model.load_model on a reference-style checkpoint directory."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from xvector_b200 import models, synthetic, tf_bundle       # noqa: E402


def test_crc32c_known_answers():
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283                 # the standard CRC-32C check value
    assert tf_bundle.crc32c(bytes(32)) == 0x8A9136AA                     # RFC 3720 B.4 test vectors
    assert tf_bundle.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tf_bundle.crc32c(bytes(range(32))) == 0x46DD794E
    c = tf_bundle.crc32c(b"foo")
    assert tf_bundle.mask_crc(c) != c and tf_bundle.unmask_crc(tf_bundle.mask_crc(c)) == c
    assert tf_bundle.crc32c(b"world", tf_bundle.crc32c(b"hello ")) == tf_bundle.crc32c(b"hello world")


def test_varint_known_answers():
    assert tf_bundle.write_varint(0) == b"\x00" and tf_bundle.write_varint(127) == b"\x7f"
    assert tf_bundle.write_varint(300) == b"\xac\x02"                     # protobuf documentation example
    for v in (0, 1, 127, 128, 300, 2 ** 31, 2 ** 40 + 17):
        assert tf_bundle.read_varint(tf_bundle.write_varint(v) + b"\x99", 0) == (v, len(tf_bundle.write_varint(v)))


def test_bundle_entry_proto_hand_assembled():
    # dtype = DT_FLOAT(1); shape {dim{size 5} dim{size 23} dim{size 512}}; offset 4096; size 235520; crc32c fixed32
    buf = (b"\x08\x01" + b"\x12\x0d" + b"\x12\x02\x08\x05" + b"\x12\x02\x08\x17" + b"\x12\x03\x08\x80\x04" +
           b"\x20\x80\x20" + b"\x28\x80\xb0\x0e" + b"\x35" + struct.pack("<I", 0xDEADBEEF))
    e = tf_bundle._parse_entry(buf)
    assert e["dtype"] == 1 and e["shape"] == [5, 23, 512] and e["offset"] == 4096 and e["size"] == 235520
    assert e["crc32c"] == 0xDEADBEEF and e["shard_id"] == 0 and not e["sliced"]


def _checkpoint(tmp_path, cls=models.ModelWithoutDropoutTdnn, num_classes=40):
    P = synthetic.make_params(cls.kernel_sizes, cls.layer_sizes, cls.embedding_sizes, num_classes=num_classes, weight_set="B")
    P["beta1_power:0"] = np.float32(0.9 ** 7)                             # scalar, as TF's Adam stores it
    P["frame_level_info_layer-1/w/Adam:0"] = np.full_like(P["frame_level_info_layer-1/w:0"], 1e-3)
    d = tmp_path / "model_final"
    d.mkdir()
    tf_bundle.write_bundle(str(d / "model"), P)
    (d / "model.meta").write_bytes(b"\x0a\x8f\x01\x0a\x0bPlaceholder\xff\xfe not json")     # a MetaGraph is binary protobuf
    (d / "done").write_text("done")
    return str(d), P


def test_round_trip_is_exact_and_checks_data(tmp_path):
    d, P = _checkpoint(tmp_path)
    got = tf_bundle.read_bundle(os.path.join(d, "model"), verify_data=True)
    assert set(got) == set(P)
    for k in P:
        assert got[k].dtype == np.float32 and got[k].shape == np.asarray(P[k]).shape, k
        np.testing.assert_array_equal(got[k], P[k])
    idx = tf_bundle.read_index(os.path.join(d, "model.index"))
    assert "" in idx and "frame_level_info_layer-0/w" in idx and "beta1_power" in idx     # names carry no ':0' on disk
    size = os.path.getsize(os.path.join(d, "model.data-00000-of-00001"))
    assert size == sum(np.asarray(v).nbytes for v in P.values())


def test_corruption_is_detected(tmp_path):
    d, _ = _checkpoint(tmp_path)
    path = os.path.join(d, "model.index")
    blob = bytearray(open(path, "rb").read())
    blob[20] ^= 0x40
    open(path, "wb").write(bytes(blob))
    with pytest.raises(ValueError, match="checksum"):
        tf_bundle.read_index(path)
    open(path, "wb").write(bytes(blob[:-3]))
    with pytest.raises(ValueError, match="magic"):
        tf_bundle.read_index(path)
    d2, _ = _checkpoint(tmp_path / "b")  if (tmp_path / "b").mkdir() is None else (None, None)
    data = os.path.join(d2, "model.data-00000-of-00001")
    raw = bytearray(open(data, "rb").read())
    raw[100] ^= 0x01
    open(data, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="data checksum"):
        tf_bundle.read_bundle(os.path.join(d2, "model"), verify_data=True)


def test_load_model_reads_a_reference_checkpoint_directory(tmp_path):
    d, P = _checkpoint(tmp_path)
    m = models.ModelWithoutDropoutTdnn()
    m.load_model(None, d, None)
    assert m.meta["source"] == "tensorflow-bundle" and m.num_classes == 40 and m.meta["input_feature_dim"] == 23
    assert m.kernel_sizes == [5, 3, 3, 1, 1] and m.dilation_rates == [1, 2, 3, 1, 1] and m.layer_sizes == [512, 512, 512, 512, 1536]
    np.testing.assert_array_equal(m.params["embed_layer-0/w:0"], P["embed_layer-0/w:0"])
    with pytest.raises(RuntimeError, match="kernel sizes"):
        models.ModelWithoutDropout().load_model(None, d, None)              # dense topology: 5,5,7,1,1
    empty = tmp_path / "empty"
    empty.mkdir()
    (empty / "model.meta").write_bytes(b"\x0a\x01\xff")
    with pytest.raises(RuntimeError, match="no model.index"):
        models.Model().load_model(None, str(empty), None)
