"""Reader (and a minimal writer) for TensorFlow checkpoint bundles, without TensorFlow.

The reference saves its models with ``tf.train.Saver`` (local/tf/models.py:131-141): ``model.index`` +
``model.data-00000-of-00001`` (+ ``model.meta``, the MetaGraph, which this build does not need: the topology comes from the
Model subclass and the variable shapes).  ``load_model`` restores variables BY NAME (models.py:143-162); this module gives
the same name -> array mapping so that an existing ``exp/xvector_nnet_*/model_final`` can be loaded directly.

Format (TensorFlow ``tensor_bundle`` / LevelDB table, public specifications):

  ``*.index`` is an immutable sorted string table: data blocks of prefix-compressed entries
  ``varint32 shared | varint32 non_shared | varint32 value_len | key_delta | value`` followed by a restart array
  (``uint32[n] | uint32 n``) and a 5-byte trailer (compression type, masked crc32c); an index block maps the last key of each
  data block to its BlockHandle (``varint64 offset | varint64 size``); a 48-byte footer holds the metaindex and index handles
  and the magic ``0xdb4775248b80fb57``.  Key ``""`` maps to a ``BundleHeaderProto`` (fields: num_shards = 1, endianness = 2, version = 3), every
  other key is a variable name (no ``:0`` suffix) mapping to a ``BundleEntryProto``:
  ``dtype = 1; shape = 2 {dim = 2 {size = 1}}; shard_id = 3; offset = 4; size = 5; crc32c = 6 (fixed32)``.
  ``*.data-SSSSS-of-NNNNN`` holds the raw little-endian tensor bytes at ``offset``.

PARITY STATUS: unpinned against TensorFlow-written files (TensorFlow is absent from the build environment and the reference
ships no checkpoint); pinned by round trips through the writer below, which follows the same specification, and by
known-answer tests of the varint / crc32c / protobuf primitives (tests/test_tf_bundle.py).
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 19: np.float16, 4: np.uint8, 6: np.int8, 10: np.bool_}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------- primitives
def _crc32c_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TABLE = _crc32c_table()


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), the checksum of LevelDB blocks and bundle tensors."""
    c = crc ^ 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in bytes(data):
        c = tbl[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xa282ead8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


def read_varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def write_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """Flat protobuf decode: list of (field_number, wire_type, value)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = read_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = read_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _parse_entry(buf):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for field, _, v in _parse_proto(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            for f2, _, dim in _parse_proto(v):
                if f2 == 2:
                    size = 0
                    for f3, _, dv in _parse_proto(dim):
                        if f3 == 1:
                            size = dv if dv < (1 << 63) else dv - (1 << 64)
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["sliced"] = True
    return e


# ---------------------------------------------------------------- table reader
def _read_block(data, offset, size, verify=True):
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        actual = crc32c(data[offset:offset + size + 1])
        if unmask_crc(stored) != actual:
            raise ValueError("index block checksum mismatch at offset %d" % offset)
    if ctype != 0:
        raise ValueError("compressed table blocks (type %d) are not supported" % ctype)
    n_restarts = struct.unpack_from("<I", raw, len(raw) - 4)[0]
    limit = len(raw) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < limit:
        shared, pos = read_varint(raw, pos)
        non_shared, pos = read_varint(raw, pos)
        vlen, pos = read_varint(raw, pos)
        key = key[:shared] + bytes(raw[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(raw[pos:pos + vlen])))
        pos += vlen
    return out


def read_index(index_path, verify=True):
    """name -> BundleEntry dict, plus the header under key ''."""
    with open(index_path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % index_path)
    footer = data[-48:]
    pos = 0
    _, pos = read_varint(footer, pos)          # metaindex handle
    _, pos = read_varint(footer, pos)
    idx_off, pos = read_varint(footer, pos)
    idx_size, pos = read_varint(footer, pos)
    entries = {}
    for _, handle in _read_block(data, idx_off, idx_size, verify):
        boff, p2 = read_varint(handle, 0)
        bsize, _ = read_varint(handle, p2)
        for key, value in _read_block(data, boff, bsize, verify):
            entries[key.decode()] = value
    return entries


def read_bundle(prefix, verify_data=False, add_suffix=":0"):
    """All tensors of the checkpoint ``prefix`` (e.g. ``.../model_final/model``): dict name(+':0') -> numpy array."""
    raw = read_index(prefix + ".index")
    header = raw.pop("", None)
    num_shards = 1
    if header is not None:
        for field, _, v in _parse_proto(header):
            if field == 1:
                num_shards = v
            elif field == 2 and v not in (0,):            # endianness: 0 = LITTLE
                raise ValueError("big-endian checkpoint bundles are not supported")
    shards = {}
    out = {}
    for name, value in raw.items():
        e = _parse_entry(value)
        if e["sliced"]:
            raise ValueError("partitioned variable %s: sliced bundle entries are not supported" % name)
        if e["dtype"] not in _DTYPES:
            continue                                        # e.g. DT_STRING bookkeeping entries
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        blob = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if verify_data and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(blob):
            raise ValueError("tensor %s: data checksum mismatch" % name)
        arr = np.frombuffer(bytes(blob), dtype=_DTYPES[e["dtype"]]).reshape(e["shape"])
        out[name + add_suffix] = arr
    return out


# ---------------------------------------------------------------- writer (tests, conversion back)
def _entry_proto(dtype_id, shape, offset, size, crc):
    dims = b"".join(b"\x12" + write_varint(len(d)) + d for d in (b"\x08" + write_varint(int(s)) for s in shape))
    out = b"\x08" + write_varint(dtype_id)
    out += b"\x12" + write_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + write_varint(offset)
    out += b"\x28" + write_varint(size)
    out += b"\x35" + struct.pack("<I", crc)
    return out


def _build_block(items, restart_interval=16):
    buf, restarts, last = bytearray(), [], b""
    for i, (key, value) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(buf))
            shared = 0
        else:
            shared = 0
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        buf += write_varint(shared) + write_varint(len(key) - shared) + write_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def write_bundle(prefix, tensors, entries_per_block=8):
    """Write ``tensors`` (name -> array; a trailing ':0' is dropped) as a single-shard bundle."""
    names = sorted(n[:-2] if n.endswith(":0") else n for n in tensors)
    src = {(n[:-2] if n.endswith(":0") else n): np.asarray(v) for n, v in tensors.items()}     # 0-d stays 0-d (tobytes copies)
    items = [(b"", b"\x08\x01\x1a\x02\x08\x01")]            # BundleHeaderProto: num_shards = 1, version {producer: 1}
    offset = 0
    with open("%s.data-00000-of-00001" % prefix, "wb") as f:
        for n in names:
            a = src[n]
            if a.dtype not in _DTYPE_IDS:
                raise ValueError("unsupported dtype %s for %s" % (a.dtype, n))
            blob = a.tobytes()
            f.write(blob)
            items.append((n.encode(), _entry_proto(_DTYPE_IDS[a.dtype], a.shape, offset, len(blob), mask_crc(crc32c(blob)))))
            offset += len(blob)
    out = bytearray()
    index_items = []

    def emit(block):
        off = len(out)
        out.extend(block)
        trailer = b"\x00"
        out.extend(trailer + struct.pack("<I", mask_crc(crc32c(block + trailer))))
        return write_varint(off) + write_varint(len(block))

    for i in range(0, len(items), entries_per_block):
        chunk = items[i:i + entries_per_block]
        index_items.append((chunk[-1][0], emit(_build_block(chunk))))
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block(index_items, restart_interval=1))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


def is_bundle_dir(model_dir):
    return os.path.isfile(os.path.join(model_dir, "model.index"))
