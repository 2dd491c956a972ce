"""Reader for TensorFlow "V2" checkpoint bundles, without TensorFlow.

The reference restores its weights with ``tf.train.Saver.restore``
(/root/reference/catfish/models/rnn_class.py:191-196, called from
/root/reference/catfish/neural_network.py:32).  TensorFlow is not a dependency
of this package, so the bundle is parsed directly:

* ``<prefix>.index`` is a LevelDB-style sorted string table: 48-byte footer
  (metaindex handle, index handle, magic), an index block pointing at data
  blocks, each block a sequence of prefix-compressed ``key -> value`` entries
  followed by a restart array.  Values are ``BundleEntryProto`` messages
  (dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6).
* ``<prefix>.data-00000-of-00001`` holds the raw little-endian tensor bytes.

Only what the shipped checkpoint needs is implemented: uncompressed blocks,
one shard, DT_FLOAT / DT_INT32 / DT_INT64 tensors.  Every tensor's masked
CRC32C is verified, so a wrong parse cannot go unnoticed.
"""

import struct

import numpy as np

_TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _block_entries(block):
    """Yield (key, value) pairs of one table block (restart array stripped)."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        value_len, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + value_len])
        pos += value_len


def _read_block(data, offset, size):
    if data[offset + size] != 0:
        raise ValueError("compressed checkpoint index blocks are not supported")
    return data[offset:offset + size]


def _parse_shape(buf):
    dims = []
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        if tag == 0x12:                      # TensorShapeProto.dim
            ln, pos = _varint(buf, pos)
            sub, spos = buf[pos:pos + ln], 0
            pos += ln
            size = 0
            while spos < len(sub):
                stag, spos = _varint(sub, spos)
                if stag == 0x08:
                    size, spos = _varint(sub, spos)
                elif stag & 7 == 2:
                    sl, spos = _varint(sub, spos)
                    spos += sl
                else:
                    _, spos = _varint(sub, spos)
            dims.append(size)
        elif tag & 7 == 0:
            _, pos = _varint(buf, pos)
        elif tag & 7 == 2:
            ln, pos = _varint(buf, pos)
            pos += ln
        else:
            raise ValueError("unexpected wire type in TensorShapeProto")
    return tuple(dims)


def _parse_entry(buf):
    entry = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None}
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
            if field == 1:
                entry["dtype"] = v
            elif field == 3:
                entry["shard_id"] = v
            elif field == 4:
                entry["offset"] = v
            elif field == 5:
                entry["size"] = v
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            if field == 2:
                entry["shape"] = _parse_shape(buf[pos:pos + ln])
            pos += ln
        elif wire == 5:
            if field == 6:
                entry["crc32c"] = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wire == 1:
            pos += 8
        else:
            raise ValueError("unexpected wire type in BundleEntryProto")
    return entry


def _make_crc_table():
    poly = 0x82F63B78
    table = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        table[i] = c
    return table


_CRC_TABLE = _make_crc_table()


def crc32c(data):
    """Castagnoli CRC of ``data`` (bytes)."""
    table = _CRC_TABLE
    crc = 0xFFFFFFFF
    for b in data:
        crc = int(table[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    """TensorFlow's stored form: rotate right by 15 bits and add a constant."""
    crc = crc32c(data)
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def read_index(prefix):
    """Return ``{tensor_name: entry}`` for the bundle at ``prefix``."""
    with open(prefix + ".index", "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _TABLE_MAGIC:
        raise ValueError("%s.index is not a TensorFlow V2 checkpoint index" % prefix)
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)            # metaindex offset
    _, pos = _varint(footer, pos)            # metaindex size
    index_off, pos = _varint(footer, pos)
    index_size, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _block_entries(_read_block(data, index_off, index_size)):
        off, hpos = _varint(handle, 0)
        size, _ = _varint(handle, hpos)
        for key, value in _block_entries(_read_block(data, off, size)):
            if key == b"":
                continue                     # BundleHeaderProto
            entries[key.decode("utf-8")] = _parse_entry(value)
    return entries


def load_checkpoint(prefix, names=None, verify_crc=True):
    """Load tensors of the bundle ``prefix`` as ``{name: np.ndarray}``.

    ``names``: optional predicate or collection restricting which tensors are
    read (the shipped bundle carries RMSProp slots inference never touches).
    """
    entries = read_index(prefix)
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        blob = f.read()
    out = {}
    for name, e in entries.items():
        if names is not None:
            if callable(names):
                if not names(name):
                    continue
            elif name not in names:
                continue
        if e["shard_id"] != 0:
            raise ValueError("multi-shard bundles are not supported")
        if e["dtype"] not in _DTYPES:
            raise ValueError("tensor %s has unsupported dtype %d" % (name, e["dtype"]))
        raw = blob[e["offset"]:e["offset"] + e["size"]]
        if verify_crc and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
            raise ValueError("crc32c mismatch for tensor %s" % name)
        out[name] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def is_inference_tensor(name):
    """True for variables the forward graph reads (drops optimizer slots)."""
    return "RMSProp" not in name and "Adam" not in name and "beta1_power" not in name \
        and "beta2_power" not in name and name != "global_step"
