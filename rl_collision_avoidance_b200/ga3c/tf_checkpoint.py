"""Reader for TensorFlow-1.x "bundle" checkpoints (`<prefix>.index` + `<prefix>.data-00000-of-00001`) without
TensorFlow, so that the reference's trained GA3C-CADRL weights can be imported into `NetworkVP_rnn`.

What it replaces: `tf.train.Saver.restore` as used by `NetworkVPCore.load` (GA3C/NetworkVPCore.py:231-242) and
`simple_load` (GCA/envs/policies/GA3C_CADRL/network.py:41-72).  Formats (public, documented in the TensorFlow and
LevelDB sources): the `.index` file is a LevelDB table (prefix-compressed key/value blocks + block-handle footer, magic
0xdb4775248b80fb57) whose values are `BundleEntryProto` messages {dtype=1, shape=2, shard_id=3, offset=4, size=5,
crc32c=6}; tensors are stored raw, little endian, in the data shard at [offset, offset+size).

Only what those checkpoints need is implemented: uncompressed blocks, float32/int32/int64 tensors, one data shard.
"""
import struct

import numpy as np

_TABLE_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _read_block(data, offset, size):
    """Returns [(key, value)] of one LevelDB block; the byte after the block is the compression type."""
    if data[offset + size] != 0:
        raise NotImplementedError("compressed checkpoint index blocks are not supported")
    block = data[offset:offset + size]
    num_restarts = struct.unpack_from("<I", block, size - 4)[0]
    end = size - 4 - 4 * num_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _parse_proto(buf):
    """Minimal protobuf wire parser: {field: [values]} (varint -> int, length-delimited -> bytes, fixed32/64 -> int)."""
    pos, out = 0, {}
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
        elif wire == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.setdefault(field, []).append(v)
    return out


def read_index(index_path):
    """{tensor name: dict(dtype, shape, shard, offset, size)} from a `.index` file."""
    data = open(index_path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % index_path)
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)      # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _read_block(data, idx_off, idx_size):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, value in _read_block(data, off, size):
            if key == b"":
                continue               # BundleHeaderProto
            msg = _parse_proto(value)
            shape = []
            for shp in msg.get(2, []):
                for dim in _parse_proto(shp).get(2, []):
                    shape.append(_parse_proto(dim).get(1, [0])[0])
            entries[key.decode()] = dict(dtype=msg.get(1, [0])[0], shape=tuple(shape), shard=msg.get(3, [0])[0],
                                         offset=msg.get(4, [0])[0], size=msg.get(5, [0])[0])
    return entries


def load_checkpoint(prefix):
    """{variable name: numpy array} for every tensor of `<prefix>.index` / `<prefix>.data-00000-of-00001`."""
    entries = read_index(prefix + ".index")
    raw = open(prefix + ".data-00000-of-00001", "rb").read()
    out = {}
    for name, e in entries.items():
        if e["dtype"] not in _DTYPES:
            continue
        if e["shard"] != 0:
            raise NotImplementedError("multi-shard checkpoints are not supported")
        dt = np.dtype(_DTYPES[e["dtype"]]).newbyteorder("<")
        arr = np.frombuffer(raw, dtype=dt, count=e["size"] // dt.itemsize, offset=e["offset"])
        out[name] = arr.reshape(e["shape"]).copy()
    return out


NETWORK_VARIABLES = ("rnn/lstm_cell/kernel", "rnn/lstm_cell/bias", "layer1/kernel", "layer1/bias", "layer2/kernel",
                     "layer2/bias", "fullyconnected1/kernel", "fullyconnected1/bias", "logits_p/kernel", "logits_p/bias",
                     "logits_v/kernel", "logits_v/bias")


def network_variables(prefix):
    """The twelve weight tensors of NetworkVP_rnn (TF variable names; Adam slots and `step` are dropped)."""
    # the reference saves with Saver({var.name: var}) (NetworkVPCore.py:57), so names carry TF's ":0" suffix
    ck = {(k[:-2] if k.endswith(":0") else k): v for k, v in load_checkpoint(prefix).items()}
    missing = [n for n in NETWORK_VARIABLES if n not in ck]
    if missing:
        raise KeyError("checkpoint %s lacks %s (has %s)" % (prefix, missing, sorted(ck)[:20]))
    return {n: ck[n].astype(np.float32) for n in NETWORK_VARIABLES}
