"""TensorFlow-free reader / writer of the reference's ``weights.tf`` checkpoints.

The reference restores its models with ``PaNWaveNet.load_weights(<model_dir>/weights.tf)`` (mel_inverter.py:203-210):
a TensorFlow "tensor bundle" (``weights.tf.index`` + ``weights.tf.data-0000N-of-0000M``) whose entries are named by the
Keras object graph stored under the key ``_CHECKPOINTABLE_OBJECT_GRAPH``.  TensorFlow is not part of this package, so the
three formats involved are restated here from their published layouts:

* the ``.index`` file is a LevelDB-style sorted string table: prefix-compressed key/value blocks with restart arrays,
  a 5-byte trailer per block (compression type + masked CRC-32C), an index block of block handles and a 48-byte
  footer ending in the magic number 0xdb4775248b80fb57;
* each value is a ``BundleEntryProto`` (dtype, shape, shard, offset, size, masked CRC-32C); the empty key holds the
  ``BundleHeaderProto``; string tensors are stored as varint lengths + a length checksum + the bytes;
* the object graph is a ``TrackableObjectGraph`` proto: nodes with named children and, for variables, the bundle key.

``import_weights`` walks the object graph along the reference's attribute names
(``block`` = MBExWN, wavegen_1d.py:433; ``pp_subnet_layers`` / ``ps_subnet_layers`` custom_pulsed_generator.py:278,365;
``pp_waveNetBlocks[i].wavenet.{start,end,cond_layer,conv_layers[i],res_skip_layers[i]}`` custom_AE_layers.py:177-259,499;
``wn_post_net[0]`` :490; weight-norm convs hold ``v``, ``g`` and ``conv1d_layer.bias`` conv_layers.py:59-101; PReLU
``alpha``) and returns the un-folded container of ``weights.py``.  ``export_weights`` writes the same structure, so that a
maintainer with TensorFlow can load this package's synthetic weights into the unmodified reference and pin parity
(tests/golden/make_tf_reference_goldens.py).  No released checkpoint is available offline: the reader is tested against
files produced by the writer and against hand-assembled tables (tests/test_tf_checkpoint.py) — "format unpinned" in the
same sense as the forward oracle.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
VAR_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow DataType enum values (types.proto) <-> numpy
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_STRING = 7
_NP_TO_DT = {np.dtype(v): k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------------------ CRC-32C
def _make_table() -> np.ndarray:
    t = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t[i] = c
    return t


_CRC_TABLE = _make_table()
_CRC_LIST = [int(x) for x in _CRC_TABLE]


def _raw_update(state: int, data: bytes) -> int:
    tab = _CRC_LIST
    for b in data:
        state = tab[(state ^ b) & 0xFF] ^ (state >> 8)
    return state


def _apply(cols: np.ndarray, states: np.ndarray) -> np.ndarray:
    """GF(2) matrix (32 column images) times a vector of 32-bit states."""
    out = np.zeros_like(states)
    for bit in range(32):
        out ^= np.where((states >> np.uint32(bit)) & np.uint32(1), cols[bit], np.uint32(0)).astype(np.uint32)
    return out


def _zero_operator(n_bytes: int) -> np.ndarray:
    """Column images of 'advance the raw CRC state through n zero bytes'."""
    one = np.array([_raw_update(1 << b, b"\0") for b in range(32)], dtype=np.uint32)
    result = np.array([1 << b for b in range(32)], dtype=np.uint32)      # identity
    power = one
    while n_bytes:
        if n_bytes & 1:
            result = _apply(power, result)
        power = _apply(power, power)
        n_bytes >>= 1
    return result


def crc32c(data, crc: int = 0) -> int:
    """CRC-32C (Castagnoli) of a bytes-like object, continuing from ``crc`` (tensorflow crc32c::Extend).

    Long buffers are split into 2^k lanes advanced together with NumPy and merged with the zero-advance operator
    (the CRC is affine in its state), so checkpoint-sized tensors take milliseconds without a C extension."""
    buf = np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data, dtype=np.uint8)
    state = (crc ^ 0xFFFFFFFF) & 0xFFFFFFFF
    n = buf.size
    if n < (1 << 14):
        return _raw_update(state, buf.tobytes()) ^ 0xFFFFFFFF
    lanes = 1 << min(12, max(4, int(np.log2(n)) - 9))
    length = n // lanes
    head = n - lanes * length
    state = _raw_update(state, buf[:head].tobytes())
    body = buf[head:].reshape(lanes, length)
    s = np.zeros(lanes, dtype=np.uint32)
    s[0] = state
    cols = np.ascontiguousarray(body.T)
    for j in range(length):
        s = _CRC_TABLE[(s ^ cols[j]) & np.uint32(0xFF)] ^ (s >> np.uint32(8))
    op = _zero_operator(length)
    while s.size > 1:
        s = _apply(op, s[0::2]) ^ s[1::2]
        op = _apply(op, op)
    return int(s[0]) ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    """tensorflow crc32c::Mask: rotate right by 15 and add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(masked: int) -> int:
    rot = (masked - 0xa282ead8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf wire
def _put_varint(n: int) -> bytes:
    if n < 0:
        n += 1 << 64
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _get_varint(buf, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _pb_fields(buf) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) triples of one protobuf message (wire types 0, 1, 2, 5)."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _get_varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            val = bytes(buf[pos:pos + ln])
            if len(val) != ln:
                raise ValueError("truncated protobuf field")
            pos += ln
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, val


def _pb_varint(field: int, value: int) -> bytes:
    return _put_varint(field << 3) + _put_varint(value)


def _pb_bytes(field: int, value: bytes) -> bytes:
    return _put_varint((field << 3) | 2) + _put_varint(len(value)) + value


def _pb_fixed32(field: int, value: int) -> bytes:
    return _put_varint((field << 3) | 5) + struct.pack("<I", value)


# ------------------------------------------------------------------------------------------------ sorted string table
def _parse_block(block: bytes) -> List[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise ValueError("table block too short")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise ValueError("bad restart array in table block")
    out, pos, key = [], 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise ValueError("bad key prefix length in table block")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + 5 > len(data):
        raise ValueError("table block handle points outside the file")
    block, ctype = data[offset:offset + size], data[offset + size]
    if verify:
        stored = unmask_crc(struct.unpack_from("<I", data, offset + size + 1)[0])
        if stored != crc32c(data[offset:offset + size + 1]):
            raise ValueError(f"table block at offset {offset}: checksum mismatch")
    if ctype != 0:
        raise NotImplementedError("compressed checkpoint index blocks (snappy) are not supported; TensorFlow writes "
                                  "bundle indices uncompressed")
    return block


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a sorted string table, in key order."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path} is not a TensorFlow checkpoint index (bad table magic)")
    footer = data[-48:]
    _, p = _get_varint(footer, 0)                   # metaindex handle (unused)
    _, p = _get_varint(footer, p)
    idx_off, p = _get_varint(footer, p)
    idx_size, p = _get_varint(footer, p)
    out = []
    for _, handle in _parse_block(_read_block(data, idx_off, idx_size, verify)):
        off, q = _get_varint(handle, 0)
        size, _ = _get_varint(handle, q)
        out.extend(_parse_block(_read_block(data, off, size, verify)))
    return out


def _build_block(entries: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (key, value) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(prev), len(key))
            while shared < m and prev[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        out += key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path: str, items: List[Tuple[bytes, bytes]], block_bytes: int = 4096) -> None:
    """Write sorted (key, value) pairs as an uncompressed sorted string table."""
    items = sorted(items)
    for a, b in zip(items, items[1:]):
        if a[0] == b[0]:
            raise ValueError(f"duplicate table key {a[0]!r}")
    out = bytearray()

    def emit(block: bytes) -> bytes:
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block + b"\0" + struct.pack("<I", mask_crc(crc32c(block + b"\0"))))
        return handle

    index, cur, cur_size = [], [], 0
    for key, value in items:
        cur.append((key, value))
        cur_size += len(key) + len(value) + 3
        if cur_size >= block_bytes:
            index.append((cur[-1][0], emit(_build_block(cur))))
            cur, cur_size = [], 0
    if cur or not index:
        index.append((cur[-1][0] if cur else b"", emit(_build_block(cur))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = meta + idx
    out.extend(footer + b"\0" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as f:
        f.write(bytes(out))


# ------------------------------------------------------------------------------------------------ tensor bundle
class BundleEntry:
    __slots__ = ("dtype", "shape", "shard", "offset", "size", "crc")

    def __init__(self, dtype=0, shape=(), shard=0, offset=0, size=0, crc=0):
        self.dtype, self.shape, self.shard, self.offset, self.size, self.crc = dtype, tuple(shape), shard, offset, size, crc


def _parse_entry(buf: bytes) -> BundleEntry:
    e = BundleEntry()
    for field, _, val in _pb_fields(buf):
        if field == 1:
            e.dtype = val
        elif field == 2:
            dims = []
            for f2, _, v2 in _pb_fields(val):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3
                    dims.append(size)
            e.shape = tuple(dims)
        elif field == 3:
            e.shard = val
        elif field == 4:
            e.offset = val
        elif field == 5:
            e.size = val
        elif field == 6:
            e.crc = val
        elif field == 7:
            raise NotImplementedError("sliced (partitioned) checkpoint variables are not supported")
    return e


def _encode_entry(e: BundleEntry) -> bytes:
    shape = b"".join(_pb_bytes(2, _pb_varint(1, d)) for d in e.shape)
    out = _pb_varint(1, e.dtype) + _pb_bytes(2, shape)
    if e.shard:
        out += _pb_varint(3, e.shard)
    if e.offset:
        out += _pb_varint(4, e.offset)
    out += _pb_varint(5, e.size) + _pb_fixed32(6, e.crc)
    return out


class BundleReader:
    """Random access to the tensors of one checkpoint prefix (e.g. ``<model_dir>/weights.tf``)."""

    def __init__(self, prefix: str, verify: bool = True):
        index = prefix + ".index"
        if not os.path.exists(index):
            raise FileNotFoundError(f"checkpoint index {index} not found")
        self.prefix, self.verify = prefix, verify
        self.num_shards = 1
        self.entries: Dict[str, BundleEntry] = {}
        for key, value in read_table(index, verify):
            if key == b"":
                for field, _, val in _pb_fields(value):
                    if field == 1:
                        self.num_shards = val
                    elif field == 2 and val != 0:
                        raise NotImplementedError("big-endian checkpoints are not supported")
            else:
                self.entries[key.decode("utf-8")] = _parse_entry(value)
        self._shards: Dict[int, np.memmap] = {}

    def keys(self) -> List[str]:
        return list(self.entries)

    def _bytes(self, e: BundleEntry) -> bytes:
        if e.shard not in self._shards:
            path = f"{self.prefix}.data-{e.shard:05d}-of-{self.num_shards:05d}"
            if not os.path.exists(path):
                raise FileNotFoundError(f"checkpoint shard {path} not found")
            self._shards[e.shard] = np.memmap(path, dtype=np.uint8, mode="r")
        shard = self._shards[e.shard]
        if e.offset + e.size > shard.size:
            raise ValueError("checkpoint entry points outside its data shard")
        raw = shard[e.offset:e.offset + e.size].tobytes()
        if self.verify and e.dtype != _DT_STRING and unmask_crc(e.crc) != crc32c(raw):
            raise ValueError("checkpoint tensor checksum mismatch")
        return raw

    def get(self, key: str):
        """numpy array, or ``bytes`` / list of ``bytes`` for string tensors."""
        if key not in self.entries:
            raise KeyError(f"{key} not found in checkpoint {self.prefix}")
        e = self.entries[key]
        raw = self._bytes(e)
        if e.dtype == _DT_STRING:
            count = int(np.prod(e.shape)) if e.shape else 1
            pos, lengths = 0, []
            for _ in range(count):
                ln, pos = _get_varint(raw, pos)
                lengths.append(ln)
            crc = 0
            for ln in lengths:                       # the checksum runs over the fixed-width lengths, not the varints
                crc = crc32c(_length_bytes(ln), crc)
            if self.verify and struct.unpack_from("<I", raw, pos)[0] != mask_crc(crc):
                raise ValueError("checkpoint string tensor: length checksum mismatch")
            crc = crc32c(raw[pos:pos + 4], crc)
            pos += 4
            strings = []
            for ln in lengths:
                strings.append(raw[pos:pos + ln])
                crc = crc32c(strings[-1], crc)
                pos += ln
            if self.verify and unmask_crc(e.crc) != crc:
                raise ValueError("checkpoint tensor checksum mismatch")
            return strings[0] if not e.shape else strings
        if e.dtype not in _DTYPES:
            raise NotImplementedError(f"checkpoint dtype enum {e.dtype} is not supported")
        arr = np.frombuffer(raw, dtype=_DTYPES[e.dtype])
        if arr.size != int(np.prod(e.shape, dtype=np.int64)):
            raise ValueError(f"{key}: {arr.size} elements stored, shape {e.shape}")
        return arr.reshape(e.shape).copy()


def _length_bytes(ln: int) -> bytes:
    """How a string length enters the length checksum of a DT_STRING tensor: TensorFlow's tensor_bundle
    (WriteStringTensor / ReadStringTensor) extends the CRC with the length as a 4-byte uint32 whenever it fits
    (<= UINT32_MAX) and as a uint64 only above that.  Every real weights.tf stores _CHECKPOINTABLE_OBJECT_GRAPH this way."""
    return struct.pack("<I", ln) if ln <= 0xFFFFFFFF else struct.pack("<Q", ln)


def _string_tensor_bytes(value: bytes) -> Tuple[bytes, int]:
    """Scalar DT_STRING payload and its CRC (varint length, the checksum over the fixed-width length, then the bytes)."""
    crc = crc32c(_length_bytes(len(value)))
    cks = struct.pack("<I", mask_crc(crc))
    crc = crc32c(cks, crc)
    crc = crc32c(value, crc)
    return _put_varint(len(value)) + cks + value, crc


def write_bundle(prefix: str, tensors: Dict[str, object]) -> None:
    """Write ``{key: ndarray | bytes}`` as a single-shard tensor bundle (index + data-00000-of-00001)."""
    items, data = [], bytearray()
    for key in sorted(tensors, key=lambda k: k.encode("utf-8")):
        value = tensors[key]
        if isinstance(value, (bytes, bytearray)):
            raw, crc = _string_tensor_bytes(bytes(value))
            e = BundleEntry(_DT_STRING, (), 0, len(data), len(raw), mask_crc(crc))
        else:
            arr = np.asarray(value)
            if arr.dtype not in _NP_TO_DT:
                raise NotImplementedError(f"dtype {arr.dtype} cannot be stored")
            raw = arr.tobytes()
            e = BundleEntry(_NP_TO_DT[arr.dtype], arr.shape, 0, len(data), len(raw), mask_crc(crc32c(raw)))
        data += raw
        items.append((key.encode("utf-8"), _encode_entry(e)))
    header = _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1))      # num_shards = 1, little endian, version.producer = 1
    items.append((b"", header))
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    write_table(prefix + ".index", items)


# ------------------------------------------------------------------------------------------------ object graph
class GraphNode:
    __slots__ = ("children", "attributes")

    def __init__(self):
        self.children: Dict[str, int] = {}
        self.attributes: Dict[str, Tuple[str, str]] = {}      # name -> (full_name, checkpoint_key)


def parse_object_graph(buf: bytes) -> List[GraphNode]:
    nodes = []
    for field, _, val in _pb_fields(buf):
        if field != 1:
            continue
        node = GraphNode()
        for f2, _, v2 in _pb_fields(val):
            if f2 == 1:
                node_id, name = 0, ""
                for f3, _, v3 in _pb_fields(v2):
                    if f3 == 1:
                        node_id = v3
                    elif f3 == 2:
                        name = v3.decode("utf-8")
                node.children[name] = node_id
            elif f2 == 2:
                name = full = key = ""
                for f3, _, v3 in _pb_fields(v2):
                    if f3 == 1:
                        name = v3.decode("utf-8")
                    elif f3 == 2:
                        full = v3.decode("utf-8")
                    elif f3 == 3:
                        key = v3.decode("utf-8")
                node.attributes[name] = (full, key)
        nodes.append(node)
    return nodes


def encode_object_graph(nodes: List[GraphNode]) -> bytes:
    out = b""
    for node in nodes:
        body = b""
        for name, node_id in node.children.items():
            body += _pb_bytes(1, _pb_varint(1, node_id) + _pb_bytes(2, name.encode("utf-8")))
        for name, (full, key) in node.attributes.items():
            body += _pb_bytes(2, _pb_bytes(1, name.encode("utf-8")) + _pb_bytes(2, full.encode("utf-8")) +
                              _pb_bytes(3, key.encode("utf-8")))
        out += _pb_bytes(1, body)
    return out


class _Graph:
    def __init__(self, reader: BundleReader):
        if OBJECT_GRAPH_KEY not in reader.entries:
            raise ValueError(f"{reader.prefix} holds no object graph; only object-based (TF2 / Keras save_weights) "
                             f"checkpoints are supported")
        self.reader = reader
        self.nodes = parse_object_graph(reader.get(OBJECT_GRAPH_KEY))
        if not self.nodes:
            raise ValueError("empty object graph")

    def child(self, node: int, *path: str) -> int:
        for name in path:
            kids = self.nodes[node].children
            if name not in kids:
                raise KeyError(f"object graph: no child '{name}' (have {sorted(kids)})")
            node = kids[name]
        return node

    def has(self, node: int, *path: str) -> bool:
        try:
            self.child(node, *path)
            return True
        except KeyError:
            return False

    def list_items(self, node: int) -> List[int]:
        kids = self.nodes[node].children
        return [kids[k] for k in sorted((k for k in kids if k.isdigit()), key=int)]

    def variable(self, node: int, *path: str) -> np.ndarray:
        n = self.child(node, *path)
        attrs = self.nodes[n].attributes
        if "VARIABLE_VALUE" not in attrs:
            raise KeyError(f"object graph node {'/'.join(path)} is not a variable")
        return self.reader.get(attrs["VARIABLE_VALUE"][1])

    def find(self, child_name: str) -> int:
        """First node (breadth first from the root) that has a child called ``child_name``."""
        seen, queue = {0}, [0]
        while queue:
            n = queue.pop(0)
            if child_name in self.nodes[n].children:
                return n
            for c in self.nodes[n].children.values():
                if c not in seen:
                    seen.add(c)
                    queue.append(c)
        raise KeyError(f"object graph: no object with a child '{child_name}'")


def _conv_weights(g: _Graph, node: int, name: str, out: Dict[str, np.ndarray]) -> None:
    """v / g / bias of one TF2C_Conv1DWeightNorm (conv_layers.py:59-101): ``v`` is the former Conv1D kernel."""
    if g.has(node, "v"):
        v = g.variable(node, "v")
    elif g.has(node, "conv1d_layer", "kernel"):
        v = g.variable(node, "conv1d_layer", "kernel")
    else:
        raise KeyError(f"{name}: neither v nor conv1d_layer/kernel in the checkpoint")
    out[f"{name}/v"] = np.asarray(v, dtype=np.float32)
    if g.has(node, "g"):
        out[f"{name}/g"] = np.asarray(g.variable(node, "g"), dtype=np.float32).reshape(-1)
    else:                                           # use_weight_norm=False: the kernel is used as stored
        out[f"{name}/g"] = np.linalg.norm(out[f"{name}/v"].reshape(-1, v.shape[-1]).astype(np.float64), axis=0).astype(np.float32)
    if g.has(node, "conv1d_layer", "bias"):
        out[f"{name}/bias"] = np.asarray(g.variable(node, "conv1d_layer", "bias"), dtype=np.float32).reshape(-1)
    else:
        out[f"{name}/bias"] = np.zeros(v.shape[-1], dtype=np.float32)


def _subnet_weights(g: _Graph, list_node: int, ops, out: Dict[str, np.ndarray]) -> None:
    """k-th weight-carrying conv / PReLU of the layer list <-> k-th conv / activation of the plan's op list."""
    from .plan import ACT_PRELU
    items = g.list_items(list_node)
    convs = [n for n in items if g.has(n, "v") or g.has(n, "conv1d_layer")]
    prelus = [n for n in items if g.has(n, "alpha")]
    plan_convs = [op.conv for op in ops if op.kind == "conv"]
    plan_acts = [op for op in ops if op.act == ACT_PRELU and op.act_name]
    if len(convs) != len(plan_convs):
        raise ValueError(f"checkpoint sub-net has {len(convs)} conv layers, the config describes {len(plan_convs)}")
    if len(prelus) != len(plan_acts):
        raise ValueError(f"checkpoint sub-net has {len(prelus)} PReLU layers, the config describes {len(plan_acts)}")
    for node, layer in zip(convs, plan_convs):
        _conv_weights(g, node, layer.name, out)
    for node, op in zip(prelus, plan_acts):
        out[f"{op.act_name}/alpha"] = np.asarray(g.variable(node, "alpha"), dtype=np.float32).reshape(-1)


def import_weights(prefix: str, plan, verify: bool = True) -> Dict[str, np.ndarray]:
    """Read a reference checkpoint (``<dir>/weights.tf``) into the container of ``weights.py`` and check it against
    the plan (shapes, completeness)."""
    from . import weights as W
    g = _Graph(BundleReader(prefix, verify))
    block = g.child(0, "block") if g.has(0, "block") else g.find("pp_waveNetBlocks")
    out: Dict[str, np.ndarray] = {}
    _subnet_weights(g, g.child(block, "pp_subnet_layers"), plan.pp_ops, out)
    if plan.ps_ops:                                        # ps_off models hold no PS sub-net
        _subnet_weights(g, g.child(block, "ps_subnet_layers"), plan.ps_ops, out)
    blocks = g.list_items(g.child(block, "pp_waveNetBlocks"))
    if len(blocks) != len(plan.blocks):
        raise ValueError(f"checkpoint holds {len(blocks)} WaveNet blocks, the config describes {len(plan.blocks)}")
    for node, spec in zip(blocks, plan.blocks):
        wn = g.child(node, "wavenet")
        base = spec.name + "_WNBlock_WN"
        _conv_weights(g, g.child(wn, "start"), f"{base}/start", out)
        _conv_weights(g, g.child(wn, "end"), f"{base}/end", out)
        _conv_weights(g, g.child(wn, "cond_layer"), f"{base}/cond_", out)
        conv_nodes = g.list_items(g.child(wn, "conv_layers"))
        rs_nodes = g.list_items(g.child(wn, "res_skip_layers"))
        if len(conv_nodes) != spec.n_layers or len(rs_nodes) != spec.n_layers:
            raise ValueError(f"checkpoint WaveNet has {len(conv_nodes)} layers, the config describes {spec.n_layers}")
        for i, (cn, rn) in enumerate(zip(conv_nodes, rs_nodes)):
            _conv_weights(g, cn, f"{base}/conv1D_{i}", out)
            _conv_weights(g, rn, f"{base}/res_skip_{i}", out)
        if spec.up > 1:                                    # WaveNetAEBlock.up_down_sample (custom_AE_layers.py:519-526)
            _conv_weights(g, g.child(node, "up_down_sample"), spec.up_name, out)
    post = g.list_items(g.child(block, "wn_post_net"))
    if len(post) != 1:
        raise ValueError("checkpoint wn_post_net does not hold exactly one layer")
    _conv_weights(g, post[0], plan.post_name, out)
    W.check(plan, out)
    return out


# ------------------------------------------------------------------------------------------------ export
def reference_subnet_layout(specs, base_name: str, final_nks: Optional[int], has_final_act: bool,
                            target_ups: Optional[int], pad_to_valid: bool = False,
                            remove_inactive_pad_layers: bool = False) -> List[Tuple[str, str]]:
    """(kind, keras name) of every entry of the reference's sub-net layer list, in list order
    (generate_subnet_from_specs, custom_pulsed_generator.py:55-146): kinds pad | conv | lin | act | final_act."""
    layers: List[Tuple[str, str]] = []
    total_ups, ii = 1, 0
    if not specs:
        return layers
    for ii, spec in enumerate(specs):
        if spec[0] == "L":
            layers.append(("lin", f"{base_name}_LinUpLayer_{ii}"))
            continue
        ks, linear_up, up = int(spec[0]), False, 1
        if len(spec) > 2:
            if isinstance(spec[2], str):
                linear_up = spec[2][0] == "L"
                up = int(spec[2][1:])
            else:
                up = int(spec[2])
        has_pad = ((ks - 1) // 2 + ((ks - 1) % 2)) > 0
        if linear_up:
            if (not remove_inactive_pad_layers) or has_pad:
                layers.append(("pad", f"{base_name}_Pad_{ii}"))
            layers.append(("conv", f"{base_name}_Layer_{ii}"))
            layers.append(("lin", f"{base_name}_LinUpLayer_{ii}"))
        elif up > 1:
            if pad_to_valid and has_pad:
                layers.append(("pad", f"{base_name}_Pad_{ii}"))
            layers.append(("conv", f"{base_name}_Layer_{ii}"))
        else:
            if (not remove_inactive_pad_layers) or has_pad:
                layers.append(("pad", f"{base_name}_Pad_{ii}"))
            layers.append(("conv", f"{base_name}_Layer_{ii}"))
        layers.append(("act", f"{base_name}_ActLayer_{ii}"))
        total_ups *= up
    if final_nks is not None:
        if pad_to_valid and ((final_nks - 1) // 2 + ((final_nks - 1) % 2)) > 0:
            layers.append(("pad", f"{base_name}_Pad_{ii}"))
        layers.append(("conv", f"{base_name}_Layer_final"))
        if target_ups is not None and total_ups != target_ups:
            layers.append(("lin", f"{base_name}_linear_interp"))
        if has_final_act:
            layers.append(("final_act", f"{base_name}_Layer_finalAct"))
    return layers


class _GraphBuilder:
    def __init__(self):
        self.nodes: List[GraphNode] = [GraphNode()]
        self.tensors: Dict[str, object] = {}

    def add(self, parent: int, name: str) -> int:
        self.nodes.append(GraphNode())
        self.nodes[parent].children[name] = len(self.nodes) - 1
        return len(self.nodes) - 1

    def variable(self, parent: int, name: str, path: str, full_name: str, value: np.ndarray) -> None:
        node = self.add(parent, name)
        key = path + VAR_SUFFIX
        self.nodes[node].attributes["VARIABLE_VALUE"] = (full_name, key)
        self.tensors[key] = np.ascontiguousarray(value, dtype=np.float32)

    def conv(self, parent: int, name: str, path: str, keras_name: str, weights: Dict[str, np.ndarray], wname: str) -> None:
        node = self.add(parent, name)
        inner = self.add(node, "conv1d_layer")
        self.variable(inner, "bias", f"{path}/conv1d_layer/bias", f"{keras_name}/bias", weights[f"{wname}/bias"])
        self.variable(node, "v", f"{path}/v", f"{keras_name}/kernel", weights[f"{wname}/v"])
        self.variable(node, "g", f"{path}/g", f"{keras_name}_base/g", weights[f"{wname}/g"])


def export_weights(prefix: str, hparams: Dict, weights: Dict[str, np.ndarray]) -> None:
    """Write the un-folded weight container as a checkpoint with the reference's object graph, loadable by
    ``PaNWaveNet.load_weights(prefix)`` on the same config (and by ``import_weights``)."""
    from .plan import build_plan
    plan = build_plan(hparams, finalize=False)
    mc = hparams["mbexwn_config"]
    b = _GraphBuilder()
    block = b.add(0, "block")
    rip = bool(mc.get("remove_inactive_pad_layers", False))
    for attr, specs, base, final_act, target, valid in (
            ("pp_subnet_layers", mc["pp_subnet"], "PulsPar", True, plan.pulse_per_frame,
             bool(mc.get("pp_subnet_use_valid_padding", False))),
            ("ps_subnet_layers", mc["ps_subnet"] if plan.ps_ops else [], "PS", False, None,
             bool(mc.get("ps_subnet_use_valid_padding", False)))):
        lst = b.add(block, attr)
        for i, (kind, name) in enumerate(reference_subnet_layout(specs, base, 1, final_act, target, valid, rip)):
            path = f"block/{attr}/{i}"
            if kind == "conv":
                b.conv(lst, str(i), path, name, weights, name)
            elif kind == "act" and plan.use_prelu:
                node = b.add(lst, str(i))
                b.variable(node, "alpha", f"{path}/alpha", f"{name}/alpha", weights[f"{name}/alpha"].reshape(1, -1))
            else:
                b.add(lst, str(i))
    blocks = b.add(block, "pp_waveNetBlocks")
    for ib, spec in enumerate(plan.blocks):
        wnb = b.add(blocks, str(ib))
        wn = b.add(wnb, "wavenet")
        base, path = spec.name + "_WNBlock_WN", f"block/pp_waveNetBlocks/{ib}/wavenet"
        b.conv(wn, "start", f"{path}/start", "start", weights, f"{base}/start")
        b.conv(wn, "end", f"{path}/end", "end", weights, f"{base}/end")
        b.conv(wn, "cond_layer", f"{path}/cond_layer", "cond_", weights, f"{base}/cond_")
        convs, rss = b.add(wn, "conv_layers"), b.add(wn, "res_skip_layers")
        for i in range(spec.n_layers):
            b.conv(convs, str(i), f"{path}/conv_layers/{i}", f"conv1D_{i}", weights, f"{base}/conv1D_{i}")
            b.conv(rss, str(i), f"{path}/res_skip_layers/{i}", f"res_skip_{i}", weights, f"{base}/res_skip_{i}")
        if spec.up > 1:
            b.conv(wnb, "up_down_sample", f"block/pp_waveNetBlocks/{ib}/up_down_sample", spec.up_name, weights, spec.up_name)
    post = b.add(block, "wn_post_net")
    b.conv(post, "0", "block/wn_post_net/0", plan.post_name, weights, plan.post_name)
    tensors = dict(b.tensors)
    tensors[OBJECT_GRAPH_KEY] = encode_object_graph(b.nodes)
    write_bundle(prefix, tensors)
