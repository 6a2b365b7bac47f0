"""Minimal ONNX (protobuf) reader/writer (oracle + fixture tooling; test infrastructure).

There is no `onnx` python package in this image, so the model files consumed by the engine
and by the oracle are written and read with this hand-rolled protobuf wire codec.  Field numbers
are those of onnx.proto3 (ModelProto.graph=7, GraphProto.node=1/initializer=5/input=11/
output=12, NodeProto.input=1/output=2/name=3/op_type=4/attribute=5, TensorProto.dims=1/
data_type=2/float_data=4/int64_data=7/name=8/raw_data=9, AttributeProto.name=1/f=2/i=3/s=4/
t=5/floats=7/ints=8/type=20).  The product's C++ loader (ocrs_b200/csrc/onnx_reader.cpp) is an
independent implementation of the same wire format.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

def _c(a):
    """C-contiguous copy that keeps 0-d arrays 0-d (np.ascontiguousarray would not)."""
    a = np.asarray(a)
    return a if a.flags.c_contiguous else a.copy(order="C")


# TensorProto.DataType
FLOAT, UINT8, INT8, INT32, INT64, BOOL = 1, 2, 3, 6, 7, 9
_NP_OF = {FLOAT: np.float32, UINT8: np.uint8, INT8: np.int8, INT32: np.int32, INT64: np.int64, BOOL: np.bool_}
_DT_OF = {np.dtype(v): k for k, v in _NP_OF.items()}

# AttributeProto.AttributeType
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS = 1, 2, 3, 4, 6, 7


# ---- wire primitives --------------------------------------------------------------------------
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fieldno: int, wt: int) -> bytes:
    return _varint((fieldno << 3) | wt)


def _f_varint(fieldno: int, v: int) -> bytes:
    return _key(fieldno, 0) + _varint(int(v))


def _f_bytes(fieldno: int, b: bytes) -> bytes:
    return _key(fieldno, 2) + _varint(len(b)) + b


def _f_str(fieldno: int, s: str) -> bytes:
    return _f_bytes(fieldno, s.encode("utf-8"))


def _f_float(fieldno: int, v: float) -> bytes:
    return _key(fieldno, 5) + struct.pack("<f", v)


def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_message(buf: bytes) -> Dict[int, List[Any]]:
    """field number -> list of raw values (int for varint/fixed, bytes for length-delimited)."""
    out: Dict[int, List[Any]] = {}
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        fieldno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.setdefault(fieldno, []).append(v)
    return out


def _packed_varints(vals: List[Any]) -> List[int]:
    out: List[int] = []
    for v in vals:
        if isinstance(v, (bytes, bytearray, memoryview)):
            pos = 0
            while pos < len(v):
                x, pos = _read_varint(v, pos)
                out.append(_signed64(x))
        else:
            out.append(_signed64(v))
    return out


# ---- model objects ---------------------------------------------------------------------------
@dataclass
class Node:
    op_type: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, Any] = field(default_factory=dict)
    name: str = ""


@dataclass
class ValueInfo:
    name: str
    elem_type: int
    shape: List[Union[int, str]]  # int = fixed, str = symbolic


@dataclass
class Graph:
    nodes: List[Node]
    initializers: Dict[str, np.ndarray]
    inputs: List[ValueInfo]
    outputs: List[ValueInfo]
    name: str = "graph"
    opset: int = 17


# ---- encode ----------------------------------------------------------------------------------
def _enc_tensor(name: str, arr: np.ndarray) -> bytes:
    arr = _c(arr)
    out = b"".join(_f_varint(1, d) for d in arr.shape)
    out += _f_varint(2, _DT_OF[arr.dtype])
    out += _f_str(8, name)
    out += _f_bytes(9, arr.tobytes())
    return out


def _enc_attr(name: str, v: Any) -> bytes:
    out = _f_str(1, name)
    if isinstance(v, bool):
        v = int(v)
    if isinstance(v, int):
        out += _f_varint(3, v) + _f_varint(20, A_INT)
    elif isinstance(v, float):
        out += _f_float(2, v) + _f_varint(20, A_FLOAT)
    elif isinstance(v, str):
        out += _f_bytes(4, v.encode()) + _f_varint(20, A_STRING)
    elif isinstance(v, np.ndarray):
        out += _f_bytes(5, _enc_tensor("", v)) + _f_varint(20, A_TENSOR)
    elif isinstance(v, (list, tuple)) and all(isinstance(x, int) for x in v):
        out += b"".join(_f_varint(8, x) for x in v) + _f_varint(20, A_INTS)
    elif isinstance(v, (list, tuple)):
        out += b"".join(_f_float(7, float(x)) for x in v) + _f_varint(20, A_FLOATS)
    else:
        raise TypeError(f"attr {name}: {type(v)}")
    return out


def _enc_value_info(vi: ValueInfo) -> bytes:
    dims = b""
    for d in vi.shape:
        dim = _f_varint(1, d) if isinstance(d, int) else _f_str(2, d)
        dims += _f_bytes(1, dim)
    tensor_type = _f_varint(1, vi.elem_type) + _f_bytes(2, dims)
    type_proto = _f_bytes(1, tensor_type)
    return _f_str(1, vi.name) + _f_bytes(2, type_proto)


def encode_model(g: Graph) -> bytes:
    gb = b""
    for n in g.nodes:
        nb = b"".join(_f_str(1, s) for s in n.inputs)
        nb += b"".join(_f_str(2, s) for s in n.outputs)
        if n.name:
            nb += _f_str(3, n.name)
        nb += _f_str(4, n.op_type)
        nb += b"".join(_f_bytes(5, _enc_attr(k, v)) for k, v in n.attrs.items())
        gb += _f_bytes(1, nb)
    gb += _f_str(2, g.name)
    for name, arr in g.initializers.items():
        gb += _f_bytes(5, _enc_tensor(name, arr))
    for vi in g.inputs:
        gb += _f_bytes(11, _enc_value_info(vi))
    for vi in g.outputs:
        gb += _f_bytes(12, _enc_value_info(vi))
    opset = _f_str(1, "") + _f_varint(2, g.opset)
    return _f_varint(1, 8) + _f_str(2, "ocrs-b200-fixture") + _f_bytes(7, gb) + _f_bytes(8, opset)


# ---- decode ----------------------------------------------------------------------------------
def _dec_tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    m = parse_message(buf)
    dims = _packed_varints(m.get(1, []))
    dt = m[2][0]
    name = m.get(8, [b""])[0].decode()
    np_dt = _NP_OF[dt]
    if 9 in m:
        arr = np.frombuffer(m[9][0], dtype=np_dt).copy()
    elif dt == FLOAT and 4 in m:
        raw = b"".join(v if len(v) != 4 or isinstance(v, bytes) else v for v in m[4])
        arr = np.frombuffer(raw, dtype=np.float32).copy()
    elif dt == INT64 and 7 in m:
        arr = np.array(_packed_varints(m[7]), dtype=np.int64)
    elif dt in (INT32, UINT8, INT8, BOOL) and 5 in m:
        arr = np.array(_packed_varints(m[5])).astype(np_dt)
    else:
        arr = np.zeros(0, dtype=np_dt)
    return name, arr.reshape(dims)


def _dec_attr(buf: bytes) -> Tuple[str, Any]:
    m = parse_message(buf)
    name = m[1][0].decode()
    t = m.get(20, [0])[0]
    if t == A_INT or (t == 0 and 3 in m):
        return name, _signed64(m[3][0])
    if t == A_FLOAT or (t == 0 and 2 in m):
        return name, struct.unpack("<f", m[2][0])[0]
    if t == A_STRING or (t == 0 and 4 in m):
        return name, m[4][0].decode()
    if t == A_TENSOR or (t == 0 and 5 in m):
        return name, _dec_tensor(m[5][0])[1]
    if t == A_INTS or (t == 0 and 8 in m):
        return name, _packed_varints(m.get(8, []))
    if t == A_FLOATS or (t == 0 and 7 in m):
        vals = []
        for v in m.get(7, []):
            vals.extend(struct.unpack("<%df" % (len(v) // 4), v))
        return name, list(vals)
    raise ValueError(f"attribute {name}: unsupported type {t}")


def _dec_value_info(buf: bytes) -> ValueInfo:
    m = parse_message(buf)
    name = m[1][0].decode()
    elem_type, shape = 0, []
    if 2 in m:
        tp = parse_message(m[2][0])
        if 1 in tp:
            tt = parse_message(tp[1][0])
            elem_type = tt.get(1, [0])[0]
            if 2 in tt:
                for dim in parse_message(tt[2][0]).get(1, []):
                    dm = parse_message(dim)
                    if 1 in dm:
                        shape.append(_signed64(dm[1][0]))
                    elif 2 in dm:
                        shape.append(dm[2][0].decode())
                    else:
                        shape.append("?")
    return ValueInfo(name, elem_type, shape)


def decode_model(buf: bytes) -> Graph:
    m = parse_message(buf)
    g = parse_message(m[7][0])
    nodes = []
    for nb in g.get(1, []):
        nm = parse_message(nb)
        attrs = dict(_dec_attr(a) for a in nm.get(5, []))
        nodes.append(Node(
            op_type=nm[4][0].decode(),
            inputs=[s.decode() for s in nm.get(1, [])],
            outputs=[s.decode() for s in nm.get(2, [])],
            attrs=attrs,
            name=nm.get(3, [b""])[0].decode(),
        ))
    inits = dict(_dec_tensor(t) for t in g.get(5, []))
    inputs = [_dec_value_info(v) for v in g.get(11, [])]
    inputs = [v for v in inputs if v.name not in inits]
    outputs = [_dec_value_info(v) for v in g.get(12, [])]
    opset = 17
    for ob in m.get(8, []):
        om = parse_message(ob)
        if not om.get(1, [b""])[0]:
            opset = om.get(2, [17])[0]
    return Graph(nodes, inits, inputs, outputs, g.get(2, [b"graph"])[0].decode(), opset)


def load_model(path: str) -> Graph:
    with open(path, "rb") as fp:
        return decode_model(fp.read())


def save_model(g: Graph, path: str) -> None:
    with open(path, "wb") as fp:
        fp.write(encode_model(g))
