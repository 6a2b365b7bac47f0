"""Host-side model inspection and parser robustness (no GPU).  The hand-written protobuf reader takes
untrusted files: truncated or corrupted input must come back as OCRS_B200_ERR_MODEL_LOAD (or parse
to something), never crash the process."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ocrs_b200 as ob
from tools.models import ensure_models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inspect_fixture_models():
    det, rec = ensure_models()
    d = ob.inspect_model(det)
    assert d["format"] == "onnx"
    assert d["inputs"] == [{"name": "image", "dims": [-1, 1, 800, 600]}]        # lib.rs:345-347
    assert d["outputs"][0]["dims"] == [-1, 1, 800, 600]
    assert d["unsupported_ops"] == [] and d["ops"]["Conv"] > 10 and d["ops"]["Sigmoid"] == 1
    r = ob.inspect_model(open(rec, "rb").read())
    assert r["inputs"] == [{"name": "line_images", "dims": [-1, 1, 64, -1]}]    # lib.rs:366-368
    assert r["outputs"][0]["dims"][-1] == 97
    assert r["ops"]["GRU"] == 2 and r["unsupported_ops"] == []
    assert r["initializer_bytes"] > 9_000_000


def test_unsupported_operator_is_reported(tmp_path):
    from oracle.onnx_io import Graph, ValueInfo, save_model
    from tools.models import _Builder
    b = _Builder()
    y = b.node("Softplus", ["x"])
    g = Graph(b.nodes, b.inits, [ValueInfo("x", 1, ["n", 4])], [ValueInfo(y, 1, ["n", 4])], name="t")
    p = str(tmp_path / "m.onnx")
    save_model(g, p)
    assert ob.inspect_model(p)["unsupported_ops"] == ["Softplus"]


def test_rten_container_and_garbage_are_rejected():
    with pytest.raises(ob.OcrsError, match="rten"):
        ob.inspect_model(b"RTEN" + bytes(60))
    with pytest.raises(ob.OcrsError):
        ob.inspect_model(b"\xff" * 64)


FUZZ = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import ocrs_b200 as ob
data = open(%r, "rb").read()
rng = np.random.default_rng(0)
ok = bad = 0
def attempt(buf):
    global ok, bad
    try:
        ob.inspect_model(bytes(buf)); ok += 1
    except ob.OcrsError:
        bad += 1
for cut in list(range(0, 600, 7)) + [int(c) for c in rng.integers(0, len(data), 60)]:
    attempt(data[:cut])                                   # truncations
head = bytearray(data[:4096])
for _ in range(300):                                      # byte flips in the structural part of the file
    buf = bytearray(data)
    for pos in rng.integers(0, 4096, int(rng.integers(1, 6))):
        buf[pos] = int(rng.integers(0, 256))
    attempt(buf)
for _ in range(100):                                      # random bytes
    attempt(rng.integers(0, 256, int(rng.integers(1, 2000)), dtype=np.uint8).tobytes())
print("done", ok, bad)
"""


def test_parser_survives_truncation_and_corruption():
    _, rec = ensure_models()
    r = subprocess.run([sys.executable, "-c", FUZZ % (ROOT, rec)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-1500:])   # a crash would be a negative return code
    tag, ok, bad = r.stdout.split()[-3:]
    assert tag == "done" and int(bad) > 100


# ---- crafted hostile graphs (dims / arity / attribute lengths are validated before
# Model::load or Model::run can index them; the validation lives in the parser, so the host-only
# inspect entry point exercises exactly what model_load runs first) ------------------------------
def _conv_graph(attrs=None, n_inputs=3, w_shape=(8, 1, 3, 3), b_shape=(8,)):
    from oracle.onnx_io import Graph, Node, ValueInfo, encode_model
    inits = {"w": np.zeros(w_shape, np.float32), "b": np.zeros(b_shape, np.float32)}
    a = {"dilations": [1, 1], "group": 1, "kernel_shape": [3, 3], "pads": [1, 1, 1, 1], "strides": [1, 1]}
    a.update(attrs or {})
    a = {k: v for k, v in a.items() if v is not None}
    node = Node("Conv", ["x", "w", "b"][:n_inputs], ["y"], a, name="c")
    return encode_model(Graph([node], inits, [ValueInfo("x", 1, ["n", 1, 8, 8])], [ValueInfo("y", 1, ["n", 8, 8, 8])]))


def test_wellformed_conv_graph_passes():
    assert ob.inspect_model(_conv_graph())["ops"] == {"Conv": 1}


@pytest.mark.parametrize("kw,msg", [
    (dict(attrs={"strides": [1]}), "strides"),
    (dict(attrs={"pads": [1, 1]}), "pads"),
    (dict(attrs={"dilations": [1, 1, 1]}), "dilations"),
    (dict(attrs={"strides": [0, 1]}), "stride"),
    (dict(attrs={"group": 0}), "group"),
    (dict(n_inputs=1), "inputs"),
    (dict(w_shape=(8, 3, 3)), "4-D"),
    (dict(b_shape=(7,)), "bias"),
])
def test_malformed_conv_attributes_are_rejected(kw, msg):
    with pytest.raises(ob.OcrsError, match=msg):
        ob.inspect_model(_conv_graph(**kw))


def _raw_tensor(dims, dtype=1, raw=b""):
    from oracle.onnx_io import _f_bytes, _f_str, _f_varint
    t = b"".join(_f_varint(1, d & 0xFFFFFFFFFFFFFFFF) for d in dims) + _f_varint(2, dtype) + _f_str(8, "t") + _f_bytes(9, raw)
    from oracle.onnx_io import _enc_value_info, ValueInfo
    gb = _f_bytes(5, t) + _f_bytes(11, _enc_value_info(ValueInfo("x", 1, [1]))) + _f_bytes(12, _enc_value_info(ValueInfo("x", 1, [1])))
    return _f_varint(1, 8) + _f_bytes(7, gb)


@pytest.mark.parametrize("dims", [[1 << 62], [-1, -1], [-4], [1 << 31, 1 << 31], [3, -1, 2]])
def test_wrapping_or_negative_tensor_dims_are_rejected(dims):
    # numel() of these dims is 0 or wraps to 0 in int64: an empty raw buffer used to pass the size check
    with pytest.raises(ob.OcrsError, match="dimension|too large"):
        ob.inspect_model(_raw_tensor(dims))


def test_gru_with_inconsistent_weights_is_rejected():
    from oracle.onnx_io import Graph, Node, ValueInfo, encode_model
    inits = {"W": np.zeros((2, 12, 5), np.float32), "R": np.zeros((2, 12, 3), np.float32)}  # R must be [2,12,4]
    node = Node("GRU", ["x", "W", "R"], ["y", "yh"], {"hidden_size": 4, "direction": "bidirectional", "linear_before_reset": 1})
    g = Graph([node], inits, [ValueInfo("x", 1, ["t", "n", 5])], [ValueInfo("y", 1, ["t", 2, "n", 4])])
    with pytest.raises(ob.OcrsError, match="GRU"):
        ob.inspect_model(encode_model(g))
