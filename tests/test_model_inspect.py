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
