"""Per-layer error of the recognition conv stack: tensor-core path vs torch fp32 vs torch fp64.
Builds truncated copies of the model (graph output = an intermediate tensor)."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrs_b200 as ob
from oracle.onnx_io import ValueInfo, load_model, save_model
from tools.models import ensure_models
from tools.synth import make_line_batch

_, rec = ensure_models()
g = load_model(rec)
x = make_line_batch(300, 16)
inits = {k: v for k, v in g.initializers.items()} if isinstance(g.initializers, dict) else {t.name: t for t in g.initializers}


def arr(name):
    t = inits[name]
    return np.asarray(t.array if hasattr(t, "array") else t)


def ref(upto, dtype):
    t = torch.from_numpy(x).to(dtype)
    for n in g.nodes[:upto + 1]:
        op = n.op_type if hasattr(n, "op_type") else n.op
        if op == "Conv":
            t = F.conv2d(t, torch.from_numpy(arr(n.inputs[1])).to(dtype), torch.from_numpy(arr(n.inputs[2])).to(dtype), padding=1)
        elif op == "Relu":
            t = F.relu(t)
        elif op == "MaxPool":
            a = n.attrs if hasattr(n, "attrs") else n.attributes
            t = F.max_pool2d(t, tuple(a["kernel_shape"]), tuple(a["strides"]))
        else:
            raise SystemExit(op)
    return t.numpy()


tmp = tempfile.mkdtemp()
for cut in (2, 5, 7, 10, 12, 15):
    out_name = g.nodes[cut].outputs[0]
    g2 = load_model(rec)
    g2.nodes = g2.nodes[:cut + 1]
    g2.outputs = [ValueInfo(out_name, 1, ["n", "c", "h", "w"])]
    path = os.path.join(tmp, f"cut{cut}.onnx")
    save_model(g2, path)
    got = ob.Model(path).run(x)
    r64, r32 = ref(cut, torch.float64), ref(cut, torch.float32)
    scale = np.abs(r64).max()
    e_tc, e_32 = np.abs(got - r64), np.abs(r32 - r64)
    print(f"after node {cut:2d} {out_name:14s} shape {got.shape} max|y|={scale:9.3f}  "
          f"tc: max {e_tc.max():.3e} mean {e_tc.mean():.3e} bias {np.mean(got - r64):+.3e} | "
          f"f32: max {e_32.max():.3e} mean {e_32.mean():.3e} bias {np.mean(r32 - r64):+.3e}")
