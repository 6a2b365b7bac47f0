"""How does the tensor core round its fp32 accumulator?  One Conv3x3 (no ReLU) of Cin -> 64 channels on the
tcgen05 path against float64, error in units of ulp(|y|), split by the sign of y and by K."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrs_b200 as ob
from oracle.onnx_io import Graph, ValueInfo, save_model
from tools.models import _Builder

tmp = tempfile.mkdtemp()
rng = np.random.default_rng(0)
for cin in (32, 64, 128, 256):
    for mode in ("mixed", "positive"):
        w = rng.normal(0, 1.0 / np.sqrt(9 * cin), (64, cin, 3, 3)).astype(np.float32)
        x = rng.normal(0, 1, (4, cin, 16, 64)).astype(np.float32)
        if mode == "positive":
            w, x = np.abs(w), np.abs(x)
        b = _Builder()
        y = b.node("Conv", ["x", b.const("w", w), b.const("b", np.zeros(64, np.float32))],
                   {"dilations": [1, 1], "group": 1, "kernel_shape": [3, 3], "pads": [1, 1, 1, 1], "strides": [1, 1]})
        g = Graph(b.nodes, b.inits, [ValueInfo("x", 1, ["n", cin, 16, 64])], [ValueInfo(y, 1, ["n", 64, 16, 64])], name="t")
        path = os.path.join(tmp, f"c{cin}{mode}.onnx")
        save_model(g, path)
        got = ob.Model(path).run(x).astype(np.float64)
        ref = F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), padding=1).numpy()
        r32 = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), padding=1).numpy().astype(np.float64)
        ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
        e, e32 = (got - ref) / ulp, (r32 - ref) / ulp
        big = np.abs(ref) > 0.5 * np.abs(ref).mean()
        pos, neg = big & (ref > 0), big & (ref < 0)
        def st(v, m):
            return f"{v[m].mean():+7.2f} (sd {v[m].std():5.2f})" if m.any() else "   n/a"
        print(f"Cin {cin:3d} K {9*cin:4d} {mode:8s} err/ulp  y>0: tc {st(e,pos)} f32 {st(e32,pos)} | y<0: tc {st(e,neg)} f32 {st(e32,neg)}")
