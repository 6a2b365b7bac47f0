"""CTA-pair conv kernel (OCRS_B200_CONV_PAIR=1) against float64 on single Conv3x3 layers."""
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(path_in, path_out, onnx_path):
    import ocrs_b200 as ob
    x = np.load(path_in)
    np.save(path_out, ob.Model(onnx_path).run(x))


def main():
    import torch
    import torch.nn.functional as F
    from oracle.onnx_io import Graph, ValueInfo, save_model
    from tools.models import _Builder
    tmp = tempfile.mkdtemp()
    rng = np.random.default_rng(0)
    for (n, cin, h, w) in [(3, 128, 8, 100), (16, 128, 16, 100), (5, 64, 16, 37)]:
        wt = rng.normal(0, 1.0 / np.sqrt(9 * cin), (128, cin, 3, 3)).astype(np.float32)
        bias = rng.normal(0, 0.1, 128).astype(np.float32)
        x = rng.normal(0, 1, (n, cin, h, w)).astype(np.float32)
        b = _Builder()
        y = b.node("Conv", ["x", b.const("w", wt), b.const("b", bias)],
                   {"dilations": [1, 1], "group": 1, "kernel_shape": [3, 3], "pads": [1, 1, 1, 1], "strides": [1, 1]})
        y = b.node("Relu", [y])
        g = Graph(b.nodes, b.inits, [ValueInfo("x", 1, ["n", cin, h, w])], [ValueInfo(y, 1, ["n", 128, h, w])], name="t")
        onnx_path = os.path.join(tmp, "c.onnx")
        save_model(g, onnx_path)
        np.save(os.path.join(tmp, "x.npy"), x)
        ref = F.relu(F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(wt).double(), torch.from_numpy(bias).double(), padding=1)).numpy()
        for pair in ("0", "1"):
            env = dict(os.environ, OCRS_B200_CONV_PAIR=pair, OCRS_B200_CONV_PAIR_DEBUG="1")
            out = os.path.join(tmp, f"y{pair}.npy")
            r = subprocess.run([sys.executable, __file__, "--child", os.path.join(tmp, "x.npy"), out, onnx_path], env=env,
                               capture_output=True, text=True, timeout=120)
            msg = [l for l in r.stderr.splitlines() if "conv pair" in l or "rror" in l][-3:]
            if r.returncode != 0 or not os.path.exists(out):
                print(f"shape {(n, cin, h, w)} pair={pair}: FAILED rc={r.returncode}", msg)
                continue
            got = np.load(out)
            os.remove(out)
            e = np.abs(got - ref)
            print(f"shape {(n, cin, h, w)} pair={pair}: max err {e.max():.3e} mean {e.mean():.3e} (|y| max {np.abs(ref).max():.2f})", msg)


if __name__ == "__main__":
    if "--child" in sys.argv:
        i = sys.argv.index("--child")
        child(sys.argv[i + 1], sys.argv[i + 2], sys.argv[i + 3])
    else:
        main()
