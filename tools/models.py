"""Synthetic stand-ins for the reference's two networks + a hand-rolled ONNX exporter.

The real `text-detection.rten` / `text-recognition.rten` are downloaded from S3 by the reference
CLI (ocrs-cli/src/main.rs:305-309) and are not available offline.  These torch modules follow the
publicly documented architectures of the ocrs-models project as recalled in SURVEY.md App. A.5
(U-Net with depthwise-separable double convs; CRNN with 2-layer bidirectional GRU) and are
exported with the operator vocabulary the reference registers for its models
(ocrs/src/wasm_api.rs:35-56), so loader, oracle and CUDA engine all consume the same `.onnx` file.
The engine itself is graph-driven and never assumes these shapes.

Fixture tooling: not part of the product path.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.onnx_io import FLOAT, Graph, Node, ValueInfo, _c, save_model  # noqa: E402

DET_INPUT_HW = (800, 600)  # lib.rs:347
REC_INPUT_H = 64  # lib.rs:366
NUM_CLASSES = 97  # 96-char alphabet + CTC blank (lib.rs:34)


# ---------------------------------------------------------------------------------------------
class DepthwiseSeparable(nn.Module):
    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.dw = nn.Conv2d(cin, cin, 3, stride=stride, padding=1, groups=cin)
        self.pw = nn.Conv2d(cin, cout, 1)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return torch.relu(self.bn(self.pw(self.dw(x))))


class DoubleConv(nn.Module):
    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.a = DepthwiseSeparable(cin, cout, stride)
        self.b = DepthwiseSeparable(cout, cout, 1)

    def forward(self, x):
        return self.b(self.a(x))


class DetectionNet(nn.Module):
    """U-Net: 7 stride-2 levels [8,16,32,32,64,128,256], ConvTranspose(k2,s2) + crop + concat +
    DoubleConv on the way up, final ConvTranspose to full resolution, 1x1 conv, sigmoid."""

    def __init__(self, depths=(8, 16, 32, 32, 64, 128, 256)):
        super().__init__()
        self.depths = list(depths)
        self.down = nn.ModuleList()
        prev = 1
        for d in self.depths:
            self.down.append(DoubleConv(prev, d, stride=2))
            prev = d
        self.up_t = nn.ModuleList()
        self.up_c = nn.ModuleList()
        for d in reversed(self.depths[:-1]):
            self.up_t.append(nn.ConvTranspose2d(prev, d, 2, stride=2))
            self.up_c.append(DoubleConv(2 * d, d))
            prev = d
        self.final_t = nn.ConvTranspose2d(prev, prev, 2, stride=2)
        self.final_c = nn.Conv2d(prev, 1, 1)

    def forward(self, x):
        skips = []
        for blk in self.down:
            x = blk(x)
            skips.append(x)
        skips.pop()
        for t, c in zip(self.up_t, self.up_c):
            skip = skips.pop()
            x = t(x)
            x = x[:, :, : skip.shape[2], : skip.shape[3]]
            x = c(torch.cat([x, skip], dim=1))
        x = torch.relu(self.final_t(x))
        return torch.sigmoid(self.final_c(x))


class RecognitionNet(nn.Module):
    """CRNN: 6 conv3x3 (+pool) -> avg-pool H -> [T,N,128] -> 2x bi-GRU(256) -> Linear(97) -> log-softmax."""

    def __init__(self, num_classes=NUM_CLASSES, hidden=256, chans=(32, 64, 128, 128, 128, 128)):
        super().__init__()
        c = list(chans)
        self.convs = nn.ModuleList([
            nn.Conv2d(1, c[0], 3, padding=1), nn.Conv2d(c[0], c[1], 3, padding=1),
            nn.Conv2d(c[1], c[2], 3, padding=1), nn.Conv2d(c[2], c[3], 3, padding=1),
            nn.Conv2d(c[3], c[4], 3, padding=1), nn.Conv2d(c[4], c[5], 3, padding=1),
        ])
        self.bns = nn.ModuleDict({"2": nn.BatchNorm2d(c[2]), "4": nn.BatchNorm2d(c[4])})
        self.pools = {0: (2, 2), 1: (2, 2), 3: (2, 1), 5: (2, 1)}
        self.gru = nn.GRU(c[5], hidden, num_layers=2, bidirectional=True)
        self.fc = nn.Linear(2 * hidden, num_classes)
        self.hidden = hidden

    def features(self, x):
        for i, conv in enumerate(self.convs):
            x = conv(x)
            if str(i) in self.bns:
                x = self.bns[str(i)](x)
            x = torch.relu(x)
            if i in self.pools:
                x = nn.functional.max_pool2d(x, self.pools[i])
        x = nn.functional.avg_pool2d(x, (x.shape[2], 1))
        return x.squeeze(2).permute(2, 0, 1)  # [T, N, C]

    def forward(self, x):
        seq = self.features(x)
        y, _ = self.gru(seq)
        return torch.log_softmax(self.fc(y), dim=2)


def init_synthetic(model: nn.Module, seed: int) -> nn.Module:
    """Seeded He-normal weights scaled so activations stay O(1); BN stats non-trivial."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                fan_in = m.weight[0].numel() if isinstance(m, nn.Conv2d) else m.weight.shape[0] * 1
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / max(fan_in, 1)) ** 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.05 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(1.0 + 0.1 * torch.rand(m.running_var.shape, generator=g))
            elif isinstance(m, nn.GRU):
                for name, p in m.named_parameters():
                    k = (1.0 / m.hidden_size) ** 0.5
                    p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * k)
            elif isinstance(m, nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (1.0 / m.in_features) ** 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
    return model.eval()


# ---------------------------------------------------------------------------------------------
# ONNX export (hand-written; mirrors what torch.onnx.export emits for these modules)
# ---------------------------------------------------------------------------------------------
class _Builder:
    def __init__(self):
        self.nodes: List[Node] = []
        self.inits: Dict[str, np.ndarray] = {}
        self._n = 0

    def name(self, prefix: str) -> str:
        self._n += 1
        return f"{prefix}_{self._n}"

    def const(self, prefix: str, arr) -> str:
        nm = self.name(prefix)
        self.inits[nm] = _c(arr)
        return nm

    def node(self, op: str, inputs: List[str], attrs=None, n_out: int = 1, prefix=None):
        outs = [self.name(prefix or op.lower()) for _ in range(n_out)]
        self.nodes.append(Node(op, list(inputs), outs, dict(attrs or {}), name=self.name("n_" + op)))
        return outs[0] if n_out == 1 else outs


def _fold_bn(w: torch.Tensor, b: torch.Tensor, bn: nn.BatchNorm2d):
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return w * scale.reshape(-1, 1, 1, 1), (b - bn.running_mean) * scale + bn.bias


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().astype(np.float32)


def _conv(b: _Builder, x: str, w, bias, *, stride=1, pad=0, group=1, k=None) -> str:
    kh, kw = w.shape[2], w.shape[3]
    return b.node("Conv", [x, b.const("w", _np(w)), b.const("b", _np(bias))], {
        "dilations": [1, 1], "group": int(group), "kernel_shape": [kh, kw],
        "pads": [pad, pad, pad, pad], "strides": [stride, stride]})


def _dws(b: _Builder, x: str, m: DepthwiseSeparable) -> str:
    x = _conv(b, x, m.dw.weight, m.dw.bias, stride=m.dw.stride[0], pad=1, group=m.dw.groups)
    w, bias = _fold_bn(m.pw.weight, m.pw.bias, m.bn)
    x = _conv(b, x, w, bias)
    return b.node("Relu", [x])


def export_detection(model: DetectionNet, path: str, in_hw=DET_INPUT_HW) -> None:
    model = model.eval()
    b = _Builder()
    x = "image"
    h, w = in_hw
    skips = []
    sizes = []
    for blk in model.down:
        x = _dws(b, _dws(b, x, blk.a), blk.b)
        h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        skips.append(x)
        sizes.append((h, w))
    skips.pop()
    sizes.pop()
    for t, c in zip(model.up_t, model.up_c):
        skip = skips.pop()
        sh, sw = sizes.pop()
        x = b.node("ConvTranspose", [x, b.const("w", _np(t.weight)), b.const("b", _np(t.bias))], {
            "dilations": [1, 1], "group": 1, "kernel_shape": [2, 2], "pads": [0, 0, 0, 0], "strides": [2, 2]})
        h, w = 2 * h, 2 * w
        if (h, w) != (sh, sw):
            pads = np.array([0, 0, 0, 0, 0, 0, sh - h, sw - w], dtype=np.int64)  # negative = crop
            x = b.node("Pad", [x, b.const("pads", pads), b.const("pad_value", np.zeros((), np.float32))], {"mode": "constant"})
            h, w = sh, sw
        x = b.node("Concat", [x, skip], {"axis": 1})
        x = _dws(b, _dws(b, x, c.a), c.b)
    x = b.node("ConvTranspose", [x, b.const("w", _np(model.final_t.weight)), b.const("b", _np(model.final_t.bias))], {
        "dilations": [1, 1], "group": 1, "kernel_shape": [2, 2], "pads": [0, 0, 0, 0], "strides": [2, 2]})
    x = b.node("Relu", [x])
    x = _conv(b, x, model.final_c.weight, model.final_c.bias)
    x = b.node("Sigmoid", [x])
    b.nodes[-1].outputs[0] = "mask"
    g = Graph(b.nodes, b.inits,
              [ValueInfo("image", FLOAT, ["batch", 1, in_hw[0], in_hw[1]])],
              [ValueInfo("mask", FLOAT, ["batch", 1, in_hw[0], in_hw[1]])], name="text-detection")
    save_model(g, path)


def _gru_params(gru: nn.GRU, layer: int):
    """torch gate order (r, z, n) -> ONNX (z, r, h); returns W [2,3H,I], R [2,3H,H], B [2,6H]."""
    H = gru.hidden_size

    def reorder(t):
        r, z, n = t[:H], t[H:2 * H], t[2 * H:]
        return torch.cat([z, r, n], 0)

    Ws, Rs, Bs = [], [], []
    for suffix in ("", "_reverse"):
        Ws.append(reorder(getattr(gru, f"weight_ih_l{layer}{suffix}")))
        Rs.append(reorder(getattr(gru, f"weight_hh_l{layer}{suffix}")))
        Bs.append(torch.cat([reorder(getattr(gru, f"bias_ih_l{layer}{suffix}")),
                             reorder(getattr(gru, f"bias_hh_l{layer}{suffix}"))], 0))
    return _np(torch.stack(Ws)), _np(torch.stack(Rs)), _np(torch.stack(Bs))


def export_recognition(model: RecognitionNet, path: str) -> None:
    model = model.eval()
    b = _Builder()
    x = "line_images"
    for i, conv in enumerate(model.convs):
        w, bias = conv.weight, conv.bias
        if str(i) in model.bns:
            w, bias = _fold_bn(w, bias, model.bns[str(i)])
        x = _conv(b, x, w, bias, pad=1)
        x = b.node("Relu", [x])
        if i in model.pools:
            k = list(model.pools[i])
            x = b.node("MaxPool", [x], {"kernel_shape": k, "strides": k, "pads": [0, 0, 0, 0], "ceil_mode": 0})
    # after pools H = 64 / 16 = 4
    x = b.node("AveragePool", [x], {"kernel_shape": [4, 1], "strides": [4, 1], "pads": [0, 0, 0, 0], "ceil_mode": 0})
    x = b.node("Reshape", [x, b.const("shape", np.array([0, 0, -1], dtype=np.int64))])  # [N,C,T]
    x = b.node("Transpose", [x], {"perm": [2, 0, 1]})  # [T,N,C]
    H = model.hidden
    for layer in range(model.gru.num_layers):
        W, R, B = _gru_params(model.gru, layer)
        # h0 = zeros([2, N, H]) built from the runtime batch size, as torch.onnx does
        shp = b.node("Shape", [x])
        n = b.node("Gather", [shp, b.const("idx", np.array(1, dtype=np.int64))], {"axis": 0})
        n = b.node("Unsqueeze", [n, b.const("axes", np.array([0], dtype=np.int64))])
        h0_shape = b.node("Concat", [b.const("two", np.array([2], np.int64)), n, b.const("hid", np.array([H], np.int64))], {"axis": 0})
        h0 = b.node("ConstantOfShape", [h0_shape], {"value": np.zeros((1,), np.float32)})
        y, _yh = b.node("GRU", [x, b.const("W", W), b.const("R", R), b.const("B", B), "", h0], {
            "hidden_size": H, "direction": "bidirectional", "linear_before_reset": 1}, n_out=2)
        y = b.node("Transpose", [y], {"perm": [0, 2, 1, 3]})  # [T,N,2,H]
        x = b.node("Reshape", [y, b.const("shape", np.array([0, 0, -1], dtype=np.int64))])  # [T,N,2H]
    x = b.node("MatMul", [x, b.const("fc_w", _np(model.fc.weight.t()))])
    x = b.node("Add", [x, b.const("fc_b", _np(model.fc.bias))])
    x = b.node("LogSoftmax", [x], {"axis": 2})
    b.nodes[-1].outputs[0] = "log_probs"
    g = Graph(b.nodes, b.inits,
              [ValueInfo("line_images", FLOAT, ["batch", 1, REC_INPUT_H, "seq"])],
              [ValueInfo("log_probs", FLOAT, ["out_seq", "batch", model.fc.out_features])], name="text-recognition")
    save_model(g, path)


def default_model_dir() -> str:
    return os.path.join(ROOT, "models")


def ensure_models(model_dir: str | None = None, verbose: bool = False):
    """Returns (detection_path, recognition_path).  Trained fixtures under `models/` win when
    present (committed); otherwise seeded synthetic-weight models are written deterministically."""
    model_dir = model_dir or default_model_dir()
    os.makedirs(model_dir, exist_ok=True)
    det = os.path.join(model_dir, "text-detection.onnx")
    rec = os.path.join(model_dir, "text-recognition.onnx")
    if not os.path.exists(det):
        if verbose:
            print("writing", det)
        export_detection(init_synthetic(DetectionNet(), 1001), det)
    if not os.path.exists(rec):
        if verbose:
            print("writing", rec)
        export_recognition(init_synthetic(RecognitionNet(), 1002), rec)
    return det, rec


if __name__ == "__main__":
    print(ensure_models(sys.argv[1] if len(sys.argv) > 1 else None, verbose=True))


# ---------------------------------------------------------------------------------------------
# The reference's fake models (lib.rs:335-422) expressed as ONNX graphs, so that its known-answer
# integration tests run through the real loader + CUDA executor.
# ---------------------------------------------------------------------------------------------
def export_fake_detection(path: str, in_hw=(200, 100)) -> None:
    """lib.rs:339-362: output = input + 0.5, input [batch,1,H,W]."""
    b = _Builder()
    b.node("Add", ["image", b.const("half", np.array([0.5], dtype=np.float32))])
    b.nodes[-1].outputs[0] = "mask"
    g = Graph(b.nodes, b.inits, [ValueInfo("image", FLOAT, ["batch", 1, in_hw[0], in_hw[1]])],
              [ValueInfo("mask", FLOAT, ["batch", 1, in_hw[0], in_hw[1]])], name="fake-detection")
    save_model(g, path)


def export_fake_recognition(path: str, height: int = 64) -> None:
    """lib.rs:372-422: max-pool width by 4, reinterpret the 64 rows as class scores -> [W/4, N, 64]."""
    b = _Builder()
    x = b.node("MaxPool", ["line_images"], {"kernel_shape": [1, 4], "strides": [1, 4], "pads": [0, 0, 0, 0], "ceil_mode": 0})
    x = b.node("Reshape", [x, b.const("shape", np.array([0, height, -1], dtype=np.int64))])
    x = b.node("Transpose", [x], {"perm": [2, 0, 1]})
    b.nodes[-1].outputs[0] = "scores"
    g = Graph(b.nodes, b.inits, [ValueInfo("line_images", FLOAT, ["batch", 1, height, "seq"])],
              [ValueInfo("scores", FLOAT, ["out_seq", "batch", height])], name="fake-recognition")
    save_model(g, path)
