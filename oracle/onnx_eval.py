"""ONNX graph interpreter on torch-CPU fp32 (oracle; test infrastructure).

Stands in for `rten::Model::run_one` (model.rs:33-40) on the two networks.  The reference
executes them with rten 0.24.0 (not vendored, no weights in this environment), so each operator
follows the ONNX operator specification; torch-CPU supplies the fp32 arithmetic.  Supported:
the 20 operators the reference registers for its models (wasm_api.rs:35-56) plus Constant,
Squeeze, Identity and Tanh.  "torch-CPU restatement, not rten."
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as Fnn

from .onnx_io import Graph, Node, _c, load_model

_TORCH_DT = {1: torch.float32, 2: torch.uint8, 3: torch.int8, 6: torch.int32, 7: torch.int64, 9: torch.bool}


def _ints(t) -> List[int]:
    return [int(v) for v in torch.as_tensor(t).reshape(-1).tolist()]


def _pool_pad(x: torch.Tensor, pads: Sequence[int], value: float) -> torch.Tensor:
    t, l, b, r = pads
    if any(pads):
        x = Fnn.pad(x, (l, r, t, b), value=value)
    return x


def _gru(X, W, R, B, initial_h, hidden_size: int, direction: str, linear_before_reset: int):
    """ONNX GRU, gate order z, r, h.  X [T,N,I] -> Y [T,D,N,H], Y_h [D,N,H]."""
    T, N, _ = X.shape
    D = W.shape[0]
    H = hidden_size
    ys = []
    yh = []
    for d in range(D):
        reverse = (direction == "reverse") or (direction == "bidirectional" and d == 1)
        Wd, Rd = W[d], R[d]
        Wb = B[d, : 3 * H] if B is not None else torch.zeros(3 * H)
        Rb = B[d, 3 * H:] if B is not None else torch.zeros(3 * H)
        h = initial_h[d] if initial_h is not None else torch.zeros(N, H)
        xw = X @ Wd.t() + Wb  # [T, N, 3H]
        out = [None] * T
        order = range(T - 1, -1, -1) if reverse else range(T)
        Rz, Rr, Rh = Rd[:H], Rd[H:2 * H], Rd[2 * H:]
        for t in order:
            g = xw[t]
            z = torch.sigmoid(g[:, :H] + h @ Rz.t() + Rb[:H])
            r = torch.sigmoid(g[:, H:2 * H] + h @ Rr.t() + Rb[H:2 * H])
            if linear_before_reset:
                n = torch.tanh(g[:, 2 * H:] + r * (h @ Rh.t() + Rb[2 * H:]))
            else:
                n = torch.tanh(g[:, 2 * H:] + (r * h) @ Rh.t() + Rb[2 * H:])
            h = (1.0 - z) * n + z * h
            out[t] = h
        ys.append(torch.stack(out, 0))
        yh.append(h)
    return torch.stack(ys, 1), torch.stack(yh, 0)


def _torch_gru_module(W, R, B, hidden_size: int, direction: str):
    """ONNX GRU (gate order z, r, h; linear_before_reset = 1) as a torch.nn.GRU (gate order r, z, n):
    the same recurrence evaluated by ATen's fused CPU kernel -- only used to make the oracle fast."""
    D, H = W.shape[0], hidden_size
    gru = torch.nn.GRU(W.shape[2], H, num_layers=1, bidirectional=(D == 2))

    def reorder(t):
        z, r, n = t[:H], t[H:2 * H], t[2 * H:]
        return torch.cat([r, z, n], 0)

    with torch.no_grad():
        for d in range(D):
            sfx = "_reverse" if d == 1 else ""
            getattr(gru, "weight_ih_l0" + sfx).copy_(reorder(W[d]))
            getattr(gru, "weight_hh_l0" + sfx).copy_(reorder(R[d]))
            if B is not None:
                getattr(gru, "bias_ih_l0" + sfx).copy_(reorder(B[d, :3 * H]))
                getattr(gru, "bias_hh_l0" + sfx).copy_(reorder(B[d, 3 * H:]))
            else:
                getattr(gru, "bias_ih_l0" + sfx).zero_()
                getattr(gru, "bias_hh_l0" + sfx).zero_()
    return gru.eval()


class OnnxModel:
    """`impl Model for rten::Model` stand-in: `input_shape()` + `run()` (model.rs:19-41)."""

    def __init__(self, graph_or_path, fused_gru: bool = True):
        self.graph: Graph = load_model(graph_or_path) if isinstance(graph_or_path, str) else graph_or_path
        self.consts = {k: torch.from_numpy(_c(v)) for k, v in self.graph.initializers.items()}
        self._gru_modules = {}
        if fused_gru:
            for i, n in enumerate(self.graph.nodes):
                if (n.op_type == "GRU" and n.attrs.get("linear_before_reset", 0) and len(n.inputs) >= 3
                        and n.inputs[1] in self.consts and n.inputs[2] in self.consts
                        and n.attrs.get("direction", "forward") in ("forward", "bidirectional")
                        and (len(n.inputs) < 4 or not n.inputs[3] or n.inputs[3] in self.consts)):
                    B = self.consts[n.inputs[3]] if len(n.inputs) > 3 and n.inputs[3] else None
                    self._gru_modules[id(n)] = _torch_gru_module(self.consts[n.inputs[1]], self.consts[n.inputs[2]], B,
                                                                 n.attrs["hidden_size"], n.attrs.get("direction", "forward"))

    def input_shape(self) -> List[Any]:
        """Fixed dims as int, symbolic dims as str (rten::Dimension)."""
        if not self.graph.inputs:
            raise ValueError("model has no inputs")
        return list(self.graph.inputs[0].shape)

    @torch.no_grad()
    def run(self, x: np.ndarray, return_all: bool = False):
        env: Dict[str, torch.Tensor] = dict(self.consts)
        env[self.graph.inputs[0].name] = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        for node in self.graph.nodes:
            outs = self._eval(node, [env[i] if i else None for i in node.inputs])
            if not isinstance(outs, (list, tuple)):
                outs = [outs]
            for name, v in zip(node.outputs, outs):
                if name:
                    env[name] = v
        if return_all:
            return {k: v for k, v in env.items() if k not in self.consts}
        return env[self.graph.outputs[0].name].contiguous().numpy()

    def _eval(self, n: Node, a: List[Optional[torch.Tensor]]):
        op, at = n.op_type, n.attrs
        if op == "Conv":
            assert at.get("auto_pad", "NOTSET") == "NOTSET"
            pads = at.get("pads", [0, 0, 0, 0])
            x = a[0]
            if pads[0] == pads[2] and pads[1] == pads[3]:
                padding = (pads[0], pads[1])
            else:
                x = Fnn.pad(x, (pads[1], pads[3], pads[0], pads[2]))
                padding = 0
            return Fnn.conv2d(x, a[1], a[2] if len(a) > 2 else None, stride=at.get("strides", [1, 1]),
                              padding=padding, dilation=at.get("dilations", [1, 1]), groups=at.get("group", 1))
        if op == "ConvTranspose":
            pads = at.get("pads", [0, 0, 0, 0])
            y = Fnn.conv_transpose2d(a[0], a[1], a[2] if len(a) > 2 else None, stride=at.get("strides", [1, 1]),
                                     padding=0, output_padding=at.get("output_padding", [0, 0]),
                                     groups=at.get("group", 1), dilation=at.get("dilations", [1, 1]))
            h, w = y.shape[-2:]
            return y[..., pads[0]:h - pads[2], pads[1]:w - pads[3]]
        if op == "MaxPool":
            assert not at.get("ceil_mode", 0)
            x = _pool_pad(a[0], at.get("pads", [0, 0, 0, 0]), float("-inf"))
            return Fnn.max_pool2d(x, at["kernel_shape"], at.get("strides", [1, 1]))
        if op == "AveragePool":
            assert not at.get("ceil_mode", 0)
            pads = at.get("pads", [0, 0, 0, 0])
            assert not any(pads) or at.get("count_include_pad", 0)
            x = _pool_pad(a[0], pads, 0.0)
            return Fnn.avg_pool2d(x, at["kernel_shape"], at.get("strides", [1, 1]))
        if op == "Relu":
            return torch.relu(a[0])
        if op == "Sigmoid":
            return torch.sigmoid(a[0])
        if op == "Tanh":
            return torch.tanh(a[0])
        if op == "Add":
            return a[0] + a[1]
        if op == "MatMul":
            return a[0] @ a[1]
        if op == "LogSoftmax":
            return torch.log_softmax(a[0], dim=at.get("axis", -1))
        if op == "Concat":
            return torch.cat([t for t in a], dim=at["axis"])
        if op == "Transpose":
            perm = at.get("perm", list(range(a[0].dim() - 1, -1, -1)))
            return a[0].permute(perm)
        if op == "Reshape":
            shape = _ints(a[1])
            if not at.get("allowzero", 0):
                shape = [a[0].shape[i] if s == 0 else s for i, s in enumerate(shape)]
            return a[0].reshape(shape)
        if op == "Identity":
            return a[0]
        if op == "Constant":
            return torch.from_numpy(_c(at["value"]))
        if op == "Shape":
            dims = list(a[0].shape)
            return torch.tensor(dims[at.get("start", 0): at.get("end", len(dims))], dtype=torch.int64)
        if op == "Gather":
            axis = at.get("axis", 0)
            idx = a[1].to(torch.int64)
            idx = torch.where(idx < 0, idx + a[0].shape[axis], idx)
            if idx.dim() == 0:
                return a[0].select(axis, int(idx))
            return torch.index_select(a[0], axis, idx.reshape(-1)).reshape(
                list(a[0].shape[:axis]) + list(idx.shape) + list(a[0].shape[axis + 1:]))
        if op == "Unsqueeze":
            axes = _ints(a[1]) if len(a) > 1 and a[1] is not None else at["axes"]
            out = a[0]
            nd = out.dim() + len(axes)
            for ax in sorted(ax % nd for ax in axes):
                out = out.unsqueeze(ax)
            return out
        if op == "Squeeze":
            axes = _ints(a[1]) if len(a) > 1 and a[1] is not None else at.get("axes")
            out = a[0]
            if axes is None:
                return out.squeeze()
            for ax in sorted((ax % out.dim() for ax in axes), reverse=True):
                out = out.squeeze(ax)
            return out
        if op == "Slice":
            starts, ends = _ints(a[1]), _ints(a[2])
            axes = _ints(a[3]) if len(a) > 3 and a[3] is not None else list(range(len(starts)))
            steps = _ints(a[4]) if len(a) > 4 and a[4] is not None else [1] * len(starts)
            out = a[0]
            for s, e, ax, st in zip(starts, ends, axes, steps):
                assert st > 0
                dim = out.shape[ax]
                s = min(max(s + dim if s < 0 else s, 0), dim)
                e = min(max(e + dim if e < 0 else e, 0), dim)
                out = out.narrow(ax, s, max(e - s, 0))[(slice(None),) * (ax % out.dim()) + (slice(None, None, st),)]
            return out
        if op == "Cast":
            return a[0].to(_TORCH_DT[at["to"]])
        if op == "ConstantOfShape":
            val = at.get("value")
            shape = _ints(a[0])
            if val is None:
                return torch.zeros(shape, dtype=torch.float32)
            v = torch.from_numpy(_c(val)).reshape(-1)[0]
            return torch.full(shape, v.item(), dtype=v.dtype)
        if op == "Pad":
            assert at.get("mode", "constant") == "constant"
            pads = _ints(a[1]) if len(a) > 1 and a[1] is not None else at["pads"]
            value = float(a[2].reshape(-1)[0]) if len(a) > 2 and a[2] is not None and a[2].numel() else float(at.get("value", 0.0))
            nd = a[0].dim()
            axes = _ints(a[3]) if len(a) > 3 and a[3] is not None else list(range(nd))
            begin = [0] * nd
            end = [0] * nd
            k = len(axes)
            for i, ax in enumerate(axes):
                begin[ax % nd] = pads[i]
                end[ax % nd] = pads[i + k]
            out = a[0]
            # negative pads crop
            for ax in range(nd):
                lo = -begin[ax] if begin[ax] < 0 else 0
                hi = out.shape[ax] + (end[ax] if end[ax] < 0 else 0)
                if lo or hi != out.shape[ax]:
                    out = out.narrow(ax, lo, hi - lo)
            tp = []
            for ax in range(nd - 1, -1, -1):
                tp += [max(begin[ax], 0), max(end[ax], 0)]
            if any(tp):
                out = Fnn.pad(out, tp, value=value)
            return out
        if op == "GRU" and id(n) in self._gru_modules:
            assert a[4] is None if len(a) > 4 else True, "sequence_lens unsupported"
            mod = self._gru_modules[id(n)]
            h0 = a[5] if len(a) > 5 and a[5] is not None else None
            y, hn = mod(a[0], h0)
            T, N = a[0].shape[0], a[0].shape[1]
            D = 2 if mod.bidirectional else 1
            return y.reshape(T, N, D, mod.hidden_size).permute(0, 2, 1, 3).contiguous(), hn
        if op == "GRU":
            assert a[4] is None if len(a) > 4 else True, "sequence_lens unsupported"
            return _gru(a[0], a[1], a[2], a[3] if len(a) > 3 else None, a[5] if len(a) > 5 else None,
                        at["hidden_size"], at.get("direction", "forward"), at.get("linear_before_reset", 0))
        raise NotImplementedError(f"ONNX op {op}")
