"""The oracle's ONNX interpreter vs the torch modules the fixtures were exported from (checks the
hand-rolled protobuf codec, BN folding, GRU gate reordering and every operator on the two graphs)."""
import os

import numpy as np
import pytest
import torch

from oracle.onnx_eval import OnnxModel
from oracle.onnx_io import decode_model, encode_model, load_model
from tools.models import (
    DetectionNet, RecognitionNet, export_detection, export_recognition, init_synthetic,
)


@pytest.fixture(scope="module")
def tmp_models(tmp_path_factory):
    d = tmp_path_factory.mktemp("models")
    det = init_synthetic(DetectionNet(), 11)
    rec = init_synthetic(RecognitionNet(), 12)
    det_p, rec_p = str(d / "det.onnx"), str(d / "rec.onnx")
    export_detection(det, det_p, in_hw=(160, 120))
    export_recognition(rec, rec_p)
    return det, rec, det_p, rec_p


def test_codec_roundtrip(tmp_models):
    _, _, det_p, rec_p = tmp_models
    for p in (det_p, rec_p):
        g = load_model(p)
        g2 = decode_model(encode_model(g))
        assert [n.op_type for n in g.nodes] == [n.op_type for n in g2.nodes]
        assert all((g.initializers[k] == g2.initializers[k]).all() for k in g.initializers)
        assert [n.attrs.keys() for n in g.nodes] == [n.attrs.keys() for n in g2.nodes]


def test_detection_graph_matches_torch(tmp_models):
    det, _, det_p, _ = tmp_models
    m = OnnxModel(det_p)
    assert m.input_shape() == ["batch", 1, 160, 120]
    x = np.random.default_rng(0).uniform(-0.5, 0.5, (2, 1, 160, 120)).astype(np.float32)
    with torch.no_grad():
        ref = det(torch.from_numpy(x)).numpy()
    got = m.run(x)
    assert got.shape == ref.shape == (2, 1, 160, 120)
    assert np.abs(got - ref).max() < 2e-4
    ops = {n.op_type for n in m.graph.nodes}
    assert {"Conv", "ConvTranspose", "Pad", "Concat", "Relu", "Sigmoid"} <= ops


def test_recognition_graph_matches_torch(tmp_models):
    _, rec, _, rec_p = tmp_models
    m = OnnxModel(rec_p)
    assert m.input_shape() == ["batch", 1, 64, "seq"]
    x = np.random.default_rng(1).uniform(-0.5, 0.5, (3, 1, 64, 100)).astype(np.float32)
    with torch.no_grad():
        ref = rec(torch.from_numpy(x)).numpy()
    got = m.run(x)
    assert got.shape == ref.shape == (25, 3, 97)
    assert np.abs(got - ref).max() < 2e-4
    assert np.allclose(np.exp(got).sum(-1), 1.0, atol=1e-4)
    ops = {n.op_type for n in m.graph.nodes}
    assert {"Conv", "MaxPool", "AveragePool", "Reshape", "Transpose", "GRU", "Shape", "Gather", "Unsqueeze",
            "Concat", "ConstantOfShape", "MatMul", "Add", "LogSoftmax"} <= ops


def test_fused_gru_equals_spec_loop(tmp_models):
    """The ATen-backed GRU used for speed matches the explicit ONNX-spec recurrence."""
    _, _, _, rec_p = tmp_models
    x = np.random.default_rng(2).uniform(-0.5, 0.5, (2, 1, 64, 200)).astype(np.float32)
    a = OnnxModel(rec_p, fused_gru=True).run(x)
    b = OnnxModel(rec_p, fused_gru=False).run(x)
    assert np.abs(a - b).max() < 1e-4
