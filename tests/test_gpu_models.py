"""Inner seam (`trait Model`): the CUDA graph executor vs the torch-CPU ONNX interpreter on the
same model files.  Tolerance: 1e-3 absolute on recognition log-probs (BASELINE.json north_star),
1e-4 on detection probabilities."""
import numpy as np
import pytest

import ocrs_b200 as ob
from oracle.onnx_eval import OnnxModel
from tests.gpu_util import model_paths

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3
PROB_TOL = 1e-3


@pytest.fixture(scope="module")
def models():
    det, rec = model_paths()
    return ob.Model(det), ob.Model(rec), OnnxModel(det), OnnxModel(rec)


def test_input_shapes(models):
    det, rec, odet, orec = models
    assert det.input_shape() == ["sym", 1, 800, 600]
    assert rec.input_shape() == ["sym", 1, 64, "sym"]


@pytest.mark.parametrize("batch", [1, 3])
def test_detection_net(models, batch):
    det, _, odet, _ = models
    x = np.random.default_rng(batch).uniform(-0.5, 0.5, (batch, 1, 800, 600)).astype(np.float32)
    got, exp = det.run(x), odet.run(x)
    assert got.shape == exp.shape == (batch, 1, 800, 600)
    assert np.abs(got - exp).max() < PROB_TOL
    assert det.last_flops() > 1e8 * batch


@pytest.mark.parametrize("batch,width", [(1, 50), (2, 100), (5, 400), (2, 1200), (1, 2400)])
def test_recognition_net(models, batch, width):
    _, rec, _, orec = models
    x = np.random.default_rng(width).uniform(-0.5, 0.5, (batch, 1, 64, width)).astype(np.float32)
    got, exp = rec.run(x), orec.run(x)
    assert got.shape == exp.shape == (width // 4, batch, 97)
    assert np.abs(got - exp).max() < LOGIT_TOL
    assert np.allclose(np.exp(got).sum(-1), 1.0, atol=1e-4)


def test_wrong_input_rank_fails(models):
    det, _, _, _ = models
    with pytest.raises(ob.OcrsError):
        det.run(np.zeros((1, 800, 600), np.float32))
    with pytest.raises(ob.OcrsError):
        det.run(np.zeros((1, 1, 100, 100), np.float32))


def test_concurrent_runs_on_one_handle(models):
    """recognition.rs:465-485 calls Model::run from several threads on one model."""
    import threading
    _, rec, _, orec = models
    xs = [np.random.default_rng(i).uniform(-0.5, 0.5, (2, 1, 64, 100 + 50 * i)).astype(np.float32) for i in range(4)]
    outs = [None] * 4

    def work(i):
        outs[i] = rec.run(xs[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for i in range(4):
        assert np.abs(outs[i] - orec.run(xs[i])).max() < LOGIT_TOL


def test_fp16_range_overflow_falls_back_to_fp32():
    """The tensor-core convolutions carry activations as split fp16.  An input that drives an
    activation past 65504 raises the device flag and the run is repeated on the fp32 kernels
    (never a silent inf/NaN); the model then stays on the fp32 path."""
    _, rec_path = model_paths()
    rec, orec = ob.Model(rec_path), OnnxModel(rec_path)
    rng = np.random.default_rng(0)
    big = (rng.uniform(-0.5, 0.5, (2, 1, 64, 100)) * 4e5).astype(np.float32)
    got, exp = rec.run(big), orec.run(big)
    assert np.isfinite(got).all()
    assert np.abs(got - exp).max() <= 2e-4 * np.abs(exp).max()
    x = rng.uniform(-0.5, 0.5, (2, 1, 64, 100)).astype(np.float32)
    assert np.abs(rec.run(x) - orec.run(x)).max() < 1e-4   # fp32 kernels from now on
