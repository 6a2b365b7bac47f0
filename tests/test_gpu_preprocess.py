"""prepare_image / resize kernels vs the oracle: bit-exact."""
import numpy as np
import pytest

import ocrs_b200 as ob
from oracle.imageops import prepare_image
from tests.gpu_util import fake_paths

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    return ob.OcrEngine(ob.OcrEngineParams())


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("order", ["hwc", "chw"])
@pytest.mark.parametrize("chans", [1, 3, 4])
@pytest.mark.parametrize("hw", [(1, 1), (7, 13), (64, 100), (768, 1024)])
def test_prepare_image_bit_exact(engine, dtype, order, chans, hw):
    rng = np.random.default_rng(hash((str(dtype), order, chans, hw)) % 2**32)
    h, w = hw
    shape = (h, w, chans) if order == "hwc" else (chans, h, w)
    if dtype == np.uint8:
        arr = rng.integers(0, 256, shape, dtype=np.uint8)
    else:
        arr = rng.random(shape, dtype=np.float32)
    got = engine.prepare_input(ob.ImageSource.from_tensor(arr, ob.DimOrder.Hwc if order == "hwc" else ob.DimOrder.Chw)).image()
    exp = prepare_image(arr, order)
    assert got.shape == exp.shape
    assert np.array_equal(got, exp)


def test_prepare_image_reference_values(engine):
    """preprocess.rs:378-594 known answers, to 1e-5 as in the reference."""
    res = engine.prepare_input(ob.ImageSource.from_tensor(np.array([0, 128, 255, 64], np.uint8).reshape(2, 2, 1), ob.DimOrder.Hwc)).image()
    assert np.abs(res.reshape(-1) - (np.array([0, 128, 255, 64]) / 255 - 0.5)).max() < 1e-5
    res = engine.prepare_input(ob.ImageSource.from_tensor(np.array([50, 100, 150, 255], np.uint8).reshape(1, 1, 4), ob.DimOrder.Hwc)).image()
    assert abs(float(res[0, 0, 0]) - (-0.5 + (50 * 0.299 + 100 * 0.587 + 150 * 0.114) / 255)) < 1e-5


def test_from_bytes_errors(engine):
    from ocrs_b200 import _lib
    import ctypes as C
    out = C.c_void_p()
    buf = np.arange(50, dtype=np.uint8)
    rc = _lib.lib.ocrs_b200_engine_prepare_input_bytes(engine._h, buf.ctypes.data_as(C.c_void_p), 50, 10, 10, C.byref(out))
    assert rc == _lib.ERR_INVALID_DATA_LENGTH
    buf = np.arange(128, dtype=np.uint8)
    rc = _lib.lib.ocrs_b200_engine_prepare_input_bytes(engine._h, buf.ctypes.data_as(C.c_void_p), 128, 8, 8, C.byref(out))
    assert rc == _lib.ERR_UNSUPPORTED_CHANNEL_COUNT
    rc = _lib.lib.ocrs_b200_engine_prepare_input_bytes(engine._h, buf.ctypes.data_as(C.c_void_p), 0, 0, 10, C.byref(out))
    assert rc == _lib.ERR_UNSUPPORTED_CHANNEL_COUNT


@pytest.mark.parametrize("hw", [(100, 200), (480, 640), (250, 90), (37, 411)])
def test_detect_text_pixels_resize_chain_bit_exact(tmp_path, hw):
    """pad -> resize -> (+0.5 model) -> slice -> resize (detection.rs:155-197) is bit-identical to the
    oracle: the fake model adds 0.5 exactly, so any difference would come from the resize kernels."""
    from oracle.engine import TextDetector
    from tests.fakes import FakeDetectionModel
    det, _ = fake_paths(tmp_path)
    rng = np.random.default_rng(3)
    img = rng.random((1, hw[0], hw[1]), dtype=np.float32)
    engine = ob.OcrEngine(ob.OcrEngineParams(detection_model=det))
    inp = engine.prepare_input(ob.ImageSource.from_tensor(img, ob.DimOrder.Chw))
    got = engine.detect_text_pixels(inp)
    exp = TextDetector(FakeDetectionModel()).detect_text_pixels(prepare_image(img, "chw"))
    assert np.array_equal(got, exp)
