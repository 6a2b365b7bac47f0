"""The reference's own integration goldens (lib.rs:437-577), run through the CUDA engine: the
fake models are expressed as ONNX graphs, loaded by the product loader and executed on the GPU."""
import numpy as np
import pytest

import ocrs_b200 as ob
from ocrs_b200 import _lib
from tests.fakes import gen_test_image
from tests.gpu_util import fake_paths

pytestmark = pytest.mark.gpu


def _bounding(r: ob.RotatedRect):
    from oracle.geometry import RotatedRect as ORect
    b = ORect.from_raw(*r.raw()).bounding_rect()
    return b.tlbr()


def test_ocr_engine_prepare_input(tmp_path):
    det, _ = fake_paths(tmp_path)
    image = gen_test_image(3)
    engine = ob.OcrEngine(ob.OcrEngineParams(detection_model=det))
    inp = engine.prepare_input(ob.ImageSource.from_tensor(image, ob.DimOrder.Chw))
    assert inp.shape == (1, image.shape[1], image.shape[2])


def test_ocr_engine_detect_words(tmp_path):
    """lib.rs:465-488: exact word boxes through pad+resize+threshold+contours+min-rect+expand."""
    det, _ = fake_paths(tmp_path)
    engine = ob.OcrEngine(ob.OcrEngineParams(detection_model=det))
    inp = engine.prepare_input(ob.ImageSource.from_tensor(gen_test_image(3), ob.DimOrder.Chw))
    words = engine.detect_words(inp)
    assert len(words) == 3
    boxes = sorted((_bounding(w) for w in words), key=lambda b: (int(b[0]), int(b[1])))
    top, height = 27, 25
    expected = [(top, -3, top + height, 53), (top, 66, top + height, 123), (top, 136, top + height, 193)]
    assert boxes == [tuple(float(v) for v in e) for e in expected]


def _recognize(engine, image):
    inp = engine.prepare_input(ob.ImageSource.from_tensor(image, ob.DimOrder.Chw))
    h, w = image.shape[1:]
    line = [ob.RotatedRect(w / 2, h / 2, 0.0, 1.0, float(w), float(h))]  # RotatedRect::from_rect
    lines = engine.recognize_text(inp, [line])
    assert len(lines) == 1 and lines[0] is not None
    return str(lines[0])


def test_ocr_engine_recognize_lines(tmp_path):
    """lib.rs:526-544"""
    _, rec = fake_paths(tmp_path)
    image = np.zeros((1, 64, 32), dtype=np.float32)
    image[:, 2, :] = 1.0
    engine = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec, alphabet=ob.DEFAULT_ALPHABET[:63]))
    assert _recognize(engine, image) == "0"


def test_ocr_engine_filter_chars(tmp_path):
    """lib.rs:546-577"""
    _, rec = fake_paths(tmp_path)
    image = np.zeros((1, 64, 32), dtype=np.float32)
    image[:, 2, :] = 0.7
    image[:, 3, :] = 0.3
    alphabet = ob.DEFAULT_ALPHABET[:63]
    engine = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec, alphabet=alphabet))
    assert _recognize(engine, image) == "0"
    engine = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec, alphabet=alphabet, allowed_chars="123456789"))
    assert _recognize(engine, image) == "1"


def test_errors_mirror_reference(tmp_path):
    det, rec = fake_paths(tmp_path)
    engine = ob.OcrEngine(ob.OcrEngineParams())
    inp = engine.prepare_input(ob.ImageSource.from_tensor(np.zeros((1, 8, 8), np.float32), ob.DimOrder.Chw))
    with pytest.raises(ob.OcrsError, match="Detection model not loaded") as ei:
        engine.detect_words(inp)
    assert ei.value.code == _lib.ERR_MODEL_NOT_LOADED
    with pytest.raises(ob.OcrsError, match="Recognition model not loaded"):
        engine.recognize_text(inp, [])
    assert engine.detection_threshold() == pytest.approx(0.2)
    # alphabet / class-count mismatch -> WrongOutput (recognition.rs:487-493)
    engine = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec))  # default alphabet = 97 classes, model has 64
    img = np.zeros((1, 64, 32), np.float32)
    inp = engine.prepare_input(ob.ImageSource.from_tensor(img, ob.DimOrder.Chw))
    with pytest.raises(ob.OcrsError, match="does not match alphabet size") as ei:
        engine.recognize_text(inp, [[ob.RotatedRect(16, 32, 0, 1, 32, 64)]])
    assert ei.value.code == _lib.ERR_WRONG_OUTPUT
    with pytest.raises(ob.OcrsError) as ei:
        ob.Model(b"RTEN\x02\x00\x00\x00")
    assert ei.value.code == _lib.ERR_MODEL_LOAD
