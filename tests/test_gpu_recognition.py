"""Recognition stages on the GPU vs the oracle: polygon crop + resize (bit-exact), the recognition
network (1e-3 on log-probs), CTC greedy and the step -> character-box mapping (exact)."""
import numpy as np
import pytest

import ocrs_b200 as ob
from oracle.engine import OcrEngine as OEngine, OcrEngineParams as OParams
from oracle.geometry import F, PointF, RotatedRect as ORect, Vec2
from oracle.onnx_eval import OnnxModel
from oracle.recognition import ctc_decode_greedy
from tests.gpu_util import (
    char_boxes, fake_paths, model_paths, oracle_char_boxes, oracle_text_of, text_of, to_oracle_rects,
)
from tools.synth import make_page

pytestmark = pytest.mark.gpu


def _words_line(rng, y, x0, n, rotated=False):
    words = []
    x = x0
    for _ in range(n):
        w = float(rng.uniform(20, 90))
        h = float(rng.uniform(14, 26))
        if rotated:
            a = float(rng.uniform(-0.2, 0.2))
            up = Vec2(np.sin(a), np.cos(a))
        else:
            up = Vec2(0.0, 1.0)
        words.append(ORect(PointF(F(x + w / 2), F(y + rng.uniform(-1.5, 1.5))), up, F(w), F(h)))
        x += w + float(rng.uniform(3, 12))
    return words


@pytest.fixture(scope="module")
def rec_engine():
    _, rec = model_paths()
    page, _ = make_page(3, 384, 640)
    eng = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec))
    inp = eng.prepare_input(ob.ImageSource.from_tensor(page, ob.DimOrder.Hwc))
    ora = OEngine(OParams(recognition_model=OnnxModel(rec)))
    return eng, inp, ora, ora.prepare_input(page, "hwc")


@pytest.mark.parametrize("seed", range(6))
def test_line_crop_bit_exact(rec_engine, seed):
    """recognition.rs:91-126 incl. lines that leave the page and rotated words."""
    eng, inp, ora, oimg = rec_engine
    rng = np.random.default_rng(seed)
    x0 = float(rng.choice([-12.0, 5.0, 200.0, 520.0]))
    y = float(rng.choice([4.0, 100.0, 380.0]))
    words = _words_line(rng, y, x0, int(rng.integers(1, 7)), rotated=bool(seed % 2))
    got = eng.prepare_recognition_input(inp, [ob.RotatedRect(*w.raw()) for w in words])
    exp = ora.prepare_recognition_input(oimg, words)
    assert got.shape == exp.shape
    assert np.array_equal(got, exp)


def test_recognize_text_matches_oracle(rec_engine):
    eng, inp, ora, oimg = rec_engine
    rng = np.random.default_rng(99)
    lines = [_words_line(rng, 30.0 + 40 * i, float(rng.uniform(5, 60)), int(rng.integers(1, 8))) for i in range(8)]
    lines.append(_words_line(rng, 200.0, 8.0, 9))  # long line -> wide bucket
    collect = []
    exp = ora.recognize_text(oimg, lines, collect=collect)
    got = eng.recognize_text(inp, [[ob.RotatedRect(*w.raw()) for w in l] for l in lines])
    assert text_of(got) == oracle_text_of(exp)
    assert char_boxes(got) == oracle_char_boxes(exp)


def test_ctc_greedy_through_fake_model(tmp_path):
    """CTC conventions (blank 0, repeat collapse, re-arming by blanks) with controlled scores: the
    fake model (lib.rs:372-422) turns image rows into class scores."""
    _, rec = fake_paths(tmp_path)
    alphabet = ob.DEFAULT_ALPHABET[:63]
    rng = np.random.default_rng(5)
    W = 200
    labels = [0, 5, 5, 0, 5, 7, 7, 7, 0, 0, 9, 1, 1, 0, 1] + [0] * 35   # one label per 4-px step (T = 50)
    img = np.zeros((1, 64, W), np.float32)
    for t, l in enumerate(labels):
        img[0, :, 4 * t:4 * t + 4] = rng.uniform(0.0, 0.3, (64, 1))
        img[0, l, 4 * t:4 * t + 4] = 0.9
    eng = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec, alphabet=alphabet))
    inp = eng.prepare_input(ob.ImageSource.from_tensor(img, ob.DimOrder.Chw))
    line = [ob.RotatedRect(W / 2, 32.0, 0.0, 1.0, float(W), 64.0)]
    got = eng.recognize_text(inp, [line])
    from tests.fakes import FakeRecognitionModel
    ora = OEngine(OParams(recognition_model=FakeRecognitionModel(), alphabet=alphabet))
    exp = ora.recognize_text(ora.prepare_input(img, "chw"), [[ORect.from_raw(*line[0].raw())]])
    assert text_of(got) == oracle_text_of(exp) == [alphabet[4] + alphabet[4] + alphabet[6] + alphabet[8] + alphabet[0] + alphabet[0]]
    assert char_boxes(got) == oracle_char_boxes(exp)


def test_empty_and_none_lines(rec_engine):
    eng, inp, ora, oimg = rec_engine
    assert eng.recognize_text(inp, []) == []
    with pytest.raises(ob.OcrsError, match="line has no words"):
        eng.recognize_text(inp, [[]])
