"""`DecodeMethod::BeamSearch` (recognition.rs:199-205, 512-514) on the GPU vs the oracle's prefix
beam search.  The decoder is in the un-vendored rten-text crate (parity unpinned, see
oracle/__init__.py); these tests pin the CUDA kernel to the oracle restatement, including its
tie-breaking (creation order) and prefix merging."""
import numpy as np
import pytest

import ocrs_b200 as ob
from oracle.engine import OcrEngine as OEngine, OcrEngineParams as OParams
from oracle.geometry import RotatedRect as ORect
from oracle.onnx_eval import OnnxModel
from tests.fakes import FakeRecognitionModel
from tests.gpu_util import char_boxes, fake_paths, model_paths, oracle_char_boxes, oracle_text_of, text_of
from tools.synth import make_page

pytestmark = pytest.mark.gpu

ALPHABET = ob.DEFAULT_ALPHABET[:63]


def _run_both(tmp_path, img, lines, width, allowed=None):
    _, rec = fake_paths(tmp_path)
    eng = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec, alphabet=ALPHABET, allowed_chars=allowed,
                                          decode_method=ob.DecodeMethod.BeamSearch, beam_width=width))
    inp = eng.prepare_input(ob.ImageSource.from_tensor(img, ob.DimOrder.Chw))
    got = eng.recognize_text(inp, lines)
    ora = OEngine(OParams(recognition_model=FakeRecognitionModel(), alphabet=ALPHABET, allowed_chars=allowed,
                          decode_method="beam", beam_width=width))
    exp = ora.recognize_text(ora.prepare_input(img, "chw"), [[ORect.from_raw(*w.raw()) for w in l] for l in lines])
    return got, exp


def _full_line(W, H=64, x0=0.0):
    return [ob.RotatedRect(x0 + W / 2, H / 2, 0.0, 1.0, float(W), float(H))]


@pytest.mark.parametrize("width", [1, 4, 100])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_scores(tmp_path, seed, width):
    """Flat random scores: thousands of near-equal paths, heavy prefix merging."""
    rng = np.random.default_rng(seed)
    W = 100
    img = rng.uniform(0.0, 1.0, (1, 64, W)).astype(np.float32) * np.float32(6.0 if seed else 1.0)
    got, exp = _run_both(tmp_path, img, [_full_line(W)], width)
    assert text_of(got) == oracle_text_of(exp)
    assert char_boxes(got) == oracle_char_boxes(exp)


def test_exact_ties_follow_creation_order(tmp_path):
    """All scores equal at every step: survivors are decided by creation order alone."""
    W = 48
    img = np.full((1, 64, W), 0.25, np.float32)
    for width in (3, 10, 100):
        got, exp = _run_both(tmp_path, img, [_full_line(W)], width)
        assert text_of(got) == oracle_text_of(exp)
        assert char_boxes(got) == oracle_char_boxes(exp)


def test_peaked_scores_equal_greedy(tmp_path):
    rng = np.random.default_rng(5)
    W = 200
    labels = [0, 5, 5, 0, 5, 7, 7, 7, 0, 0, 9, 1, 1, 0, 1] + [0] * 35
    img = np.zeros((1, 64, W), np.float32)
    for t, l in enumerate(labels):
        img[0, :, 4 * t:4 * t + 4] = rng.uniform(0.0, 0.3, (64, 1))
        img[0, l, 4 * t:4 * t + 4] = 30.0
    got, exp = _run_both(tmp_path, img, [_full_line(W)], 10)
    want = ALPHABET[4] + ALPHABET[4] + ALPHABET[6] + ALPHABET[8] + ALPHABET[0] + ALPHABET[0]
    assert text_of(got) == oracle_text_of(exp) == [want]
    assert char_boxes(got) == oracle_char_boxes(exp)


def test_beam_sums_paths_greedy_does_not(tmp_path):
    """Two steps where blank wins each argmax but label 3 carries more total mass:
    p(blank,blank) < p(3,blank)+p(blank,3)+p(3,3).  Greedy reads nothing, beam reads one char."""
    W = 200
    img = np.full((1, 64, W), -40.0, np.float32)
    img[0, 0, :] = 0.0
    img[0, 0, :8] = np.log(0.55)
    img[0, 3, :8] = np.log(0.45)
    got, exp = _run_both(tmp_path, img, [_full_line(W)], 8)
    assert text_of(got) == oracle_text_of(exp) == [ALPHABET[2]]
    _, rec = fake_paths(tmp_path)
    greedy = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec, alphabet=ALPHABET))
    ginp = greedy.prepare_input(ob.ImageSource.from_tensor(img, ob.DimOrder.Chw))
    assert text_of(greedy.recognize_text(ginp, [_full_line(W)])) == [None]  # no chars -> None (recognition.rs:305-309)


def test_several_lines_and_allowed_chars(tmp_path):
    rng = np.random.default_rng(11)
    W = 400
    img = (rng.uniform(0.0, 1.0, (1, 64, W)) * 4.0).astype(np.float32)
    lines = [_full_line(100, x0=0.0), _full_line(60, x0=120.0), _full_line(200, x0=200.0)]
    allowed = ALPHABET[:20]
    got, exp = _run_both(tmp_path, img, lines, 16, allowed=allowed)
    assert text_of(got) == oracle_text_of(exp)
    assert char_boxes(got) == oracle_char_boxes(exp)
    assert all(ch in allowed for t in text_of(got) for ch in t)


def test_trained_model_page_lines():
    """Real CRNN log-probs: beam (width 10) on a synthetic page's lines matches the oracle."""
    det, rec = model_paths()
    page, _ = make_page(5, 256, 512, n_rows=3)
    eng = ob.OcrEngine(ob.OcrEngineParams(detection_model=det, recognition_model=rec,
                                          decode_method=ob.DecodeMethod.BeamSearch, beam_width=10))
    inp = eng.prepare_input(ob.ImageSource.from_tensor(page, ob.DimOrder.Hwc))
    lines = eng.find_text_lines(inp, eng.detect_words(inp))
    got = eng.recognize_text(inp, lines)
    ora = OEngine(OParams(recognition_model=OnnxModel(rec), decode_method="beam", beam_width=10))
    exp = ora.recognize_text(ora.prepare_input(page, "hwc"), [[ORect.from_raw(*w.raw()) for w in l] for l in lines])
    assert len(lines) >= 2
    assert text_of(got) == oracle_text_of(exp)
    assert char_boxes(got) == oracle_char_boxes(exp)


def test_beam_width_limits(tmp_path):
    _, rec = fake_paths(tmp_path)
    for bad in (0, 1025):
        with pytest.raises(ob.OcrsError, match="beam width"):
            ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec, alphabet=ALPHABET,
                                            decode_method=ob.DecodeMethod.BeamSearch, beam_width=bad))
