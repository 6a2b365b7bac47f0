"""BASELINE.json configs 2 (detect-only, 16 pages) and 4 (recognition-only, 4096 line regions) at
full size.  The oracle checks a sample; size-independent properties cover the rest:
batched == per-page, batch-composition invariance, duplicate lines read identically,
probabilities sum to one."""
import numpy as np
import pytest

import ocrs_b200 as ob
from oracle.engine import OcrEngine as OEngine, OcrEngineParams as OParams, find_connected_component_rects
from oracle.geometry import RotatedRect as ORect
from oracle.imageops import threshold_mask
from oracle.onnx_eval import OnnxModel
from tests.gpu_util import char_boxes, model_paths, oracle_char_boxes, oracle_text_of, raw32, text_of
from tools.synth import make_line_batch, make_page

pytestmark = pytest.mark.gpu

TOL = 1e-3  # north_star tolerance on probabilities / log-probs


def test_config2_detect_only_16_pages():
    det, _ = model_paths()
    eng = ob.OcrEngine(ob.OcrEngineParams(detection_model=det))
    ora = OEngine(OParams(detection_model=OnnxModel(det)))
    pages = [make_page(seed)[0] for seed in range(100, 116)]
    inputs = [eng.prepare_input(ob.ImageSource.from_tensor(p, ob.DimOrder.Hwc)) for p in pages]
    batched = eng.detect_words_batch(inputs)
    assert len(batched) == 16 and all(len(w) > 20 for w in batched)
    for i, inp in enumerate(inputs):  # batching changes nothing
        assert raw32(eng.detect_words(inp)) == raw32(batched[i])
    for i in (0, 7, 15):  # oracle on a sample
        oimg = ora.prepare_input(pages[i], "hwc")
        assert np.array_equal(inputs[i].image(), oimg)
        prob, oprob = eng.detect_text_pixels(inputs[i]), ora.detect_text_pixels(oimg)
        assert np.abs(prob - oprob).max() < TOL
        diff = threshold_mask(prob) != threshold_mask(oprob)
        assert np.all(np.abs(oprob[diff] - np.float32(0.2)) <= TOL)
        owords = find_connected_component_rects(threshold_mask(prob), 3.0, 100.0)
        assert raw32(batched[i]) == [tuple(np.float32(v) for v in r.raw()) for r in owords]


def test_config4_recognition_net_4096_lines():
    _, rec = model_paths()
    model, omodel = ob.Model(rec), OnnxModel(rec)
    x = make_line_batch(300, 64)
    x = np.ascontiguousarray(np.tile(x, (64, 1, 1, 1)))          # [4096,1,64,400], 64 distinct lines
    logp = np.concatenate([model.run(x[i:i + 1024]) for i in range(0, 4096, 1024)], axis=1)
    assert logp.shape == (100, 4096, 97)
    assert np.allclose(np.exp(logp.astype(np.float64)).sum(-1), 1.0, atol=1e-4)
    # oracle on the distinct lines
    exp = omodel.run(x[:64])
    assert np.abs(logp[:, :64] - exp).max() < TOL
    # batch-composition invariance: a line's labels do not depend on where it sits in a batch
    lab = logp.argmax(-1)
    for k in range(1, 64):
        assert np.array_equal(lab[:, k * 64:(k + 1) * 64], lab[:, :64])
    assert np.abs(logp.reshape(100, 64, 64, 97) - logp[:, None, :64]).max() < 1e-4
    alone = model.run(x[5:6])
    assert np.abs(alone[:, 0] - logp[:, 5]).max() < 1e-4


def test_config4_crop_path_4096_regions():
    """A tall synthetic page of 4096 32x200 line regions (-> 64x400, T=100) through
    recognize_text: crop + resize + CRNN + CTC + box mapping."""
    _, rec = model_paths()
    lines64 = make_line_batch(301, 64, height=32, width=200)[:, 0] + np.float32(0.5)   # [64,32,200] in [0,1]
    page = np.ascontiguousarray(np.tile(lines64, (64, 1, 1)).reshape(1, 4096 * 32, 200))
    eng = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec))
    inp = eng.prepare_input(ob.ImageSource.from_tensor(page, ob.DimOrder.Chw))
    rects = [[ob.RotatedRect(100.0, 32.0 * i + 16.0, 0.0, 1.0, 200.0, 32.0)] for i in range(4096)]
    got = eng.recognize_text(inp, rects)
    assert len(got) == 4096
    texts, boxes = text_of(got), char_boxes(got)
    assert sum(1 for t in texts[:64] if t) >= 60
    for k in range(1, 64):  # duplicates read identically, boxes shifted by the region offset
        assert texts[k * 64:(k + 1) * 64] == texts[:64]
        for j in (0, 31, 63):
            b0, bk = boxes[j], boxes[k * 64 + j]
            off = 32 * 64 * k
            assert (b0 is None and bk is None) or [(t + off, l, b + off, r) for (t, l, b, r) in b0] == bk
    # oracle on a sample of regions
    ora = OEngine(OParams(recognition_model=OnnxModel(rec)))
    oimg = ora.prepare_input(page, "chw")
    sample = [0, 1, 63, 64 * 17 + 5, 4095]
    exp = ora.recognize_text(oimg, [[ORect.from_raw(*rects[i][0].raw())] for i in sample])
    assert [texts[i] for i in sample] == oracle_text_of(exp)
    assert [boxes[i] for i in sample] == oracle_char_boxes(exp)
