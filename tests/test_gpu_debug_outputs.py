"""Batched debug outputs (ocrs-cli --text-map / --text-mask / --text-line-images, main.rs:423-443) against the
single-page entry points and the oracle."""
import numpy as np
import pytest

import ocrs_b200 as ob
from oracle.engine import OcrEngine as OEngine, OcrEngineParams as OParams
from oracle.onnx_eval import OnnxModel
from tests.gpu_util import model_paths, to_oracle_rects
from tools.synth import make_page

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    det, rec = model_paths()
    eng = ob.OcrEngine(ob.OcrEngineParams(detection_model=det, recognition_model=rec))
    pages = [make_page(700 + i, 320 + 32 * i, 448, n_rows=5)[0] for i in range(3)]
    inputs = [eng.prepare_input(ob.ImageSource.from_tensor(p, ob.DimOrder.Hwc)) for p in pages]
    return det, rec, eng, pages, inputs


def test_text_map_and_mask_batch(setup):
    det, rec, eng, pages, inputs = setup
    maps, masks = eng.detect_text_pixels_batch(inputs)
    thr = np.float32(eng.detection_threshold())
    for inp, m, k in zip(inputs, maps, masks):
        single = eng.detect_text_pixels(inp)
        assert m.shape == single.shape and np.array_equal(m, single)      # batch-invariant: same kernels, same bits
        assert np.array_equal(k, (m > thr).astype(np.uint8))             # main.rs:431-433
    only_masks = eng.detect_text_pixels_batch(inputs, maps=False)
    assert only_masks[0] is None and all(np.array_equal(a, b) for a, b in zip(only_masks[1], masks))
    ora = OEngine(OParams(detection_model=OnnxModel(det), recognition_model=OnnxModel(rec)))
    ref = ora.detect_text_pixels(ora.prepare_input(pages[0], "hwc"))
    assert np.max(np.abs(ref - maps[0])) < 1e-3


def test_text_line_images_batch(setup):
    det, rec, eng, pages, inputs = setup
    words = eng.detect_words(inputs[1])
    lines = eng.find_text_lines(inputs[1], words)
    assert len(lines) >= 3
    imgs = eng.prepare_recognition_inputs(inputs[1], lines)
    assert len(imgs) == len(lines)
    ora = OEngine(OParams(detection_model=OnnxModel(det), recognition_model=OnnxModel(rec)))
    oimg = ora.prepare_input(pages[1], "hwc")
    for line, im in zip(lines, imgs):
        one = eng.prepare_recognition_input(inputs[1], line)
        assert im.shape == one.shape and np.array_equal(im, one)
        ref = ora.prepare_recognition_input(oimg, to_oracle_rects(line))
        assert np.array_equal(np.asarray(ref, np.float32).reshape(im.shape), im)   # exact stage
    assert eng.prepare_recognition_inputs(inputs[1], []) == []
