"""Detection post-processing on the GPU (threshold -> connected components -> outer contours ->
RDP -> min-area rect) vs the oracle, bit-exact, by injecting masks through the public API: the
fake detection model (x + 0.5, lib.rs:339-362) at the image's own size makes prob == image."""
import numpy as np
import pytest

import ocrs_b200 as ob
from oracle.engine import find_connected_component_rects
from tests.fakes import gen_rect_grid
from tests.gpu_util import fake_paths, raw32

pytestmark = pytest.mark.gpu


def _engine_for(tmp_path, hw):
    det, _ = fake_paths(tmp_path, hw)
    return ob.OcrEngine(ob.OcrEngineParams(detection_model=det))


def _inject(engine, mask):
    img = mask.astype(np.float32)[None]  # 1 -> prob 1.0, 0 -> prob 0.0
    return engine.prepare_input(ob.ImageSource.from_tensor(img, ob.DimOrder.Chw))


def _check(tmp_path, mask):
    engine = _engine_for(tmp_path, mask.shape)
    inp = _inject(engine, mask)
    prob = engine.detect_text_pixels(inp)
    assert np.array_equal(prob > np.float32(0.2), mask)
    got = engine.detect_words(inp)
    exp = find_connected_component_rects(mask, 3.0, 100.0)
    assert raw32(got) == [tuple(np.float32(v) for v in r.raw()) for r in exp]
    return got


def test_reference_grid_fixture(tmp_path):
    """detection.rs:212-246 (with the engine's expand_dist = 3)"""
    mask = np.zeros((400, 400), bool)
    for (t, l, b, r) in gen_rect_grid((10, 10), (5, 5), (10, 50), (10, 5)):
        mask[t:b + 1, l:r + 1] = True
    rects = _check(tmp_path, mask)
    assert len(rects) == 25
    for r in rects:
        assert sorted([round(r.h), round(r.w)]) == [16, 56]


def _blobs(rng, h, w, n):
    m = np.zeros((h, w), bool)
    for _ in range(n):
        y, x = rng.integers(0, h - 4), rng.integers(0, w - 8)
        hh, ww = rng.integers(3, 24), rng.integers(6, 90)
        m[y:y + hh, x:x + ww] = True
    # carve some holes and put components inside them
    for _ in range(n // 4):
        y, x = rng.integers(0, h - 30), rng.integers(0, w - 60)
        m[y:y + 28, x:x + 56] = True
        m[y + 3:y + 25, x + 3:x + 53] = False
        m[y + 9:y + 19, x + 14:x + 42] = True
    return m


@pytest.mark.parametrize("seed", range(4))
def test_random_blob_masks(tmp_path, seed):
    rng = np.random.default_rng(seed)
    _check(tmp_path, _blobs(rng, 192, 256, 30))


def test_rotated_and_noisy_components(tmp_path):
    import cv2
    rng = np.random.default_rng(11)
    canvas = np.zeros((256, 320), np.uint8)
    for _ in range(14):
        c = (int(rng.integers(30, 290)), int(rng.integers(30, 226)))
        box = cv2.boxPoints((c, (float(rng.uniform(30, 90)), float(rng.uniform(8, 22))), float(rng.uniform(-40, 40))))
        cv2.fillPoly(canvas, [box.astype(np.int32)], 1)
    mask = canvas.astype(bool) | (rng.random(canvas.shape) < 0.01)
    _check(tmp_path, mask)


@pytest.mark.parametrize("name", ["empty", "full", "frame_touching", "thin_ring_nested", "salt"])
def test_edge_masks(tmp_path, name):
    h, w = 96, 128
    m = np.zeros((h, w), bool)
    if name == "full":
        m[:] = True
    elif name == "frame_touching":
        m[0:12, 0:40] = True
        m[h - 12:, w - 40:] = True
        m[40:60, 0:30] = True
    elif name == "thin_ring_nested":
        m[10, 10:80] = m[70, 10:80] = True
        m[10:71, 10] = m[10:71, 79] = True
        m[30:50, 25:65] = True       # inside the ring: must not be reported
        m[30:50, 90:125] = True      # outside, right of the ring
    elif name == "salt":
        m = np.random.default_rng(5).random((h, w)) < 0.3
    _check(tmp_path, m)


def test_full_size_page_mask(tmp_path):
    rng = np.random.default_rng(21)
    _check(tmp_path, _blobs(rng, 768, 1024, 300))


def test_batched_detect_matches_single(tmp_path):
    rng = np.random.default_rng(31)
    masks = [_blobs(rng, 192, 256, 20) for _ in range(3)]
    engine = _engine_for(tmp_path, masks[0].shape)
    inputs = [_inject(engine, m) for m in masks]
    batched = engine.detect_words_batch(inputs)
    for inp, b in zip(inputs, batched):
        assert raw32(engine.detect_words(inp)) == raw32(b)
