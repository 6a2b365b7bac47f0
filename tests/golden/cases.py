"""Seeded inputs of the golden vectors (tests/golden/vectors.json).  Shared by the generator
(make_goldens.py), the CPU check (oracle == vectors, host library == vectors) and the GPU check
(CUDA engine == vectors).  Everything here is exact arithmetic territory: no network is involved,
so the vectors do not depend on torch or on the machine."""
import numpy as np

from tests.fakes import gen_rect_grid


def f32_bits(values):
    """float32 values -> list of uint32 bit patterns (JSON-safe, exact)."""
    return [int(v) for v in np.asarray(values, dtype=np.float32).reshape(-1).view(np.uint32)]


def preprocess_cases():
    out = []
    rng = np.random.default_rng(7)
    out.append(("u8_hwc_rgb_64x100", rng.integers(0, 256, (64, 100, 3), dtype=np.uint8), "hwc"))
    rng = np.random.default_rng(8)
    out.append(("f32_chw_grey_7x13", rng.random((1, 7, 13), dtype=np.float32), "chw"))
    rng = np.random.default_rng(9)
    out.append(("u8_hwc_rgba_768x1024", rng.integers(0, 256, (768, 1024, 4), dtype=np.uint8), "hwc"))
    rng = np.random.default_rng(10)
    out.append(("f32_hwc_rgb_33x17", rng.random((33, 17, 3), dtype=np.float32), "hwc"))
    return out


def blob_mask(seed, h=192, w=256, n=30):
    """Same generator as tests/test_gpu_detection.py::_blobs (rectangles, rings, components in holes)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), bool)
    for _ in range(n):
        y, x = rng.integers(0, h - 4), rng.integers(0, w - 8)
        hh, ww = rng.integers(3, 24), rng.integers(6, 90)
        m[y:y + hh, x:x + ww] = True
    for _ in range(n // 4):
        y, x = rng.integers(0, h - 30), rng.integers(0, w - 60)
        m[y:y + 28, x:x + 56] = True
        m[y + 3:y + 25, x + 3:x + 53] = False
        m[y + 9:y + 19, x + 14:x + 42] = True
    return m


def mask_cases():
    grid = np.zeros((400, 400), bool)
    for (t, l, b, r) in gen_rect_grid((10, 10), (5, 5), (10, 50), (10, 5)):  # detection.rs:212-246
        grid[t:b + 1, l:r + 1] = True
    return [("reference_grid_400x400", grid), ("blobs_seed0", blob_mask(0)), ("blobs_seed1", blob_mask(1))]


def layout_cases():
    """Word rects as raw (cx, cy, ux, uy, w, h) tuples."""
    out = []
    words = []
    for (t, l, b, r) in gen_rect_grid((10, 10), (4, 6), (12, 40), (14, 8)):
        words.append(((l + r) / 2.0, (t + b) / 2.0, 0.0, 1.0, float(r - l), float(b - t)))
    out.append(("grid_4x6", words))
    rng = np.random.default_rng(21)
    words = []
    for row in range(7):
        x = float(rng.uniform(5, 40))
        for _ in range(int(rng.integers(2, 9))):
            w, h = float(rng.uniform(20, 80)), float(rng.uniform(12, 22))
            if 250 < x < 300 and row % 2 == 0:
                x = 320.0  # a gutter on even rows
            words.append((x + w / 2, 30.0 + 34.0 * row + float(rng.uniform(-2, 2)), 0.0, 1.0, w, h))
            x += w + float(rng.uniform(4, 14))
    order = rng.permutation(len(words))
    out.append(("ragged_rows_shuffled", [words[i] for i in order]))
    return out


def crop_cases():
    """(name, page seed, word rects of one line): the cases of tests/test_gpu_recognition.py::test_line_crop_bit_exact."""
    out = []
    for seed in range(6):
        rng = np.random.default_rng(seed)
        x0 = float(rng.choice([-12.0, 5.0, 200.0, 520.0]))
        y = float(rng.choice([4.0, 100.0, 380.0]))
        n = int(rng.integers(1, 7))
        rotated = bool(seed % 2)
        words = []
        x = x0
        for _ in range(n):
            w = float(rng.uniform(20, 90))
            h = float(rng.uniform(14, 26))
            if rotated:
                a = float(rng.uniform(-0.2, 0.2))
                ux, uy = float(np.sin(a)), float(np.cos(a))
            else:
                ux, uy = 0.0, 1.0
            words.append((np.float32(x + w / 2), np.float32(y + rng.uniform(-1.5, 1.5)), ux, uy, np.float32(w), np.float32(h)))
            x += w + float(rng.uniform(3, 12))
        out.append((f"line_seed{seed}", words))
    return out


CTC_LABELS = [0, 5, 5, 0, 5, 7, 7, 7, 0, 0, 9, 1, 1, 0, 1] + [0] * 35   # one label per 4-px step (T = 50)


def ctc_image():
    """The controlled-score image of tests/test_gpu_recognition.py::test_ctc_greedy_through_fake_model."""
    rng = np.random.default_rng(5)
    W = 200
    img = np.zeros((1, 64, W), np.float32)
    for t, l in enumerate(CTC_LABELS):
        img[0, :, 4 * t:4 * t + 4] = rng.uniform(0.0, 0.3, (64, 1))
        img[0, l, 4 * t:4 * t + 4] = 0.9
    return img


def text_item_cases():
    def chars(text, width, top=0, height=25):
        return [(ch, (top, i * width, top + height, (i + 1) * width)) for i, ch in enumerate(text)]
    rng = np.random.default_rng(4)
    ragged, x = [], 3
    for ch in "ragged line":
        w, t, h = int(rng.integers(4, 15)), int(rng.integers(0, 9)), int(rng.integers(8, 30))
        ragged.append((ch, (t, x, t + h, x + w)))
        x += w + int(rng.integers(0, 4))
    return [("line_one", chars("line one", 10)), ("foo_bar_baz", chars("foo bar  baz ", 10)), ("ragged", ragged)]
