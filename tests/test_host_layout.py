"""Host-side layout analysis of the product (C++ through the C ABI; no GPU needed) against the
oracle: bit-identical rect lists, line by line, on the reference's fixtures and on random pages."""
import numpy as np
import pytest

import ocrs_b200 as ob
from oracle.geometry import F, PointF, Rect, RotatedRect as ORect, Vec2
from oracle.layout import find_text_lines as oracle_find_text_lines
from tests.fakes import gen_rect_grid


def _to_product(rects):
    return [ob.RotatedRect(*r.raw()) for r in rects]


def _raw(lines):
    return [[tuple(np.float32(v) for v in r.raw()) for r in line] for line in lines]


def _check(words):
    got = ob.find_text_lines(_to_product(words))
    exp = oracle_find_text_lines(words)
    assert _raw(got) == _raw(exp)
    return got


def _axis(tlbr):
    return ORect.from_rect(Rect(*tlbr).to_f32())


@pytest.mark.parametrize("seed", [1234, 5, 6])
def test_two_column_fixture(seed):
    """layout_analysis.rs:294-350"""
    left = gen_rect_grid((0, 0), (10, 5), (5, 5), (3, 2))
    right = gen_rect_grid((0, 33 + 20), (10, 5), (5, 5), (3, 2))
    words = [_axis(r) for r in left + right]
    np.random.default_rng(seed).shuffle(words)
    lines = _check(words)
    assert len(lines) == 20 and all(len(l) == 5 for l in lines)


def test_overlapping_words_fixture():
    """layout_analysis.rs:256-264"""
    _check([_axis(r) for r in gen_rect_grid((0, 0), (2, 2), (10, 20), (50, -5))])


def test_empty_and_single():
    assert ob.find_text_lines([]) == []
    _check([_axis((10, 10, 20, 60))])


def _random_page(rng, n_rows, rotated):
    words = []
    y = 20.0
    two_col = rng.random() < 0.5
    for _ in range(n_rows):
        h = float(rng.uniform(12, 28))
        spans = [(20, 480), (540, 1000)] if two_col else [(20, 1000)]
        for (x0, x1) in spans:
            x = x0 + float(rng.uniform(0, 20))
            while x < x1 - 30:
                w = float(rng.uniform(20, 120))
                if x + w > x1:
                    break
                cy = y + h / 2 + float(rng.uniform(-2, 2))
                if rotated:
                    ang = float(rng.uniform(-0.15, 0.15))
                    up = Vec2(np.sin(ang), np.cos(ang))
                else:
                    up = Vec2(0.0, 1.0)
                words.append(ORect(PointF(F(x + w / 2), F(cy)), up, F(w + 6), F(h + 6)))
                x += w + float(rng.uniform(4, 14))
        y += h + float(rng.uniform(6, 18))
    order = rng.permutation(len(words))
    return [words[i] for i in order]


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("rotated", [False, True])
def test_random_pages_bit_exact(seed, rotated):
    rng = np.random.default_rng(100 + seed)
    words = _random_page(rng, n_rows=int(rng.integers(3, 14)), rotated=rotated)
    _check(words)


def test_find_text_lines_survives_degenerate_rects():
    """NaN / inf / zero-sized / zero-axis / far-away rects: the host layout code returns (or reports an
    error) without crashing or looping; every input rect appears at most once in the output."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = int(rng.integers(0, 60))
        rects = []
        for _ in range(n):
            v = list(rng.normal(0, 300, 2)) + list(rng.normal(0, 1, 2)) + list(np.abs(rng.normal(30, 40, 2)))
            mode = int(rng.integers(0, 12))
            if mode == 0:
                v[int(rng.integers(0, 6))] = float("nan")
            elif mode == 1:
                v[int(rng.integers(0, 6))] = float("inf")
            elif mode == 2:
                v[4] = 0.0
            elif mode == 3:
                v[5] = -5.0
            elif mode == 4:
                v[2] = v[3] = 0.0
            elif mode == 5:
                v[0] = 1e30
            rects.append(ob.RotatedRect(*[float(x) for x in v]))
        try:
            lines = ob.find_text_lines(rects)
        except ob.OcrsError:
            continue
        assert sum(len(l) for l in lines) <= n
