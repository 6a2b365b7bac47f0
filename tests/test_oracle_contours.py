"""Cross-checks the oracle's Suzuki-Abe restatement against OpenCV's implementation of the same
paper (RETR_EXTERNAL + CHAIN_APPROX_NONE), including point order, on random and adversarial masks
(nested components inside holes, 1-px walls, components touching the frame)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle.contours import find_contours_external


def _cv_contours(mask):
    cs, _ = cv2.findContours(mask.astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    out = [c.reshape(-1, 2).astype(np.int32) for c in cs]
    # OpenCV returns contours last-found-first
    out.sort(key=lambda c: (c[:, 1].min(), c[c[:, 1] == c[:, 1].min(), 0].min()))
    return out


def _masks():
    rng = np.random.default_rng(7)
    yield "empty", np.zeros((16, 16), bool)
    yield "full", np.ones((9, 13), bool)
    yield "single", np.pad(np.ones((1, 1), bool), 3)
    for p in (0.1, 0.3, 0.5, 0.7):
        yield f"noise{p}", rng.random((40, 56)) < p
    m = np.zeros((40, 40), bool)
    m[2:38, 2:38] = True
    m[4:36, 4:36] = False          # ring with 2-px wall
    m[10:20, 10:20] = True         # nested component in the hole
    m[12:18, 12:18] = False
    m[14:16, 14:16] = True         # nested twice
    yield "nested_thick", m
    m = np.zeros((30, 30), bool)
    m[5, 5:25] = m[24, 5:25] = True
    m[5:25, 5] = m[5:25, 24] = True  # 1-px ring
    m[10:14, 10:14] = True           # nested
    m[12, 26:29] = True              # outside, to the right of the ring
    yield "nested_thin", m
    # blobs like a text mask
    m = np.zeros((96, 128), bool)
    for _ in range(25):
        y, x = rng.integers(0, 90), rng.integers(0, 120)
        m[y:y + rng.integers(2, 9), x:x + rng.integers(3, 30)] = True
    yield "blobs", m
    yield "diag", np.eye(12, dtype=bool)


@pytest.mark.parametrize("name,mask", list(_masks()), ids=lambda v: v if isinstance(v, str) else "")
def test_contours_match_opencv(name, mask):
    ours = find_contours_external(mask)
    ref = _cv_contours(mask)
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        assert a.shape == b.shape and (a == b).all(), name
