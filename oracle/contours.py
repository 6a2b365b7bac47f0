"""Outer-contour extraction (oracle; test infrastructure).

Restates `rten_imageproc::find_contours(mask, RetrievalMode::External)` (call site
detection.rs:46) from the published algorithm: S. Suzuki, K. Abe, "Topological structural
analysis of digitized binary images by border following", CVGIP 30(1), 1985 -- Algorithm 1 with
the Appendix-II restriction to outermost borders (start only where LNBD <= 0, mark with 2/-2).
rten-imageproc 0.24.0 (Cargo.lock:726) is not vendored: parity with it is pinned only through
the reference goldens that consume contours (detection.rs:212-246, lib.rs:437-445), i.e. for
contour *extents*; point order follows the paper (start = first pixel in raster order, walk
counter-clockwise in image coordinates, every border-pixel visit is emitted).

As an independent cross-check `tests/test_oracle_contours.py` compares this routine with
OpenCV's implementation of the same paper (cv2.findContours RETR_EXTERNAL, CHAIN_APPROX_NONE).
"""
from __future__ import annotations

from typing import List

import numpy as np

# 8-neighbourhood in CLOCKWISE order (image coordinates, y down): E, SE, S, SW, W, NW, N, NE
_NBR = [(0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1)]
_NBR_INDEX = {d: i for i, d in enumerate(_NBR)}


def find_contours_external(mask: np.ndarray) -> List[np.ndarray]:
    """mask: bool/uint8 [H, W].  Returns a list of int32 arrays [n_points, 2] holding (x, y) pixel
    coordinates, in raster discovery order of each component's first pixel."""
    h, w = mask.shape
    f = np.zeros((h + 2, w + 2), dtype=np.int32)
    f[1:-1, 1:-1] = mask.astype(bool)
    out: List[np.ndarray] = []

    # Candidate outer-border start points never change (labelling keeps non-zero pixels
    # non-zero): pixel != 0 and left neighbour == 0.  We still visit them in raster order and
    # re-evaluate the paper's conditions on the *current* label image.
    nz = f != 0
    cand = nz[:, 1:] & ~nz[:, :-1]
    ys, xs = np.nonzero(cand)
    xs = xs + 1
    fl = f  # alias

    # LNBD needs "the last pixel on this row with a value other than 0 or 1".  Rather than
    # scanning every pixel we look it up when needed from the row contents.
    for y, x in zip(ys.tolist(), xs.tolist()):
        if fl[y, x] != 1:  # already labelled by an earlier border
            continue
        # LNBD: value of the nearest pixel to the left on this row whose value is not 0 or 1.
        row = fl[y, :x]
        idx = np.nonzero((row != 0) & (row != 1))[0]
        lnbd = int(row[idx[-1]]) if idx.size else 0
        if lnbd > 0:
            continue  # inside a hole of a traced component: not an outermost border
        pts = _follow_border(fl, y, x)
        out.append(np.asarray(pts, dtype=np.int32) - 1)  # remove the 1-px padding
    return out


def _follow_border(f: np.ndarray, i: int, j: int):
    """Steps (3.1)-(3.5) of Algorithm 1 with NBD fixed to 2 (Appendix II)."""
    nbd = 2
    pts = []
    # (3.1) clockwise from (i, j-1) around (i, j): first non-zero pixel
    start_dir = _NBR_INDEX[(0, -1)]
    i1 = j1 = None
    for k in range(8):
        dy, dx = _NBR[(start_dir + k) % 8]
        if f[i + dy, j + dx] != 0:
            i1, j1 = i + dy, j + dx
            break
    if i1 is None:
        f[i, j] = -nbd
        return [(j, i)]
    i2, j2 = i1, j1
    i3, j3 = i, j
    while True:
        # (3.3) counter-clockwise from the element after (i2, j2) around (i3, j3)
        d0 = _NBR_INDEX[(i2 - i3, j2 - j3)]
        examined_east_zero = False
        i4 = j4 = None
        for k in range(1, 9):
            dy, dx = _NBR[(d0 - k) % 8]
            v = f[i3 + dy, j3 + dx]
            if v != 0:
                i4, j4 = i3 + dy, j3 + dx
                break
            if (dy, dx) == (0, 1):
                examined_east_zero = True
        # (3.4)
        if examined_east_zero:
            f[i3, j3] = -nbd
        elif f[i3, j3] == 1:
            f[i3, j3] = nbd
        pts.append((j3, i3))
        # (3.5)
        if (i4, j4) == (i, j) and (i3, j3) == (i1, j1):
            break
        i2, j2 = i3, j3
        i3, j3 = i4, j4
    return pts
