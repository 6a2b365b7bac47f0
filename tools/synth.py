"""Seeded synthetic pages and line crops for tests and bench.
Fixture tooling: not part of the product path."""
from __future__ import annotations

from typing import List, Tuple

import cv2
import numpy as np

VOCAB = ("the of and to in is that it was for on are as with his they at be this from have or by one had not but "
         "what all were when we there can an your which their said if do will each about how up out them then she "
         "many some so these would other into has more her two like him see time could no make than first been its "
         "who now people my made over did down only way find use may water long little very after words called just "
         "where most know get through back much before go good new write our used me man too any day same right look "
         "think also around another came come work three word must because does part even place well such here take "
         "why things help put years different away again off went old number great tell men say small every found "
         "still between name should Home big give air line set own under read last never us left end along while "
         "OCR B200 GPU 2024 1985 3.14 42 100% (a) [b] x=y+z").split()


def make_page(seed: int, height: int = 768, width: int = 1024, n_rows: int | None = None,
              two_col: bool | None = None, with_boxes: bool = False):
    """White (250 +/- 5) page with rows of dark words rendered by cv2.putText.
    Returns (HWC u8 RGB image, list of row strings); with_boxes adds a third item: per row, the
    list of word boxes (top, left, bottom, right)."""
    rng = np.random.default_rng(seed)
    img = np.clip(250 + rng.normal(0, 2.0, (height, width, 1)), 0, 255).repeat(3, axis=2).astype(np.uint8)
    if two_col is None:
        two_col = bool(seed % 2)
    cols = [(24, width // 2 - 24), (width // 2 + 24, width - 24)] if two_col else [(24, width - 24)]
    texts: List[str] = []
    all_boxes = []
    y = 30
    rows = 0
    while y < height - 20 and (n_rows is None or rows < n_rows):
        scale = float(rng.uniform(0.5, 0.85))
        thick = 1 if scale < 0.7 else 2
        line_h = int(30 * scale) + 12
        for (x0, x1) in cols:
            x = x0 + int(rng.integers(0, 12))
            words = []
            boxes = []
            while True:
                w = VOCAB[int(rng.integers(0, len(VOCAB)))]
                (tw, th), base = cv2.getTextSize(w, cv2.FONT_HERSHEY_SIMPLEX, scale, thick)
                if x + tw > x1:
                    break
                shade = int(rng.integers(10, 60))
                cv2.putText(img, w, (x, y + th), cv2.FONT_HERSHEY_SIMPLEX, scale, (shade, shade, shade), thick,
                            cv2.LINE_AA)
                words.append(w)
                boxes.append((y - 1, x - 1, y + th + base // 2 + 2, x + tw + 1))
                x += tw + int(rng.integers(12, 20))
                if rng.random() < 0.04:
                    break
            if words:
                texts.append(" ".join(words))
                all_boxes.append(boxes)
        y += line_h + int(rng.integers(4, 14))
        rows += 1
    if with_boxes:
        return img, texts, all_boxes
    return img, texts


def make_pages(seeds, height: int = 768, width: int = 1024) -> List[np.ndarray]:
    return [make_page(s, height, width)[0] for s in seeds]


def make_line_batch(seed: int, n: int, height: int = 64, width: int = 400) -> np.ndarray:
    """Config 4: pre-cropped line images as the recognition network sees them, f32 [n,1,64,400]."""
    rng = np.random.default_rng(seed)
    out = np.full((n, 1, height, width), 0.48, dtype=np.float32)
    for i in range(n):
        canvas = np.full((height, width), 250, dtype=np.uint8)
        x = 4
        while x < width - 40:
            w = VOCAB[int(rng.integers(0, len(VOCAB)))]
            (tw, th), _ = cv2.getTextSize(w, cv2.FONT_HERSHEY_SIMPLEX, 1.4, 2)
            if x + tw > width - 4:
                break
            cv2.putText(canvas, w, (x, 46), cv2.FONT_HERSHEY_SIMPLEX, 1.4, 30, 2, cv2.LINE_AA)
            x += tw + 24
        out[i, 0] = canvas.astype(np.float32) / 255.0 - 0.5
    return out
