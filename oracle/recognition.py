"""Recognition pre/post-processing and CTC decoding (oracle; test infrastructure).

Restates ocrs/src/recognition.rs (line polygons, width bucketing, polygon crop, batch packing,
CTC step -> char box mapping) and `rten::ctc::CtcDecoder::{decode_greedy, decode_beam}`
(rten 0.24.0, Cargo.lock:682, not vendored: greedy/beam restated from the CTC definition --
Graves et al. 2006; blank = class 0, see SURVEY App. A.6 -- parity with rten unpinned beyond
lib.rs:526-577).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import BLACK_VALUE
from .geometry import (
    F, Line, Point, PointF, Rect, RotatedRect, as_i32, as_u32, bounding_rect_of_rotated, polygon_bounding_rect,
    polygon_edges, polygon_fill_rows, round_f32,
)
from .imageops import resize_bilinear
from .layout import downwards_line, leftmost_edge, rightmost_edge


def line_polygon(words: Sequence[RotatedRect]) -> List[Point]:
    """recognition.rs:29-55.  `p.y as i32` truncates toward zero (:32)."""

    def trunc_point(p: PointF) -> Point:
        return Point.from_yx(as_i32(p.y), as_i32(p.x))

    poly: List[Point] = []
    for w in words:
        left = downwards_line(leftmost_edge(w))
        right = downwards_line(rightmost_edge(w))
        poly.append(trunc_point(left.start))
        poly.append(trunc_point(right.start))
    for w in reversed(words):
        left = downwards_line(leftmost_edge(w))
        right = downwards_line(rightmost_edge(w))
        poly.append(trunc_point(right.end))
        poly.append(trunc_point(left.end))
    return poly


def resized_line_width(orig_width: int, orig_height: int, height: int) -> int:
    """recognition.rs:58-75"""
    aspect_ratio = F(orig_width) / F(orig_height)
    v = F(height) * aspect_ratio
    # f32::clamp(10., 2400.) -- NaN stays NaN -> `as u32` = 0
    if not math.isnan(float(v)):
        v = min(max(v, F(10.0)), F(2400.0))
    return as_u32(v)


def next_multiple_of(v: int, m: int) -> int:
    return ((v + m - 1) // m) * m


@dataclass
class TextRecLine:
    """recognition.rs:79-89"""
    index: int
    region: List[Point]
    resized_width: int


def prepare_text_line(image: np.ndarray, line_region: Sequence[Point], resized_width: int,
                      output_height: int) -> np.ndarray:
    """recognition.rs:91-126.  image: f32 [1, H, W].  Returns f32 [output_height, resized_width]."""
    _, img_h, img_w = image.shape
    page_index_rect = Rect.from_hw(img_h, img_w).adjust_tlbr(0, 0, -1, -1)
    grey = image[0]
    line_rect = polygon_bounding_rect(line_region)
    lh, lw = max(line_rect.height(), 0), max(line_rect.width(), 0)
    line_img = np.full((lh, lw), F(BLACK_VALUE), dtype=np.float32)
    for y, xs in polygon_fill_rows(line_region):
        oy = y - line_rect.top
        # both the page pixel and the canvas pixel are tested against the *page* rect (:112)
        if not (page_index_rect.top <= y <= page_index_rect.bottom and page_index_rect.top <= oy <= page_index_rect.bottom):
            continue
        ox = xs - line_rect.left
        ok = (xs >= page_index_rect.left) & (xs <= page_index_rect.right) & (ox >= page_index_rect.left) & (ox <= page_index_rect.right)
        line_img[oy, ox[ok]] = grey[y, xs[ok]]
    return resize_bilinear(line_img, output_height, resized_width)


def prepare_text_line_batch(image: np.ndarray, lines: Sequence[TextRecLine], output_height: int,
                            output_width: int) -> np.ndarray:
    """recognition.rs:135-158 -> f32 [B, 1, H, output_width], right-padded with BLACK_VALUE."""
    out = np.full((len(lines), 1, output_height, output_width), F(BLACK_VALUE), dtype=np.float32)
    for i, line in enumerate(lines):
        out[i, 0, :, : line.resized_width] = prepare_text_line(image, line.region, line.resized_width, output_height)
    return out


def polygon_slice_bounding_rect(poly: Sequence[Point], min_x: int, max_x: int) -> Optional[Rect]:
    """recognition.rs:162-193"""
    out: Optional[Rect] = None
    for e in polygon_edges(poly):
        e = e.rightwards()
        if (e.start.x < min_x and e.end.x < min_x) or (e.start.x > max_x and e.end.x > max_x):
            continue
        ef = e.to_f32()
        y0 = ef.y_for_x(F(min_x))
        start = e.start if y0 is None else Point.from_yx(as_i32(round_f32(y0)), min_x)
        y1 = ef.y_for_x(F(max_x))
        end = e.end if y1 is None else Point.from_yx(as_i32(round_f32(y1)), max_x)
        br = Line(start, end).bounding_rect()
        out = br if out is None else out.union(br)
    return out


# ---- CTC ------------------------------------------------------------------------------------
@dataclass
class DecodeStep:
    label: int
    pos: int


def filter_excluded_char_labels(seq: np.ndarray, excluded: Optional[Sequence[int]]) -> np.ndarray:
    """recognition.rs:547-561 (in place on a copy)."""
    if excluded:
        seq = seq.copy()
        seq[:, list(excluded)] = -np.inf
    return seq


def ctc_decode_greedy(seq: np.ndarray) -> Tuple[List[DecodeStep], float]:
    """`CtcDecoder::decode_greedy` (call recognition.rs:511).  seq: [T, C] log-probs.
    Per-step argmax (first index on ties), collapse repeats, drop blank 0; a blank between two
    equal labels re-arms emission; `pos` = first timestep of the run."""
    labels = np.argmax(seq, axis=1)  # first maximum
    steps: List[DecodeStep] = []
    score = 0.0
    last = 0
    for pos, l in enumerate(labels.tolist()):
        score += float(seq[pos, l])
        if l == last:
            continue
        last = l
        if l > 0:
            steps.append(DecodeStep(label=int(l), pos=pos))
    return steps, score


def _logaddexp(a: float, b: float) -> float:
    if a == -math.inf:
        return b
    if b == -math.inf:
        return a
    m = max(a, b)
    return m + math.log1p(math.exp(-abs(a - b)))


def ctc_decode_beam(seq: np.ndarray, width: int) -> Tuple[List[DecodeStep], float]:
    """`CtcDecoder::decode_beam` (call recognition.rs:512-514): CTC prefix beam search
    (Hannun et al. 2014) in log space, no language model.  Each prefix keeps
    (log p ending in blank, log p ending in non-blank); after every timestep the `width` most
    probable prefixes survive (ties: earlier-created prefix first).  `pos` of an emitted label
    is the timestep at which that prefix first entered the beam (it is kept while the prefix
    survives, like the greedy decoder's "first timestep of the run")."""
    T, C = seq.shape
    NEG = -math.inf
    # prefix (tuple of labels) -> [p_blank, p_non_blank, positions tuple, creation order]
    beams: Dict[tuple, list] = {(): [0.0, NEG, (), 0]}
    order = 1
    for t in range(T):
        nxt: Dict[tuple, list] = {}

        def get(prefix, positions):
            nonlocal order
            e = nxt.get(prefix)
            if e is None:
                if prefix in beams:  # re-reached through its parent: keep the original timesteps
                    positions = beams[prefix][2]
                e = [NEG, NEG, positions, order]
                order += 1
                nxt[prefix] = e
            return e

        for prefix, (pb, pnb, positions, _) in beams.items():
            total = _logaddexp(pb, pnb)
            for c in range(C):
                p = float(seq[t, c])
                if p == NEG:
                    continue
                if c == 0:
                    e = get(prefix, positions)
                    e[0] = _logaddexp(e[0], total + p)
                    continue
                end = prefix[-1] if prefix else None
                new_prefix = prefix + (c,)
                if c == end:
                    # repeat of last label: extends the same prefix via the non-blank path,
                    # creates a new label only after a blank
                    e = get(prefix, positions)
                    e[1] = _logaddexp(e[1], pnb + p)
                    e2 = get(new_prefix, positions + (t,))
                    e2[1] = _logaddexp(e2[1], pb + p)
                else:
                    e2 = get(new_prefix, positions + (t,))
                    e2[1] = _logaddexp(e2[1], total + p)
        ranked = sorted(nxt.items(), key=lambda kv: (-_logaddexp(kv[1][0], kv[1][1]), kv[1][3]))
        beams = dict(ranked[:width])
    best_prefix, (pb, pnb, positions, _) = next(iter(beams.items()))
    steps = [DecodeStep(label=int(l), pos=int(p)) for l, p in zip(best_prefix, positions)]
    return steps, _logaddexp(pb, pnb)


# ---- results --------------------------------------------------------------------------------
@dataclass
class TextChar:
    """text_items.rs:47-53"""
    char: str
    rect: Rect


@dataclass
class LineRecResult:
    """recognition.rs:221-234"""
    line: TextRecLine
    rec_input_len: int
    ctc_input_len: int
    steps: List[DecodeStep]


def text_lines_from_recognition_results(results: Sequence[LineRecResult], alphabet: str) -> List[Optional[List[TextChar]]]:
    """recognition.rs:241-311"""
    out: List[Optional[List[TextChar]]] = []
    for result in results:
        line_rect = polygon_bounding_rect(result.line.region)
        x_scale_factor = F(line_rect.width()) / F(result.line.resized_width)
        downsample_factor = as_u32(round_f32(F(result.rec_input_len) / F(result.ctc_input_len)))
        steps = result.steps
        chars: List[TextChar] = []
        for i, step in enumerate(steps):
            start_x = step.pos * downsample_factor
            end_x = steps[i + 1].pos * downsample_factor if i + 1 < len(steps) else result.line.resized_width
            start_x, end_x = (line_rect.left + as_i32(F(x) * x_scale_factor) for x in (start_x, end_x))
            if start_x >= line_rect.right:
                continue
            idx = step.label - 1
            ch = alphabet[idx] if 0 <= idx < len(alphabet) else "?"
            rect = polygon_slice_bounding_rect(result.line.region, start_x, end_x)
            assert rect is not None, "invalid X coords"  # recognition.rs:299
            chars.append(TextChar(ch, rect))
        out.append(chars if chars else None)
    return out
