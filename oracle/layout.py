"""Layout analysis: words -> lines in reading order (oracle; test infrastructure).

Line-by-line restatement of ocrs/src/layout_analysis.rs, layout_analysis/empty_rects.rs and
geom_util.rs.  Rust semantics that decide integer-exact results are kept:
stable sorts, first-minimum `min_by_key`, `as i32` truncation, `round` half away from zero,
and std::collections::BinaryHeap's exact sift order for equal scores.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from .geometry import (
    F, LineF, Point, PointF, Rect, RotatedRect, as_i32, bounding_rect_of_rotated, fsqrt, log2f, round_f32,
)

_F0 = F(0.0)


# ---- geom_util.rs ---------------------------------------------------------------------------
def _corners_sorted_by_x(r: RotatedRect):
    # geom_util.rs:8,15: `corners.sort_by(|a, b| a.x.total_cmp(&b.x))` (stable)
    return sorted(r.corners(), key=lambda p: float(p.x))


def rightmost_edge(r: RotatedRect) -> LineF:
    """geom_util.rs:6-10"""
    c = _corners_sorted_by_x(r)
    return LineF(c[2], c[3])


def leftmost_edge(r: RotatedRect) -> LineF:
    """geom_util.rs:13-17"""
    c = _corners_sorted_by_x(r)
    return LineF(c[0], c[1])


def downwards_line(l: LineF) -> LineF:
    """geom_util.rs:20-26"""
    return l.downwards()


# ---- layout_analysis.rs ---------------------------------------------------------------------
def rects_separated_by_line(a: RotatedRect, b: RotatedRect, l: LineF) -> bool:
    """layout_analysis.rs:8-11"""
    return LineF(a.center(), b.center()).intersects(l)


def group_into_lines(rects: Sequence[RotatedRect], separators: Sequence[LineF]) -> List[List[RotatedRect]]:
    """layout_analysis.rs:19-71"""
    sorted_rects = sorted(rects, key=lambda r: as_i32(r.bounding_rect().left))  # stable (:21)
    lines: List[List[RotatedRect]] = []
    overlap_threshold = F(5)  # :27
    max_h_overlap = F(5)  # :35
    while sorted_rects:
        line = [sorted_rects.pop(0)]
        while True:
            last = line[-1]
            last_edge = rightmost_edge(last)
            last_edge_cx = last_edge.center().x
            best_i, best_key = -1, None
            for i, r in enumerate(sorted_rects):
                edge = leftmost_edge(r)
                if not (r.cx > last.cx):
                    continue
                if not (edge.center().x - last_edge_cx >= -max_h_overlap):
                    continue
                if not (last_edge.vertical_overlap(edge) >= overlap_threshold):
                    continue
                if any(rects_separated_by_line(last, r, s) for s in separators):
                    continue
                key = as_i32(r.cx)
                if best_key is None or key < best_key:  # first minimum (:59)
                    best_i, best_key = i, key
            if best_i < 0:
                break
            line.append(sorted_rects.pop(best_i))
        lines.append(line)
    return lines


def _score(r: Rect) -> np.float32:
    """layout_analysis.rs:127-135"""
    aspect_ratio = F(r.height()) / F(r.width())
    v = abs(log2f(aspect_ratio))
    if v < F(3.0):
        weight = F(0.5)
    elif v < F(5.0):
        weight = F(1.5)
    else:
        weight = v
    return fsqrt(F(r.area()) * weight)


def find_block_separators(words: Sequence[RotatedRect]) -> List[Rect]:
    """layout_analysis.rs:83-155"""
    br = bounding_rect_of_rotated(words)
    if br is None:
        return []
    page_rect = br.integral_bounding_rect()

    lines = group_into_lines(words, [])
    lines.sort(key=lambda l: as_i32(round_f32(l[0].bounding_rect().top)))  # :90 (stable)

    all_word_spacings: List[int] = []
    for line in lines:
        if len(line) > 1:
            spacings = []
            for cur, nxt in zip(line, line[1:]):
                d = nxt.bounding_rect().left - cur.bounding_rect().right
                d = d if d > _F0 else _F0  # f32::max(0.)
                spacings.append(as_i32(round_f32(d)))
            spacings.sort()
            all_word_spacings.extend(spacings)
    all_word_spacings.sort()

    median_word_spacing = all_word_spacings[len(all_word_spacings) // 2] if all_word_spacings else 10
    # :116-119 -- words[len/2] of the *unsorted* input
    median_height = as_i32(round_f32(words[len(words) // 2].height() if len(words) else F(10.0)))

    object_bboxes = [r.bounding_rect().integral_bounding_rect() for r in words]
    min_width = median_word_spacing * 3
    min_height = 3 * max(median_height, 0)
    assert min_width >= 0  # `.try_into().unwrap()` (:148)

    out: List[Rect] = []
    it = filter_overlapping(max_empty_rects(object_bboxes, page_rect, _score, min_width, min_height), F(0.5))
    for r in it:
        out.append(r)
        if len(out) == 80:  # .take(80) (:153)
            break
    return out


def find_text_lines(words: Sequence[RotatedRect]) -> List[List[RotatedRect]]:
    """layout_analysis.rs:158-233"""
    separators = find_block_separators(words)
    vertical_separators = []
    horizontal_separators = []
    for r in separators:
        c = r.center()
        vertical_separators.append(LineF(Point.from_yx(r.top, c.x).to_f32(), Point.from_yx(r.bottom, c.x).to_f32()))
        horizontal_separators.append(LineF(Point.from_yx(c.y, r.left).to_f32(), Point.from_yx(c.y, r.right).to_f32()))

    lines = group_into_lines(words, vertical_separators)

    def midpoint_line(ws: Sequence[RotatedRect]) -> LineF:
        assert ws
        return LineF(ws[0].bounding_rect().left_edge().center(), ws[-1].bounding_rect().right_edge().center())

    lines.sort(key=lambda ws: as_i32(midpoint_line(ws).center().y))  # :195 (stable)

    def is_separated_by(a: LineF, b: LineF, seps: Sequence[LineF]) -> bool:
        a_to_b = LineF(a.center(), b.center())
        return any(s.intersects(a_to_b) for s in seps)

    paragraphs: List[List[List[RotatedRect]]] = []
    while lines:
        seed = lines.pop(0)
        para = [seed]
        prev_line = midpoint_line(seed)
        index = 0
        while index < len(lines):
            cand = midpoint_line(lines[index])
            if prev_line.horizontal_overlap(cand) > _F0 and not is_separated_by(prev_line, cand, horizontal_separators):
                para.append(lines.pop(index))
                prev_line = cand
            else:
                index += 1
        paragraphs.append(para)
    return [line for para in paragraphs for line in para]


# ---- layout_analysis/empty_rects.rs ---------------------------------------------------------
class _Partition:
    __slots__ = ("score", "boundary", "obstacles")

    def __init__(self, score, boundary: Rect, obstacles: List[Rect]):
        self.score = F(score)
        self.boundary = boundary
        self.obstacles = obstacles


class _RustBinaryHeap:
    """std::collections::BinaryHeap (max-heap) with the exact sift order of the Rust standard
    library, so that pops among equal scores come out in the same order.
    Ordering key: `score` (f32::total_cmp, empty_rects.rs:20-24; scores here are finite >= 0)."""

    def __init__(self):
        self.data: List[_Partition] = []

    def __len__(self):
        return len(self.data)

    @staticmethod
    def _le(a: _Partition, b: _Partition) -> bool:
        return a.score <= b.score

    def push(self, item: _Partition) -> None:
        old_len = len(self.data)
        self.data.append(item)
        self._sift_up(0, old_len)

    def pop(self) -> Optional[_Partition]:
        if not self.data:
            return None
        item = self.data.pop()
        if self.data:
            item, self.data[0] = self.data[0], item
            self._sift_down_to_bottom(0)
        return item

    def _sift_up(self, start: int, pos: int) -> int:
        d = self.data
        elem = d[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if self._le(elem, d[parent]):
                break
            d[pos] = d[parent]
            pos = parent
        d[pos] = elem
        return pos

    def _sift_down_to_bottom(self, pos: int) -> None:
        d = self.data
        end = len(d)
        start = pos
        elem = d[pos]
        child = 2 * pos + 1
        while child <= max(end - 2, 0) and child + 1 < end:
            if self._le(d[child], d[child + 1]):
                child += 1
            d[pos] = d[child]
            pos = child
            child = 2 * pos + 1
        if child == end - 1:
            d[pos] = d[child]
            pos = child
        d[pos] = elem
        self._sift_up(start, pos)


def max_empty_rects(obstacles: Sequence[Rect], boundary: Rect, score: Callable[[Rect], np.float32],
                    min_width: int, min_height: int):
    """Breuel's maximal-empty-rectangle search, best-first (empty_rects.rs:37-138)."""
    queue = _RustBinaryHeap()
    # :58-61 stable sort by (center.x, center.y)
    obs = sorted(obstacles, key=lambda o: (o.center().x, o.center().y))
    if not boundary.is_empty():
        queue.push(_Partition(score(boundary), boundary, list(obs)))

    while True:
        part = queue.pop()
        if part is None:
            return
        b, obstacles_ = part.boundary, part.obstacles
        if not obstacles_:
            yield b
            continue
        pivot = obstacles_[len(obstacles_) // 2]
        right_rect = Rect.from_tlbr(b.top, pivot.right, b.bottom, b.right)
        left_rect = Rect.from_tlbr(b.top, b.left, b.bottom, pivot.left)
        top_rect = Rect.from_tlbr(b.top, b.left, pivot.top, b.right)
        bottom_rect = Rect.from_tlbr(pivot.bottom, b.left, b.bottom, b.right)
        for sr in (top_rect, left_rect, bottom_rect, right_rect):  # :106
            if max(sr.width(), 0) < min_width or max(sr.height(), 0) < min_height or sr.is_empty():
                continue
            sr_obs = [o for o in obstacles_ if o.intersects(sr)]
            assert len(sr_obs) < len(obstacles_)
            queue.push(_Partition(score(sr), sr, sr_obs))


def filter_overlapping(source, factor):
    """empty_rects.rs:184-221"""
    found: List[Rect] = []
    for r in source:
        if any(f.iou(r) >= factor for f in found):
            continue
        found.append(r)
        yield r
