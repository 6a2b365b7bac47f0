"""Geometry types and polygon algorithms of the hot path (oracle; test infrastructure).

Restates the subset of the un-vendored crate `rten-imageproc` 0.24.0 (Cargo.lock:726) that
`ocrs/src/{detection,layout_analysis,recognition,text_items,geom_util}.rs` call.  The crate
source is not in /root/reference, so these follow the published algorithms and the semantics
pinned by the reference's own tests; parity with rten is otherwise unpinned.

All float arithmetic is IEEE binary32 evaluated in source order with no FMA contraction
(Rust never contracts); that is why every value is an `np.float32` scalar.
Integer conversions follow Rust `as` casts (truncate toward zero, saturating, NaN -> 0)
and `f32::round` (half away from zero).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

F = np.float32
_F0 = F(0.0)
_F1 = F(1.0)
_F2 = F(2.0)
_HALF = F(0.5)

I32_MAX = 2**31 - 1
I32_MIN = -(2**31)

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.log2f.restype = ctypes.c_float
_libm.log2f.argtypes = [ctypes.c_float]


def log2f(x) -> np.float32:
    """glibc log2f: the same libm routine Rust's `f32::log2` lowers to on linux."""
    return F(_libm.log2f(float(x)))


def as_i32(x) -> int:
    """Rust `f32 as i32`."""
    x = float(x)
    if math.isnan(x):
        return 0
    if x >= 2147483648.0:
        return I32_MAX
    if x <= -2147483648.0:
        return I32_MIN
    return int(x)  # trunc toward zero


def as_u32(x) -> int:
    """Rust `f32 as u32`."""
    x = float(x)
    if math.isnan(x) or x <= 0.0:
        return 0
    if x >= 4294967296.0:
        return 2**32 - 1
    return int(x)


def round_f32(x) -> np.float32:
    """Rust `f32::round`: half away from zero (exact, unlike x+0.5)."""
    x = F(x)
    r = np.trunc(x)
    if abs(x - r) >= _HALF:
        r = r + (F(1.0) if x > 0 else F(-1.0))
    return F(r)


def fsqrt(x) -> np.float32:
    return np.sqrt(F(x))


# ----------------------------------------------------------------------------------------------
# Points / vectors
# ----------------------------------------------------------------------------------------------
class Point:
    """Integer point (rten_imageproc::Point<i32>)."""

    __slots__ = ("x", "y")

    def __init__(self, x: int, y: int):
        self.x = int(x)
        self.y = int(y)

    @staticmethod
    def from_yx(y, x) -> "Point":
        return Point(x, y)

    def to_f32(self) -> "PointF":
        return PointF(F(self.x), F(self.y))

    def __eq__(self, o):
        return isinstance(o, Point) and self.x == o.x and self.y == o.y

    def __hash__(self):
        return hash((self.x, self.y))

    def __repr__(self):
        return f"Point(y={self.y}, x={self.x})"


class PointF:
    """Float point (rten_imageproc::PointF)."""

    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x = F(x)
        self.y = F(y)

    @staticmethod
    def from_yx(y, x) -> "PointF":
        return PointF(x, y)

    def distance(self, o: "PointF") -> np.float32:
        dx = self.x - o.x
        dy = self.y - o.y
        return fsqrt(dx * dx + dy * dy)

    def __eq__(self, o):
        return isinstance(o, PointF) and self.x == o.x and self.y == o.y

    def __repr__(self):
        return f"PointF(y={float(self.y)}, x={float(self.x)})"


class Vec2:
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x = F(x)
        self.y = F(y)

    @staticmethod
    def from_yx(y, x) -> "Vec2":
        return Vec2(x, y)

    def length(self) -> np.float32:
        return fsqrt(self.x * self.x + self.y * self.y)

    def normalized(self) -> "Vec2":
        ln = self.length()
        return Vec2(self.x / ln, self.y / ln)

    def dot(self, o: "Vec2") -> np.float32:
        return self.x * o.x + self.y * o.y

    def perpendicular(self) -> "Vec2":
        """(x, y) -> (-y, x): pinned by text_items.rs:139-156 (corner order)."""
        return Vec2(-self.y, self.x)

    def __repr__(self):
        return f"Vec2(y={float(self.y)}, x={float(self.x)})"


# ----------------------------------------------------------------------------------------------
# Lines
# ----------------------------------------------------------------------------------------------
class LineF:
    __slots__ = ("start", "end")

    def __init__(self, start: PointF, end: PointF):
        self.start = start
        self.end = end

    def center(self) -> PointF:
        return PointF((self.start.x + self.end.x) / _F2, (self.start.y + self.end.y) / _F2)

    def downwards(self) -> "LineF":
        # geom_util.rs:20-26 semantics (start.y <= end.y keeps the order)
        if self.start.y <= self.end.y:
            return self
        return LineF(self.end, self.start)

    def rightwards(self) -> "LineF":
        if self.start.x <= self.end.x:
            return self
        return LineF(self.end, self.start)

    def vertical_overlap(self, o: "LineF") -> np.float32:
        a, b = self.downwards(), o.downwards()
        ov = min(a.end.y, b.end.y) - max(a.start.y, b.start.y)
        return ov if ov > _F0 else _F0

    def horizontal_overlap(self, o: "LineF") -> np.float32:
        a, b = self.rightwards(), o.rightwards()
        ov = min(a.end.x, b.end.x) - max(a.start.x, b.start.x)
        return ov if ov > _F0 else _F0

    def intersects(self, o: "LineF") -> bool:
        """Segment/segment test, Cramer's rule, division free; parallel/coincident -> False;
        parameters inclusive on [0, 1]."""
        x1, x2, y1, y2 = self.start.x, self.end.x, self.start.y, self.end.y
        x3, x4, y3, y4 = o.start.x, o.end.x, o.start.y, o.end.y
        a = x2 - x1
        b = -(x4 - x3)
        c = y2 - y1
        d = -(y4 - y3)
        b0 = x3 - x1
        b1 = y3 - y1
        det_a = a * d - b * c
        if det_a == _F0:
            return False
        det_a0 = b0 * d - b * b1
        det_a1 = a * b1 - b0 * c
        s_ok = ((det_a0 >= _F0) == (det_a > _F0)) and abs(det_a0) <= abs(det_a)
        t_ok = ((det_a1 >= _F0) == (det_a > _F0)) and abs(det_a1) <= abs(det_a)
        return bool(s_ok and t_ok)

    def is_empty(self) -> bool:
        return self.start.x == self.end.x and self.start.y == self.end.y

    def distance(self, p: PointF) -> np.float32:
        """Distance from `p` to the closest point of the *segment* (comp.graphics.algorithms
        FAQ 1.02); degenerates to point distance when start == end."""
        if self.is_empty():
            return self.start.distance(p)
        abx = self.end.x - self.start.x
        aby = self.end.y - self.start.y
        acx = p.x - self.start.x
        acy = p.y - self.start.y
        ab_len = fsqrt(abx * abx + aby * aby)
        scalar_proj = (acx * abx + acy * aby) / (ab_len * ab_len)
        if scalar_proj <= _F0:
            return self.start.distance(p)
        if scalar_proj >= _F1:
            return self.end.distance(p)
        ix = self.start.x + abx * scalar_proj
        iy = self.start.y + aby * scalar_proj
        dx = ix - p.x
        dy = iy - p.y
        return fsqrt(dx * dx + dy * dy)

    def y_for_x(self, x) -> Optional[np.float32]:
        """None for vertical lines or x outside the segment (recognition.rs:176-184 relies on it)."""
        x = F(x)
        lo, hi = (self.start.x, self.end.x) if self.start.x <= self.end.x else (self.end.x, self.start.x)
        if x < lo or x > hi:
            return None
        dx = self.end.x - self.start.x
        if dx == _F0:
            return None
        slope = (self.end.y - self.start.y) / dx
        intercept = self.start.y - slope * self.start.x
        return slope * x + intercept

    def x_for_y(self, y) -> Optional[np.float32]:
        y = F(y)
        lo, hi = (self.start.y, self.end.y) if self.start.y <= self.end.y else (self.end.y, self.start.y)
        if y < lo or y > hi:
            return None
        dy = self.end.y - self.start.y
        if dy == _F0:
            return None
        inv_slope = (self.end.x - self.start.x) / dy
        intercept = self.start.x - inv_slope * self.start.y
        return inv_slope * y + intercept

    def __repr__(self):
        return f"LineF({self.start} -> {self.end})"


class Line:
    """Integer line segment."""

    __slots__ = ("start", "end")

    def __init__(self, start: Point, end: Point):
        self.start = start
        self.end = end

    def rightwards(self) -> "Line":
        if self.start.x <= self.end.x:
            return self
        return Line(self.end, self.start)

    def downwards(self) -> "Line":
        if self.start.y <= self.end.y:
            return self
        return Line(self.end, self.start)

    def to_f32(self) -> LineF:
        return LineF(self.start.to_f32(), self.end.to_f32())

    def bounding_rect(self) -> "Rect":
        return Rect(
            min(self.start.y, self.end.y),
            min(self.start.x, self.end.x),
            max(self.start.y, self.end.y),
            max(self.start.x, self.end.x),
        )


# ----------------------------------------------------------------------------------------------
# Rects
# ----------------------------------------------------------------------------------------------
def _idiv2(v: int) -> int:
    """Rust i32 `/ 2` (truncates toward zero)."""
    return int(v / 2) if v < 0 else v // 2


class Rect:
    """Integer rect, rten_imageproc::Rect<i32>; right/bottom exclusive by convention."""

    __slots__ = ("top", "left", "bottom", "right")

    def __init__(self, top: int, left: int, bottom: int, right: int):
        self.top, self.left, self.bottom, self.right = int(top), int(left), int(bottom), int(right)

    @staticmethod
    def from_tlbr(t, l, b, r) -> "Rect":
        return Rect(t, l, b, r)

    @staticmethod
    def from_tlhw(t, l, h, w) -> "Rect":
        return Rect(t, l, t + h, l + w)

    @staticmethod
    def from_hw(h, w) -> "Rect":
        return Rect(0, 0, h, w)

    def width(self) -> int:
        return self.right - self.left

    def height(self) -> int:
        return self.bottom - self.top

    def area(self) -> int:
        return self.width() * self.height()

    def is_empty(self) -> bool:
        return self.right <= self.left or self.bottom <= self.top

    def center(self) -> Point:
        return Point(_idiv2(self.left + self.right), _idiv2(self.top + self.bottom))

    def adjust_tlbr(self, t, l, b, r) -> "Rect":
        return Rect(self.top + t, self.left + l, self.bottom + b, self.right + r)

    def contains_point(self, p: Point) -> bool:
        """Inclusive on all four sides (hence adjust_tlbr(0,0,-1,-1) at recognition.rs:100)."""
        return self.top <= p.y <= self.bottom and self.left <= p.x <= self.right

    def contains(self, o: "Rect") -> bool:
        return self.left <= o.left and self.right >= o.right and self.top <= o.top and self.bottom >= o.bottom

    def intersects(self, o: "Rect") -> bool:
        """Strict: touching rects do not intersect."""
        return self.left < o.right and self.right > o.left and self.top < o.bottom and self.bottom > o.top

    def union(self, o: "Rect") -> "Rect":
        return Rect(min(self.top, o.top), min(self.left, o.left), max(self.bottom, o.bottom), max(self.right, o.right))

    def intersection(self, o: "Rect") -> "Rect":
        t, l = max(self.top, o.top), max(self.left, o.left)
        b, r = min(self.bottom, o.bottom), min(self.right, o.right)
        if b < t or r < l:
            return Rect(t, l, t, l)  # empty
        return Rect(t, l, b, r)

    def iou(self, o: "Rect") -> np.float32:
        inter = self.intersection(o).area()
        union = self.area() + o.area() - inter
        return F(inter) / F(union)

    def to_f32(self) -> "RectF":
        return RectF(F(self.top), F(self.left), F(self.bottom), F(self.right))

    def corners(self) -> List[Point]:
        """top-left, top-right, bottom-right, bottom-left."""
        return [
            Point(self.left, self.top),
            Point(self.right, self.top),
            Point(self.right, self.bottom),
            Point(self.left, self.bottom),
        ]

    def tlbr(self) -> Tuple[int, int, int, int]:
        return (self.top, self.left, self.bottom, self.right)

    def __eq__(self, o):
        return isinstance(o, Rect) and self.tlbr() == o.tlbr()

    def __hash__(self):
        return hash(self.tlbr())

    def __repr__(self):
        return f"Rect(t={self.top}, l={self.left}, b={self.bottom}, r={self.right})"


class RectF:
    __slots__ = ("top", "left", "bottom", "right")

    def __init__(self, top, left, bottom, right):
        self.top, self.left, self.bottom, self.right = F(top), F(left), F(bottom), F(right)

    def width(self) -> np.float32:
        return self.right - self.left

    def height(self) -> np.float32:
        return self.bottom - self.top

    def center(self) -> PointF:
        return PointF((self.left + self.right) / _F2, (self.top + self.bottom) / _F2)

    def union(self, o: "RectF") -> "RectF":
        return RectF(min(self.top, o.top), min(self.left, o.left), max(self.bottom, o.bottom), max(self.right, o.right))

    def integral_bounding_rect(self) -> Rect:
        """floor(top,left), ceil(bottom,right) -> i32."""
        return Rect(
            as_i32(np.floor(self.top)), as_i32(np.floor(self.left)),
            as_i32(np.ceil(self.bottom)), as_i32(np.ceil(self.right)),
        )

    def left_edge(self) -> LineF:
        return LineF(PointF(self.left, self.top), PointF(self.left, self.bottom))

    def right_edge(self) -> LineF:
        return LineF(PointF(self.right, self.top), PointF(self.right, self.bottom))

    def tlbr(self):
        return (float(self.top), float(self.left), float(self.bottom), float(self.right))

    def __eq__(self, o):
        return isinstance(o, RectF) and self.tlbr() == o.tlbr()

    def __repr__(self):
        return "RectF(t=%g, l=%g, b=%g, r=%g)" % self.tlbr()


class RotatedRect:
    """Oriented rect: centre, unit `up` axis, width (extent perpendicular to up), height."""

    __slots__ = ("cx", "cy", "ux", "uy", "w", "h")

    def __init__(self, center: PointF, up: Vec2, width, height, _normalize=True):
        if _normalize:
            up = up.normalized()
        self.cx, self.cy = F(center.x), F(center.y)
        self.ux, self.uy = F(up.x), F(up.y)
        self.w, self.h = F(width), F(height)

    @staticmethod
    def from_raw(cx, cy, ux, uy, w, h) -> "RotatedRect":
        r = RotatedRect.__new__(RotatedRect)
        r.cx, r.cy, r.ux, r.uy, r.w, r.h = F(cx), F(cy), F(ux), F(uy), F(w), F(h)
        return r

    @staticmethod
    def from_rect(r: RectF) -> "RotatedRect":
        return RotatedRect(r.center(), Vec2.from_yx(1.0, 0.0), r.width(), r.height())

    def raw(self) -> Tuple[float, ...]:
        return tuple(float(v) for v in (self.cx, self.cy, self.ux, self.uy, self.w, self.h))

    def center(self) -> PointF:
        return PointF(self.cx, self.cy)

    def up_axis(self) -> Vec2:
        return Vec2(self.ux, self.uy)

    def width(self):
        return self.w

    def height(self):
        return self.h

    def area(self) -> np.float32:
        return self.h * self.w

    def resize(self, width, height) -> None:
        self.w, self.h = F(width), F(height)

    def corners(self) -> List[PointF]:
        """[c-U+P, c-U-P, c+U-P, c+U+P], U = up*h/2, P = perp(up)*w/2 (text_items.rs:139-156)."""
        hw = self.w / _F2
        hh = self.h / _F2
        px, py = (-self.uy) * hw, self.ux * hw  # par_offset = perpendicular(up) * (w/2)
        ux, uy = self.ux * hh, self.uy * hh  # perp_offset = up * (h/2)
        cx, cy = self.cx, self.cy
        return [
            PointF(cx - ux + px, cy - uy + py),
            PointF(cx - ux - px, cy - uy - py),
            PointF(cx + ux - px, cy + uy - py),
            PointF(cx + ux + px, cy + uy + py),
        ]

    def bounding_rect(self) -> RectF:
        cs = self.corners()
        xs = [c.x for c in cs]
        ys = [c.y for c in cs]
        return RectF(min(ys), min(xs), max(ys), max(xs))

    def orient_towards(self, up: Vec2) -> "RotatedRect":
        """Among the four 90-degree rotations pick the one whose up is most aligned with `up`
        (Rust `max_by` keeps the LAST maximum)."""
        target = up.normalized()
        cands = [
            (self.ux, self.uy),
            (self.uy, -self.ux),
            (-self.ux, -self.uy),
            (-self.uy, self.ux),
        ]
        best, best_d = 0, None
        for i, (x, y) in enumerate(cands):
            d = x * target.x + y * target.y
            if best_d is None or d >= best_d:
                best, best_d = i, d
        x, y = cands[best]
        if best in (0, 2):
            return RotatedRect(self.center(), Vec2(x, y), self.w, self.h)
        return RotatedRect(self.center(), Vec2(x, y), self.h, self.w)

    def __repr__(self):
        return "RotatedRect(c=(y=%g,x=%g), up=(y=%g,x=%g), w=%g, h=%g)" % (
            float(self.cy), float(self.cx), float(self.uy), float(self.ux), float(self.w), float(self.h))


def bounding_rect_of_rects(rects: Iterable[RectF]) -> Optional[RectF]:
    out = None
    for r in rects:
        out = r if out is None else out.union(r)
    return out


def bounding_rect_of_rotated(rects: Iterable[RotatedRect]) -> Optional[RectF]:
    return bounding_rect_of_rects(r.bounding_rect() for r in rects)


def bounding_rect_of_int_rects(rects: Iterable[Rect]) -> Optional[Rect]:
    out = None
    for r in rects:
        out = r if out is None else out.union(r)
    return out


# ----------------------------------------------------------------------------------------------
# Polygon algorithms
# ----------------------------------------------------------------------------------------------
def polygon_edges(points: Sequence[Point]) -> List[Line]:
    n = len(points)
    return [Line(points[i], points[(i + 1) % n]) for i in range(n)]


def polygon_bounding_rect(points: Sequence[Point]) -> Rect:
    xs = [p.x for p in points]
    ys = [p.y for p in points]
    return Rect(min(ys), min(xs), max(ys), max(xs))


def polygon_fill_rows(points: Sequence[Point]) -> List[Tuple[int, np.ndarray]]:
    """Even-odd scanline fill (rten-imageproc `Polygon::fill_iter`, call recognition.rs:110).

    Non-horizontal edges only, each active for start.y <= y < end.y (downwards orientation);
    pixel (y, x), x in [left, right), is inside when the number of active edges whose
    x-intercept `round(x_for_y(y))` is <= x is odd.  Returns [(y, xs int array)] in raster order.
    """
    br = polygon_bounding_rect(points)
    edges = [e.downwards() for e in polygon_edges(points) if e.start.y != e.end.y]
    rows = []
    if br.is_empty():
        return rows
    xs_all = np.arange(br.left, br.right, dtype=np.int64)
    for y in range(br.top, br.bottom):
        cross = []
        for e in edges:
            if e.start.y <= y < e.end.y:
                xf = e.to_f32().x_for_y(F(y))
                cross.append(as_i32(round_f32(xf)))
        if not cross:
            continue
        cnt = np.zeros(xs_all.shape, dtype=np.int64)
        for c in cross:
            cnt += (c <= xs_all)
        sel = xs_all[(cnt & 1) == 1]
        if sel.size:
            rows.append((y, sel))
    return rows


def polygon_fill_points(points: Sequence[Point]) -> List[Point]:
    return [Point(int(x), y) for y, xs in polygon_fill_rows(points) for x in xs]


def polygon_contains_pixel(points: Sequence[Point], p: Point) -> bool:
    for y, xs in polygon_fill_rows(points):
        if y == p.y:
            return bool((xs == p.x).any())
    return False


def polygon_is_simple(points: Sequence[Point]) -> bool:
    """No two non-adjacent edges intersect."""
    edges = polygon_edges(points)
    n = len(edges)
    for i in range(n):
        for j in range(i + 1, n):
            if j == i + 1 or (i == 0 and j == n - 1):
                continue
            if edges[i].to_f32().intersects(edges[j].to_f32()):
                return False
    return True


def _simplify_polyline(points: Sequence[PointF], eps, out: List[PointF], keep_last: bool) -> None:
    if len(points) <= 1:
        if points:
            out.append(points[0])
        return
    seg = LineF(points[0], points[-1])
    max_i, max_d = 0, _F0
    for i in range(1, len(points) - 1):
        d = seg.distance(points[i])
        if d >= max_d:  # keeps the LAST farthest point
            max_i, max_d = i, d
    if max_d > eps:
        _simplify_polyline(points[: max_i + 1], eps, out, False)
        _simplify_polyline(points[max_i:], eps, out, keep_last)
    else:
        out.append(seg.start)
        if keep_last:
            out.append(seg.end)


def simplify_polygon(points: Sequence[PointF], eps) -> List[PointF]:
    """Ramer-Douglas-Peucker on the polygon closed by repeating points[0]
    (rten-imageproc `simplify_polygon`, call detection.rs:50 with eps = 2)."""
    eps = F(eps)
    polyline = list(points) + [points[0]]
    out: List[PointF] = []
    _simplify_polyline(polyline, eps, out, True)
    return out[:-1]


def convex_hull(points: Sequence[PointF]) -> List[PointF]:
    """Gift wrapping from the first left-most point.  The next hull vertex is the candidate with
    every other point on or to one side (cross >= 0 test below); among collinear candidates the
    farthest wins.  Duplicate points collapse.  (rten-imageproc `convex_hull`; restated.)"""
    n = len(points)
    if n == 0:
        return []
    start = 0
    for i in range(1, n):
        if points[i].x < points[start].x:
            start = i
    hull: List[PointF] = []
    cur = start
    while True:
        hull.append(points[cur])
        nxt = -1
        for i in range(n):
            if points[i].x == points[cur].x and points[i].y == points[cur].y:
                continue
            if nxt < 0:
                nxt = i
                continue
            ax = points[nxt].x - points[cur].x
            ay = points[nxt].y - points[cur].y
            bx = points[i].x - points[cur].x
            by = points[i].y - points[cur].y
            cross = ax * by - ay * bx
            if cross < _F0:
                nxt = i
            elif cross == _F0:
                if bx * bx + by * by > ax * ax + ay * ay:
                    nxt = i
        if nxt < 0:
            break
        if points[nxt].x == points[start].x and points[nxt].y == points[start].y:
            break
        cur = nxt
        if len(hull) > n:  # safety against numeric cycles
            break
    return hull


def min_area_rect(points: Sequence[PointF]) -> Optional[RotatedRect]:
    """Exhaustive search over hull edges (geometrictools "MinimumAreaRectangle");
    first strictly-smaller area wins (rten-imageproc `min_area_rect`, call detection.rs:52)."""
    hull = convex_hull(points)
    if len(hull) < 2:
        return None
    best = None
    best_area = None
    m = len(hull)
    for i in range(m):
        s, e = hull[i], hull[(i + 1) % m]
        par = Vec2(e.x - s.x, e.y - s.y).normalized()
        perp = par.perpendicular()
        min_par, max_par, max_perp = F(np.finfo(np.float32).max), F(np.finfo(np.float32).min), F(np.finfo(np.float32).min)
        for p in hull:
            dx = p.x - s.x
            dy = p.y - s.y
            par_proj = par.x * dx + par.y * dy
            perp_proj = perp.x * dx + perp.y * dy
            min_par = min(min_par, par_proj)
            max_par = max(max_par, par_proj)
            max_perp = max(max_perp, perp_proj)
        height = max_perp
        width = max_par - min_par
        area = height * width
        if best_area is None or area < best_area:
            cy = s.y + (par.y * (min_par + max_par) / _F2) + (perp.y * height / _F2)
            cx = s.x + (par.x * (min_par + max_par) / _F2) + (perp.x * height / _F2)
            best = RotatedRect(PointF(cx, cy), perp, width, height)
            best_area = area
    return best
