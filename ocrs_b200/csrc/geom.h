// Geometry types and polygon algorithms shared by host (layout, result assembly) and device
// (contour -> rect) code.  Mirrors the subset of rten-imageproc 0.24 that
// ocrs/src/{detection,layout_analysis,recognition,text_items,geom_util}.rs call.
//
// Every routine evaluates IEEE binary32 in source order.  The translation units that include
// this header are compiled with FMA contraction disabled (-fmad=false / -ffp-contract=off) so
// host and device produce identical bits, and the same bits as the reference's Rust, which never
// contracts.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define OCRS_HD __host__ __device__ __forceinline__
#else
#define OCRS_HD inline
#endif

namespace ocrs {
namespace geom {

// Rust `f32 as i32`: truncate toward zero, saturate, NaN -> 0.
OCRS_HD int32_t f2i(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (-2147483647 - 1);
  return (int32_t)x;
}
// Rust `f32 as u32`.
OCRS_HD uint32_t f2u(float x) {
  if (x != x || x <= 0.0f) return 0u;
  if (x >= 4294967296.0f) return 4294967295u;
  return (uint32_t)x;
}
OCRS_HD float fmin2(float a, float b) { return a < b ? a : b; }
OCRS_HD float fmax2(float a, float b) { return a > b ? a : b; }
OCRS_HD int imin2(int a, int b) { return a < b ? a : b; }
OCRS_HD int imax2(int a, int b) { return a > b ? a : b; }
// Rust i32 `/ 2` truncates toward zero (C++ does too).
OCRS_HD int idiv2(int v) { return v / 2; }

struct PointI { int x, y; };
struct PointF { float x, y; };
struct Vec2 { float x, y; };

OCRS_HD float vlength(Vec2 v) { return sqrtf(v.x * v.x + v.y * v.y); }
OCRS_HD Vec2 vnormalized(Vec2 v) {
  float l = vlength(v);
  return Vec2{v.x / l, v.y / l};
}
OCRS_HD Vec2 vperp(Vec2 v) { return Vec2{-v.y, v.x}; }  // pinned by text_items.rs:139-156
OCRS_HD float pdist(PointF a, PointF b) {
  float dx = a.x - b.x, dy = a.y - b.y;
  return sqrtf(dx * dx + dy * dy);
}

struct LineF {
  PointF s, e;
};
OCRS_HD PointF line_center(LineF l) { return PointF{(l.s.x + l.e.x) / 2.0f, (l.s.y + l.e.y) / 2.0f}; }
OCRS_HD LineF line_downwards(LineF l) { return (l.s.y <= l.e.y) ? l : LineF{l.e, l.s}; }   // geom_util.rs:20-26
OCRS_HD LineF line_rightwards(LineF l) { return (l.s.x <= l.e.x) ? l : LineF{l.e, l.s}; }
OCRS_HD float line_vertical_overlap(LineF a, LineF b) {
  a = line_downwards(a);
  b = line_downwards(b);
  float ov = fmin2(a.e.y, b.e.y) - fmax2(a.s.y, b.s.y);
  return ov > 0.0f ? ov : 0.0f;
}
OCRS_HD float line_horizontal_overlap(LineF a, LineF b) {
  a = line_rightwards(a);
  b = line_rightwards(b);
  float ov = fmin2(a.e.x, b.e.x) - fmax2(a.s.x, b.s.x);
  return ov > 0.0f ? ov : 0.0f;
}
// Segment intersection, Cramer's rule, division free; parallel/coincident -> false.
OCRS_HD bool line_intersects(LineF p, LineF q) {
  float x1 = p.s.x, x2 = p.e.x, y1 = p.s.y, y2 = p.e.y;
  float x3 = q.s.x, x4 = q.e.x, y3 = q.s.y, y4 = q.e.y;
  float a = x2 - x1, b = -(x4 - x3), c = y2 - y1, d = -(y4 - y3);
  float b0 = x3 - x1, b1 = y3 - y1;
  float det_a = a * d - b * c;
  if (det_a == 0.0f) return false;
  float det_a0 = b0 * d - b * b1;
  float det_a1 = a * b1 - b0 * c;
  bool s_ok = ((det_a0 >= 0.0f) == (det_a > 0.0f)) && fabsf(det_a0) <= fabsf(det_a);
  bool t_ok = ((det_a1 >= 0.0f) == (det_a > 0.0f)) && fabsf(det_a1) <= fabsf(det_a);
  return s_ok && t_ok;
}
// Distance from p to the closest point of the segment.
OCRS_HD float line_distance(LineF l, PointF p) {
  if (l.s.x == l.e.x && l.s.y == l.e.y) return pdist(l.s, p);
  float abx = l.e.x - l.s.x, aby = l.e.y - l.s.y;
  float acx = p.x - l.s.x, acy = p.y - l.s.y;
  float ab_len = sqrtf(abx * abx + aby * aby);
  float proj = (acx * abx + acy * aby) / (ab_len * ab_len);
  if (proj <= 0.0f) return pdist(l.s, p);
  if (proj >= 1.0f) return pdist(l.e, p);
  float ix = l.s.x + abx * proj;
  float iy = l.s.y + aby * proj;
  float dx = ix - p.x, dy = iy - p.y;
  return sqrtf(dx * dx + dy * dy);
}
// y on the line at x; false for vertical lines or x outside the segment.
OCRS_HD bool line_y_for_x(LineF l, float x, float* y) {
  float lo = l.s.x <= l.e.x ? l.s.x : l.e.x;
  float hi = l.s.x <= l.e.x ? l.e.x : l.s.x;
  if (x < lo || x > hi) return false;
  float dx = l.e.x - l.s.x;
  if (dx == 0.0f) return false;
  float slope = (l.e.y - l.s.y) / dx;
  float intercept = l.s.y - slope * l.s.x;
  *y = slope * x + intercept;
  return true;
}
OCRS_HD bool line_x_for_y(LineF l, float y, float* x) {
  float lo = l.s.y <= l.e.y ? l.s.y : l.e.y;
  float hi = l.s.y <= l.e.y ? l.e.y : l.s.y;
  if (y < lo || y > hi) return false;
  float dy = l.e.y - l.s.y;
  if (dy == 0.0f) return false;
  float inv_slope = (l.e.x - l.s.x) / dy;
  float intercept = l.s.x - inv_slope * l.s.y;
  *x = inv_slope * y + intercept;
  return true;
}

struct RectI {
  int top, left, bottom, right;
};
OCRS_HD int rwidth(RectI r) { return r.right - r.left; }
OCRS_HD int rheight(RectI r) { return r.bottom - r.top; }
OCRS_HD int rarea(RectI r) { return rwidth(r) * rheight(r); }
OCRS_HD bool rempty(RectI r) { return r.right <= r.left || r.bottom <= r.top; }
OCRS_HD PointI rcenter(RectI r) { return PointI{idiv2(r.left + r.right), idiv2(r.top + r.bottom)}; }
OCRS_HD bool rintersects(RectI a, RectI b) {
  return a.left < b.right && a.right > b.left && a.top < b.bottom && a.bottom > b.top;
}
OCRS_HD RectI runion(RectI a, RectI b) {
  return RectI{imin2(a.top, b.top), imin2(a.left, b.left), imax2(a.bottom, b.bottom), imax2(a.right, b.right)};
}
OCRS_HD RectI rintersection(RectI a, RectI b) {
  int t = imax2(a.top, b.top), l = imax2(a.left, b.left);
  int bo = imin2(a.bottom, b.bottom), r = imin2(a.right, b.right);
  if (bo < t || r < l) return RectI{t, l, t, l};
  return RectI{t, l, bo, r};
}
OCRS_HD float riou(RectI a, RectI b) {
  int inter = rarea(rintersection(a, b));
  int uni = rarea(a) + rarea(b) - inter;
  return (float)inter / (float)uni;
}
OCRS_HD bool rcontains_point(RectI r, int y, int x) {  // inclusive on all sides
  return r.top <= y && y <= r.bottom && r.left <= x && x <= r.right;
}

struct RectF {
  float top, left, bottom, right;
};
OCRS_HD RectF rfunion(RectF a, RectF b) {
  return RectF{fmin2(a.top, b.top), fmin2(a.left, b.left), fmax2(a.bottom, b.bottom), fmax2(a.right, b.right)};
}
OCRS_HD RectI rf_integral(RectF r) {
  return RectI{f2i(floorf(r.top)), f2i(floorf(r.left)), f2i(ceilf(r.bottom)), f2i(ceilf(r.right))};
}
OCRS_HD PointF rf_left_edge_center(RectF r) { return PointF{(r.left + r.left) / 2.0f, (r.top + r.bottom) / 2.0f}; }
OCRS_HD PointF rf_right_edge_center(RectF r) { return PointF{(r.right + r.right) / 2.0f, (r.top + r.bottom) / 2.0f}; }

// Oriented rect; layout identical to the C ABI's ocrs_b200_rotated_rect.
struct RotatedRect {
  float cx, cy;  // centre
  float ux, uy;  // unit "up" axis
  float w, h;    // extent perpendicular to / along `up`
};
OCRS_HD RotatedRect rr_new(PointF c, Vec2 up, float w, float h) {
  Vec2 n = vnormalized(up);
  return RotatedRect{c.x, c.y, n.x, n.y, w, h};
}
OCRS_HD RotatedRect rr_from_rect(RectF r) {
  PointF c{(r.left + r.right) / 2.0f, (r.top + r.bottom) / 2.0f};
  return rr_new(c, Vec2{0.0f, 1.0f}, r.right - r.left, r.bottom - r.top);
}
// [c-U+P, c-U-P, c+U-P, c+U+P], U = up*h/2, P = perp(up)*w/2
OCRS_HD void rr_corners(const RotatedRect& r, PointF out[4]) {
  float hw = r.w / 2.0f, hh = r.h / 2.0f;
  float px = (-r.uy) * hw, py = r.ux * hw;
  float ux = r.ux * hh, uy = r.uy * hh;
  out[0] = PointF{r.cx - ux + px, r.cy - uy + py};
  out[1] = PointF{r.cx - ux - px, r.cy - uy - py};
  out[2] = PointF{r.cx + ux - px, r.cy + uy - py};
  out[3] = PointF{r.cx + ux + px, r.cy + uy + py};
}
OCRS_HD RectF rr_bounding_rect(const RotatedRect& r) {
  PointF c[4];
  rr_corners(r, c);
  RectF o{c[0].y, c[0].x, c[0].y, c[0].x};
  for (int i = 1; i < 4; ++i) {
    o.top = fmin2(o.top, c[i].y);
    o.left = fmin2(o.left, c[i].x);
    o.bottom = fmax2(o.bottom, c[i].y);
    o.right = fmax2(o.right, c[i].x);
  }
  return o;
}
OCRS_HD float rr_area(const RotatedRect& r) { return r.h * r.w; }

// Stable sort of the 4 corners by x (geom_util.rs:8,15); insertion sort keeps ties in order.
OCRS_HD void rr_corners_sorted_by_x(const RotatedRect& r, PointF c[4]) {
  rr_corners(r, c);
  for (int i = 1; i < 4; ++i) {
    PointF k = c[i];
    int j = i - 1;
    while (j >= 0 && c[j].x > k.x) {
      c[j + 1] = c[j];
      --j;
    }
    c[j + 1] = k;
  }
}
OCRS_HD LineF rightmost_edge(const RotatedRect& r) {  // geom_util.rs:6-10
  PointF c[4];
  rr_corners_sorted_by_x(r, c);
  return LineF{c[2], c[3]};
}
OCRS_HD LineF leftmost_edge(const RotatedRect& r) {  // geom_util.rs:13-17
  PointF c[4];
  rr_corners_sorted_by_x(r, c);
  return LineF{c[0], c[1]};
}

// Gift-wrapping convex hull from the first left-most point; `hull` must hold n entries.
// Returns the number of hull vertices.  See oracle/geometry.py:convex_hull for the contract.
OCRS_HD int convex_hull(const PointF* pts, int n, PointF* hull) {
  if (n == 0) return 0;
  int start = 0;
  for (int i = 1; i < n; ++i)
    if (pts[i].x < pts[start].x) start = i;
  int m = 0;
  int cur = start;
  while (true) {
    hull[m++] = pts[cur];
    int nxt = -1;
    for (int i = 0; i < n; ++i) {
      if (pts[i].x == pts[cur].x && pts[i].y == pts[cur].y) continue;
      if (nxt < 0) { nxt = i; continue; }
      float ax = pts[nxt].x - pts[cur].x, ay = pts[nxt].y - pts[cur].y;
      float bx = pts[i].x - pts[cur].x, by = pts[i].y - pts[cur].y;
      float cross = ax * by - ay * bx;
      if (cross < 0.0f) {
        nxt = i;
      } else if (cross == 0.0f) {
        if (bx * bx + by * by > ax * ax + ay * ay) nxt = i;
      }
    }
    if (nxt < 0) break;
    if (pts[nxt].x == pts[start].x && pts[nxt].y == pts[start].y) break;
    cur = nxt;
    if (m >= n) break;
  }
  return m;
}

// Exhaustive hull-edge min-area rectangle; first strictly smaller area wins.
OCRS_HD bool min_area_rect_of_hull(const PointF* hull, int m, RotatedRect* out) {
  if (m < 2) return false;
  bool have = false;
  float best_area = 0.0f;
  for (int i = 0; i < m; ++i) {
    PointF s = hull[i], e = hull[(i + 1) % m];
    Vec2 par = vnormalized(Vec2{e.x - s.x, e.y - s.y});
    Vec2 perp = vperp(par);
    float min_par = 3.402823466e+38f, max_par = -3.402823466e+38f, max_perp = -3.402823466e+38f;
    for (int k = 0; k < m; ++k) {
      float dx = hull[k].x - s.x, dy = hull[k].y - s.y;
      float par_proj = par.x * dx + par.y * dy;
      float perp_proj = perp.x * dx + perp.y * dy;
      min_par = fmin2(min_par, par_proj);
      max_par = fmax2(max_par, par_proj);
      max_perp = fmax2(max_perp, perp_proj);
    }
    float height = max_perp;
    float width = max_par - min_par;
    float area = height * width;
    if (!have || area < best_area) {
      float cy = s.y + (par.y * (min_par + max_par) / 2.0f) + (perp.y * height / 2.0f);
      float cx = s.x + (par.x * (min_par + max_par) / 2.0f) + (perp.x * height / 2.0f);
      *out = rr_new(PointF{cx, cy}, perp, width, height);
      best_area = area;
      have = true;
    }
  }
  return have;
}

}  // namespace geom
}  // namespace ocrs
