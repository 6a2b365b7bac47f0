// Host-side layout analysis (exact).  Compiled with FMA contraction off.
#include "layout.h"

#include <algorithm>
#include <cmath>
#include <functional>

#include "common.h"

namespace ocrs {
namespace layout {

using namespace geom;

namespace {

bool rects_separated_by_line(const RotatedRect& a, const RotatedRect& b, const LineF& l) {  // layout_analysis.rs:8-11
  LineF a_to_b{PointF{a.cx, a.cy}, PointF{b.cx, b.cy}};
  return line_intersects(a_to_b, l);
}

// f32::round (half away from zero) then `as i32`
int round_i32(float v) { return f2i(roundf(v)); }

struct Partition {  // empty_rects.rs:6-10
  float score;
  RectI boundary;
  std::vector<RectI> obstacles;
};

// std::collections::BinaryHeap with Rust's exact sift order: equal scores must
// pop in the same order as the reference.  Ordering: score via f32::total_cmp (finite >= 0 here).
class RustBinaryHeap {
 public:
  bool empty() const { return data_.empty(); }
  void push(Partition&& p) {
    size_t old_len = data_.size();
    data_.push_back(std::move(p));
    sift_up(0, old_len);
  }
  Partition pop() {
    Partition item = std::move(data_.back());
    data_.pop_back();
    if (!data_.empty()) {
      std::swap(item, data_[0]);
      sift_down_to_bottom(0);
    }
    return item;
  }

 private:
  static bool le(const Partition& a, const Partition& b) { return a.score <= b.score; }
  size_t sift_up(size_t start, size_t pos) {
    Partition elem = std::move(data_[pos]);
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (le(elem, data_[parent])) break;
      data_[pos] = std::move(data_[parent]);
      pos = parent;
    }
    data_[pos] = std::move(elem);
    return pos;
  }
  void sift_down_to_bottom(size_t pos) {
    size_t end = data_.size();
    size_t start = pos;
    Partition elem = std::move(data_[pos]);
    size_t child = 2 * pos + 1;
    while (end >= 2 && child <= end - 2) {
      if (le(data_[child], data_[child + 1])) child += 1;
      data_[pos] = std::move(data_[child]);
      pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1) {
      data_[pos] = std::move(data_[child]);
      pos = child;
    }
    data_[pos] = std::move(elem);
    sift_up(start, pos);
  }
  std::vector<Partition> data_;
};

float score_rect(const RectI& r) {  // layout_analysis.rs:127-135
  float aspect_ratio = (float)rheight(r) / (float)rwidth(r);
  float v = fabsf(log2f(aspect_ratio));
  float weight;
  if (v < 3.0f) weight = 0.5f;
  else if (v < 5.0f) weight = 1.5f;
  else weight = v;
  return sqrtf((float)rarea(r) * weight);
}

// Iterator over maximal empty rects (empty_rects.rs:37-138).
class MaxEmptyRects {
 public:
  MaxEmptyRects(const std::vector<RectI>& obstacles, RectI boundary, uint32_t min_w, uint32_t min_h)
      : min_w_(min_w), min_h_(min_h) {
    std::vector<RectI> obs = obstacles;
    std::stable_sort(obs.begin(), obs.end(), [](const RectI& a, const RectI& b) {  // :58-61
      PointI ca = rcenter(a), cb = rcenter(b);
      return ca.x != cb.x ? ca.x < cb.x : ca.y < cb.y;
    });
    if (!rempty(boundary)) queue_.push(Partition{score_rect(boundary), boundary, std::move(obs)});
  }
  bool next(RectI* out) {
    while (!queue_.empty()) {
      Partition part = queue_.pop();
      const RectI b = part.boundary;
      if (part.obstacles.empty()) {
        *out = b;
        return true;
      }
      RectI pivot = part.obstacles[part.obstacles.size() / 2];
      RectI right_rect{b.top, pivot.right, b.bottom, b.right};
      RectI left_rect{b.top, b.left, b.bottom, pivot.left};
      RectI top_rect{b.top, b.left, pivot.top, b.right};
      RectI bottom_rect{pivot.bottom, b.left, b.bottom, b.right};
      const RectI subs[4] = {top_rect, left_rect, bottom_rect, right_rect};  // :106
      for (const RectI& sr : subs) {
        if ((uint32_t)std::max(rwidth(sr), 0) < min_w_ || (uint32_t)std::max(rheight(sr), 0) < min_h_ ||
            rempty(sr))
          continue;
        std::vector<RectI> sr_obs;
        for (const RectI& o : part.obstacles)
          if (rintersects(o, sr)) sr_obs.push_back(o);
        OCRS_CHECK(sr_obs.size() < part.obstacles.size(), kInternal, "max_empty_rects: pivot not excluded");
        queue_.push(Partition{score_rect(sr), sr, std::move(sr_obs)});
      }
    }
    return false;
  }

 private:
  RustBinaryHeap queue_;
  uint32_t min_w_, min_h_;
};

RectF bounding_rect_of(const std::vector<RotatedRect>& rects, bool* ok) {
  *ok = !rects.empty();
  RectF out{0, 0, 0, 0};
  for (size_t i = 0; i < rects.size(); ++i) {
    RectF b = rr_bounding_rect(rects[i]);
    out = i == 0 ? b : rfunion(out, b);
  }
  return out;
}

}  // namespace

std::vector<std::vector<RotatedRect>> group_into_lines(const std::vector<RotatedRect>& rects,
                                                       const std::vector<LineF>& separators) {
  // Per-rect quantities the reference recomputes inside its O(n^2) loop (layout_analysis.rs:46-59)
  // are pure functions of the rect: cache them once.  Semantics (stable sort, filter order,
  // first-minimum tie break) are unchanged.
  struct Item {
    RotatedRect r;
    LineF left_edge, right_edge;
    float left_cx;  // leftmost_edge(r).center().x
    int key;        // r.center().x as i32
    int sort_key;   // bounding_rect().left() as i32
  };
  std::vector<Item> sorted;
  sorted.reserve(rects.size());
  for (const RotatedRect& r : rects) {
    Item it;
    it.r = r;
    it.left_edge = leftmost_edge(r);
    it.right_edge = rightmost_edge(r);
    it.left_cx = line_center(it.left_edge).x;
    it.key = f2i(r.cx);
    it.sort_key = f2i(rr_bounding_rect(r).left);
    sorted.push_back(it);
  }
  std::stable_sort(sorted.begin(), sorted.end(), [](const Item& a, const Item& b) { return a.sort_key < b.sort_key; });  // :21
  std::vector<std::vector<RotatedRect>> lines;
  const float overlap_threshold = 5.0f;  // :27
  const float max_h_overlap = 5.0f;      // :35
  while (!sorted.empty()) {
    std::vector<RotatedRect> line;
    Item last = sorted.front();
    line.push_back(last.r);
    sorted.erase(sorted.begin());
    while (true) {
      const LineF last_edge = last.right_edge;
      const float last_edge_cx = line_center(last_edge).x;
      int best_i = -1, best_key = 0;
      for (size_t i = 0; i < sorted.size(); ++i) {
        const Item& c = sorted[i];
        if (!(c.r.cx > last.r.cx)) continue;
        if (!(c.left_cx - last_edge_cx >= -max_h_overlap)) continue;
        if (!(line_vertical_overlap(last_edge, c.left_edge) >= overlap_threshold)) continue;
        bool separated = false;
        for (const LineF& s : separators)
          if (rects_separated_by_line(last.r, c.r, s)) { separated = true; break; }
        if (separated) continue;
        if (best_i < 0 || c.key < best_key) { best_i = (int)i; best_key = c.key; }  // first minimum (:59)
      }
      if (best_i < 0) break;
      last = sorted[(size_t)best_i];
      line.push_back(last.r);
      sorted.erase(sorted.begin() + best_i);
    }
    lines.push_back(std::move(line));
  }
  return lines;
}

std::vector<RectI> find_block_separators(const std::vector<RotatedRect>& words) {
  bool ok;
  RectF br = bounding_rect_of(words, &ok);
  if (!ok) return {};
  RectI page_rect = rf_integral(br);

  auto lines = group_into_lines(words, {});
  std::stable_sort(lines.begin(), lines.end(),
                   [](const std::vector<RotatedRect>& a, const std::vector<RotatedRect>& b) {  // :90
                     return round_i32(rr_bounding_rect(a.front()).top) < round_i32(rr_bounding_rect(b.front()).top);
                   });
  std::vector<int> all_word_spacings;
  for (const auto& line : lines) {
    if (line.size() > 1) {
      std::vector<int> spacings;
      for (size_t i = 0; i + 1 < line.size(); ++i) {
        float d = rr_bounding_rect(line[i + 1]).left - rr_bounding_rect(line[i]).right;
        d = d > 0.0f ? d : 0.0f;
        spacings.push_back(round_i32(d));
      }
      std::sort(spacings.begin(), spacings.end());
      all_word_spacings.insert(all_word_spacings.end(), spacings.begin(), spacings.end());
    }
  }
  std::sort(all_word_spacings.begin(), all_word_spacings.end());
  int median_word_spacing = all_word_spacings.empty() ? 10 : all_word_spacings[all_word_spacings.size() / 2];
  // :116-119 -- words[len/2] of the unsorted input
  int median_height = round_i32(words.empty() ? 10.0f : words[words.size() / 2].h);

  std::vector<RectI> object_bboxes;
  for (const auto& r : words) object_bboxes.push_back(rf_integral(rr_bounding_rect(r)));
  int min_width = median_word_spacing * 3;
  uint32_t min_height = (uint32_t)(3 * std::max(median_height, 0));
  OCRS_CHECK(min_width >= 0, kInternal, "negative min_width");  // .try_into().unwrap() (:148)

  MaxEmptyRects it(object_bboxes, page_rect, (uint32_t)min_width, min_height);
  std::vector<RectI> found;  // filter_overlapping(0.5).take(80) (:152-153)
  RectI r;
  while (found.size() < 80 && it.next(&r)) {
    bool overlaps = false;
    for (const RectI& f : found)
      if (riou(f, r) >= 0.5f) { overlaps = true; break; }
    if (!overlaps) found.push_back(r);
  }
  return found;
}

std::vector<std::vector<RotatedRect>> find_text_lines(const std::vector<RotatedRect>& words) {
  std::vector<RectI> separators = find_block_separators(words);
  std::vector<LineF> vertical, horizontal;
  for (const RectI& r : separators) {
    PointI c = rcenter(r);
    vertical.push_back(LineF{PointF{(float)c.x, (float)r.top}, PointF{(float)c.x, (float)r.bottom}});
    horizontal.push_back(LineF{PointF{(float)r.left, (float)c.y}, PointF{(float)r.right, (float)c.y}});
  }
  auto lines = group_into_lines(words, vertical);

  auto midpoint_line = [](const std::vector<RotatedRect>& ws) -> LineF {  // :186-192
    RectF first = rr_bounding_rect(ws.front());
    RectF last = rr_bounding_rect(ws.back());
    return LineF{rf_left_edge_center(first), rf_right_edge_center(last)};
  };
  std::stable_sort(lines.begin(), lines.end(),
                   [&](const std::vector<RotatedRect>& a, const std::vector<RotatedRect>& b) {  // :195
                     return f2i(line_center(midpoint_line(a)).y) < f2i(line_center(midpoint_line(b)).y);
                   });
  auto is_separated_by = [&](const LineF& a, const LineF& b) {
    LineF a_to_b{line_center(a), line_center(b)};
    for (const LineF& s : horizontal)
      if (line_intersects(s, a_to_b)) return true;
    return false;
  };
  std::vector<std::vector<RotatedRect>> out;
  while (!lines.empty()) {
    std::vector<RotatedRect> seed = std::move(lines.front());
    lines.erase(lines.begin());
    LineF prev_line = midpoint_line(seed);
    out.push_back(std::move(seed));
    size_t index = 0;
    while (index < lines.size()) {
      LineF cand = midpoint_line(lines[index]);
      if (line_horizontal_overlap(prev_line, cand) > 0.0f && !is_separated_by(prev_line, cand)) {
        out.push_back(std::move(lines[index]));
        lines.erase(lines.begin() + (long)index);
        prev_line = cand;
      } else {
        ++index;
      }
    }
  }
  return out;
}

std::vector<PointI> line_polygon(const std::vector<RotatedRect>& words) {  // recognition.rs:29-55
  std::vector<PointI> poly;
  auto trunc_pt = [](PointF p) { return PointI{f2i(p.x), f2i(p.y)}; };  // `p.y as i32` (:32)
  for (const auto& w : words) {
    LineF left = line_downwards(leftmost_edge(w));
    LineF right = line_downwards(rightmost_edge(w));
    poly.push_back(trunc_pt(left.s));
    poly.push_back(trunc_pt(right.s));
  }
  for (auto it = words.rbegin(); it != words.rend(); ++it) {
    LineF left = line_downwards(leftmost_edge(*it));
    LineF right = line_downwards(rightmost_edge(*it));
    poly.push_back(trunc_pt(right.e));
    poly.push_back(trunc_pt(left.e));
  }
  return poly;
}

uint32_t resized_line_width(int orig_width, int orig_height, int height) {  // recognition.rs:58-75
  float aspect_ratio = (float)orig_width / (float)orig_height;
  float v = (float)height * aspect_ratio;
  if (v == v) v = fminf(fmaxf(v, 10.0f), 2400.0f);  // f32::clamp keeps NaN
  return f2u(v);
}

RectI polygon_bounding_rect(const std::vector<PointI>& poly) {
  RectI r{poly[0].y, poly[0].x, poly[0].y, poly[0].x};
  for (const auto& p : poly) {
    r.top = std::min(r.top, p.y);
    r.left = std::min(r.left, p.x);
    r.bottom = std::max(r.bottom, p.y);
    r.right = std::max(r.right, p.x);
  }
  return r;
}

bool polygon_slice_bounding_rect(const std::vector<PointI>& poly, int min_x, int max_x, RectI* out) {
  bool have = false;
  size_t n = poly.size();
  for (size_t i = 0; i < n; ++i) {
    PointI s = poly[i], e = poly[(i + 1) % n];
    if (s.x > e.x) std::swap(s, e);  // rightwards
    if ((s.x < min_x && e.x < min_x) || (s.x > max_x && e.x > max_x)) continue;
    LineF ef{PointF{(float)s.x, (float)s.y}, PointF{(float)e.x, (float)e.y}};
    float y;
    PointI ts = s, te = e;
    if (line_y_for_x(ef, (float)min_x, &y)) ts = PointI{min_x, round_i32(y)};
    if (line_y_for_x(ef, (float)max_x, &y)) te = PointI{max_x, round_i32(y)};
    RectI br{std::min(ts.y, te.y), std::min(ts.x, te.x), std::max(ts.y, te.y), std::max(ts.x, te.x)};
    *out = have ? runion(*out, br) : br;
    have = true;
  }
  return have;
}

bool line_integral_rect(const std::vector<RotatedRect>& words, RectI* out) {
  bool ok;
  RectF br = bounding_rect_of(words, &ok);
  if (!ok) return false;
  *out = rf_integral(br);
  return true;
}

}  // namespace layout
}  // namespace ocrs
