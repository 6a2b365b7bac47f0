// See text_output.h.  Compiled with -fmad=false / -ffp-contract=off like the rest of the exact host code.
#include "text_output.h"

#include <cmath>
#include <cstdio>

namespace ocrs {
namespace textout {

using geom::PointF;
using geom::RectI;
using geom::RotatedRect;
using geom::Vec2;

RotatedRect orient_towards(const RotatedRect& r, Vec2 up) {
  const Vec2 target = geom::vnormalized(up);
  const float cand[4][2] = {{r.ux, r.uy}, {r.uy, -r.ux}, {-r.ux, -r.uy}, {-r.uy, r.ux}};
  int best = 0;
  float best_d = 0.f;
  for (int i = 0; i < 4; ++i) {
    const float d = cand[i][0] * target.x + cand[i][1] * target.y;
    if (i == 0 || d >= best_d) {  // Rust `max_by` keeps the last maximum
      best = i;
      best_d = d;
    }
  }
  RotatedRect o = r;
  o.ux = cand[best][0];
  o.uy = cand[best][1];
  if (best == 1 || best == 3) {
    o.w = r.h;
    o.h = r.w;
  }
  return o;
}

RotatedRect item_rotated_rect(const RectI* rects, size_t n) {
  OCRS_CHECK(n > 0, kInvalidArg, "text item has no characters");
  std::vector<PointF> pts;
  pts.reserve(n * 4);
  for (size_t i = 0; i < n; ++i) {  // Rect::corners: top-left, top-right, bottom-right, bottom-left
    const RectI& r = rects[i];
    pts.push_back(PointF{(float)r.left, (float)r.top});
    pts.push_back(PointF{(float)r.right, (float)r.top});
    pts.push_back(PointF{(float)r.right, (float)r.bottom});
    pts.push_back(PointF{(float)r.left, (float)r.bottom});
  }
  std::vector<PointF> hull(pts.size());
  const int m = geom::convex_hull(pts.data(), (int)pts.size(), hull.data());
  RotatedRect rr;
  OCRS_CHECK(geom::min_area_rect_of_hull(hull.data(), m, &rr), kInvalidArg, "expected valid rect");  // text_items.rs:25
  return orient_towards(rr, Vec2{0.0f, -1.0f});  // Vec2::from_yx(-1., 0.)
}

void rounded_vertices(const RotatedRect& r, int32_t xy[8]) {
  PointF c[4];
  geom::rr_corners(r, c);
  for (int i = 0; i < 4; ++i) {
    xy[2 * i] = geom::f2i(roundf(c[i].x));      // f32::round: half away from zero; `as i32` saturates
    xy[2 * i + 1] = geom::f2i(roundf(c[i].y));
  }
}

namespace {

void append_utf8(std::string& s, uint32_t cp) {
  if (cp < 0x80) {
    s.push_back((char)cp);
  } else if (cp < 0x800) {
    s.push_back((char)(0xC0 | (cp >> 6)));
    s.push_back((char)(0x80 | (cp & 0x3F)));
  } else if (cp < 0x10000) {
    s.push_back((char)(0xE0 | (cp >> 12)));
    s.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
    s.push_back((char)(0x80 | (cp & 0x3F)));
  } else {
    s.push_back((char)(0xF0 | (cp >> 18)));
    s.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
    s.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
    s.push_back((char)(0x80 | (cp & 0x3F)));
  }
}

std::string chars_to_utf8(const TextChar* c, size_t n) {
  std::string s;
  for (size_t i = 0; i < n; ++i) append_utf8(s, c[i].ch);
  return s;
}

// JSON string literal with serde_json's escaping rules
std::string json_string(const std::string& in) {
  std::string o = "\"";
  for (unsigned char ch : in) {
    switch (ch) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\b': o += "\\b"; break;
      case '\f': o += "\\f"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (ch < 0x20) {
          char buf[8];
          std::snprintf(buf, sizeof buf, "\\u%04x", ch);
          o += buf;
        } else {
          o.push_back((char)ch);
        }
    }
  }
  o.push_back('"');
  return o;
}

struct Pretty {  // serde_json::to_string_pretty layout: 2 spaces, one element per line, `[]` when empty
  std::string out;
  int depth = 0;
  void nl() {
    out.push_back('\n');
    out.append((size_t)depth * 2, ' ');
  }
};

void put_vertices(Pretty& p, const RotatedRect& rr) {
  int32_t xy[8];
  rounded_vertices(rr, xy);
  p.out += "[";
  ++p.depth;
  for (int i = 0; i < 4; ++i) {
    p.nl();
    p.out += "[";
    ++p.depth;
    p.nl();
    p.out += std::to_string(xy[2 * i]) + ",";
    p.nl();
    p.out += std::to_string(xy[2 * i + 1]);
    --p.depth;
    p.nl();
    p.out += i < 3 ? "]," : "]";
  }
  --p.depth;
  p.nl();
  p.out += "]";
}

std::vector<RectI> rects_of(const TextChar* c, size_t n) {
  std::vector<RectI> r(n);
  for (size_t i = 0; i < n; ++i) r[i] = c[i].rect;
  return r;
}

// one {"text", "vertices"[, "words"]} object; keys in serde_json's (sorted) map order
void put_item(Pretty& p, const TextChar* c, size_t n, bool with_words) {
  p.out += "{";
  ++p.depth;
  p.nl();
  p.out += "\"text\": " + json_string(chars_to_utf8(c, n)) + ",";
  p.nl();
  p.out += "\"vertices\": ";
  auto rects = rects_of(c, n);
  put_vertices(p, item_rotated_rect(rects.data(), rects.size()));
  if (with_words) {
    p.out += ",";
    p.nl();
    p.out += "\"words\": [";
    // TextLine::words (text_items.rs:76-82): split on ' ', drop empty pieces
    std::vector<std::pair<size_t, size_t>> words;
    size_t start = 0;
    for (size_t i = 0; i <= n; ++i) {
      if (i == n || c[i].ch == (uint32_t)' ') {
        if (i > start) words.emplace_back(start, i - start);
        start = i + 1;
      }
    }
    if (words.empty()) {
      p.out += "]";
    } else {
      ++p.depth;
      for (size_t w = 0; w < words.size(); ++w) {
        p.nl();
        put_item(p, c + words[w].first, words[w].second, false);
        if (w + 1 < words.size()) p.out += ",";
      }
      --p.depth;
      p.nl();
      p.out += "]";
    }
  }
  --p.depth;
  p.nl();
  p.out += "}";
}

}  // namespace

std::string format_text(const std::vector<TextLine>& lines) {
  std::string out;
  bool first = true;
  for (const auto& l : lines) {
    if (!l.present) continue;  // `.flatten()` drops None lines
    if (!first) out.push_back('\n');
    first = false;
    out += chars_to_utf8(l.chars.data(), l.chars.size());
  }
  return out;
}

std::string format_json(const std::vector<TextLine>& lines, const std::string& input_path, int image_height,
                        int image_width) {
  Pretty p;
  p.out += "{";
  ++p.depth;
  p.nl();
  p.out += "\"image_height\": " + std::to_string(image_height) + ",";
  p.nl();
  p.out += "\"image_width\": " + std::to_string(image_width) + ",";
  p.nl();
  p.out += "\"paragraphs\": [";
  ++p.depth;
  p.nl();
  p.out += "{";
  ++p.depth;
  p.nl();
  p.out += "\"lines\": [";
  size_t n_present = 0;
  for (const auto& l : lines) n_present += l.present ? 1 : 0;
  if (n_present == 0) {
    p.out += "]";
  } else {
    ++p.depth;
    size_t k = 0;
    for (const auto& l : lines) {
      if (!l.present) continue;
      p.nl();
      put_item(p, l.chars.data(), l.chars.size(), true);
      if (++k < n_present) p.out += ",";
    }
    --p.depth;
    p.nl();
    p.out += "]";
  }
  --p.depth;
  p.nl();
  p.out += "}";
  --p.depth;
  p.nl();
  p.out += "],";
  p.nl();
  p.out += "\"url\": " + json_string(input_path);
  --p.depth;
  p.nl();
  p.out += "}";
  return p.out;
}

}  // namespace textout
}  // namespace ocrs
