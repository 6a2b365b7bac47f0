// Host-side layout analysis and recognition geometry (exact integer / fp32 semantics).
//   find_text_lines            <-> ocrs/src/layout_analysis.rs:158-233
//   group_into_lines           <-> layout_analysis.rs:19-71
//   find_block_separators      <-> layout_analysis.rs:83-155
//   max_empty_rects / filter   <-> layout_analysis/empty_rects.rs:37-221
//   line_polygon               <-> recognition.rs:29-55
//   resized_line_width         <-> recognition.rs:58-75
//   polygon_slice_bounding_rect<-> recognition.rs:162-193
#pragma once
#include <cstdint>
#include <vector>

#include "geom.h"

namespace ocrs {
namespace layout {

using geom::LineF;
using geom::PointI;
using geom::RectI;
using geom::RotatedRect;

std::vector<std::vector<RotatedRect>> group_into_lines(const std::vector<RotatedRect>& rects,
                                                       const std::vector<LineF>& separators);
std::vector<RectI> find_block_separators(const std::vector<RotatedRect>& words);
std::vector<std::vector<RotatedRect>> find_text_lines(const std::vector<RotatedRect>& words);

std::vector<PointI> line_polygon(const std::vector<RotatedRect>& words);
uint32_t resized_line_width(int orig_width, int orig_height, int height);
RectI polygon_bounding_rect(const std::vector<PointI>& poly);
// Returns false when no edge overlaps [min_x, max_x].
bool polygon_slice_bounding_rect(const std::vector<PointI>& poly, int min_x, int max_x, RectI* out);
// bounding rect of the words' rotated rects -> integral rect (recognition.rs:432-434)
bool line_integral_rect(const std::vector<RotatedRect>& words, RectI* out);

}  // namespace layout
}  // namespace ocrs
