// Text items and the CLI's output formats (host code).
//   item_rotated_rect   <->  TextItem::rotated_rect        ocrs/src/text_items.rs:18-30
//   rounded_vertices    <->  rounded_vertex_coords         ocrs-cli/src/output.rs:24-27
//   format_text         <->  format_text_output            ocrs-cli/src/output.rs:87-94
//   format_json         <->  format_json_output / ocr_json ocrs-cli/src/output.rs:34-76, 97-100
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "engine.h"
#include "geom.h"

namespace ocrs {
namespace textout {

// Min-area rect of the corners of `n` character boxes, oriented towards up = (x 0, y -1).
geom::RotatedRect item_rotated_rect(const geom::RectI* rects, size_t n);
// RotatedRect::orient_towards (rten-imageproc; call text_items.rs:29)
geom::RotatedRect orient_towards(const geom::RotatedRect& r, geom::Vec2 up);
void rounded_vertices(const geom::RotatedRect& r, int32_t xy[8]);

std::string format_text(const std::vector<TextLine>& lines);
std::string format_json(const std::vector<TextLine>& lines, const std::string& input_path, int image_height,
                        int image_width);

}  // namespace textout
}  // namespace ocrs
