// Bandwidth-bound, bit-exact pixel and index kernels of the hot path (sm_100a), compiled with
// -fmad=false so that every float expression rounds exactly as the reference's Rust does.
//   prepare_image        <-> ocrs/src/preprocess.rs:149-248
//   resize (pad + bilinear, half-pixel)  <-> detection.rs:155-171,187-194 ; recognition.rs:119-122
//   threshold            <-> detection.rs:110
//   connected components + outer contours + RDP + min-area rect <-> detection.rs:41-62
//   polygon crop + resize into the recognition batch <-> recognition.rs:91-158
//   CTC greedy           <-> recognition.rs:498-523 (rten::ctc::CtcDecoder::decode_greedy)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "geom.h"

namespace ocrs {
namespace img {

constexpr float kBlackValue = -0.5f;  // preprocess.rs:128

// dtype: 0 = u8, 1 = f32.  order: 0 = HWC, 1 = CHW.  channels in {1,3,4}.  out: f32 [H*W].
void prepare_image(const void* pixels, int dtype, int order, int H, int W, int C, float* out, cudaStream_t st);

// out[oh, ow] = bilinear(pad(src [H,W] -> [padH,padW] with `pad_value`), half-pixel).
// `n` images with strides (elements) src_stride / dst_stride.
void resize_padded(const float* src, int H, int W, int padH, int padW, float pad_value, float* dst, int OH, int OW,
                   int n, int64_t src_stride, int64_t dst_stride, cudaStream_t st);

// Detection epilogue: takes the network output [inH, inW], uses the top-left [sliceH, sliceW]
// window, resizes to [H, W]; writes the probability map (if prob != null) and the u8 mask
// (prob > threshold).
void resize_threshold(const float* net_out, int inH, int inW, int sliceH, int sliceW, float* prob, uint8_t* mask,
                      int H, int W, float threshold, cudaStream_t st);

void threshold(const float* prob, uint8_t* mask, int64_t n, float thr, cudaStream_t st);

// ---- the same pixel stages for a whole batch of pages in ONE launch each (device tables, page = grid z / y) ----
struct PageResizeIn {   // input of the detection resize: grey page [H, W], virtually padded to [padH, padW]
  const float* src;
  int32_t H, W, padH, padW;
  float sy, sx;         // (float)padH / (float)OH, (float)padW / (float)OW (fill with resize_scale)
};
struct PageResizeOut {  // output of the detection epilogue: slice of the net output -> [H, W] mask (+ prob map)
  const float* net;     // this page's [inH, inW] network output
  float* prob;          // may be null
  uint8_t* mask;
  uint32_t* bits;       // bit-packed mask [H][ceil(W/32)] (may be null)
  int32_t sliceH, sliceW, H, W;
  float sy, sx;         // (float)sliceH / (float)H, (float)sliceW / (float)W
};
inline float resize_scale(int n_in, int n_out) { return (float)n_in / (float)n_out; }
struct PagePrepare {    // u8 HWC RGB page -> grey f32
  const void* src;
  float* dst;
};
void resize_padded_batch(const PageResizeIn* d_tab, int n, float pad_value, float* dst, int OH, int OW, int64_t dst_stride,
                         cudaStream_t st);
void resize_threshold_batch(const PageResizeOut* d_tab, int n, int inW, int maxH, int maxW, float thr, cudaStream_t st);
// true when a page can take the batched fast path of prepare_image (u8, HWC, RGB, aligned, H*W % 4 == 0)
bool prepare_image_batchable(const void* pixels, int dtype, int order, int H, int W, int C, const float* out);
void prepare_image_rgb8_batch(const PagePrepare* d_tab, int n, int H, int W, cudaStream_t st);

// ---- connected components -> word rects -------------------------------------------------------
struct ComponentBuffers {
  int32_t* labels;       // [H*W + 1] union-find parents of run-start pixels (+1 virtual frame node); other entries unused
  int32_t* comp_roots;   // [max_comps] root pixel index of each foreground component (unordered)
  int32_t* counters;     // [8]: [0] n_comps, [2] error flag, [3] n_rects, [4..5] pool top (u64)
  int16_t* pts;          // [pool_cap * 2] traced border points (x, y)
  int32_t* simp_idx;     // [pool_cap] indices of simplified points
  int32_t* stack;        // [pool_cap * 3] RDP work stack
  float* fpts;           // [pool_cap * 2] simplified points as floats
  float* hull;           // [pool_cap * 2]
  geom::RotatedRect* rects;  // [max_comps]
  int32_t* rect_root;    // [max_comps] root pixel index (discovery order key) of each emitted rect
  int64_t pool_cap;
  int32_t max_comps;
};

// One page of a detection batch for labelling + contour extraction.
struct CclPage {
  const uint32_t* bits;  // bit-packed mask [H][wd], bit (x & 31) of word (x >> 5); bits beyond W are 0
  uint16_t* wstart;      // [H][wd] scratch: start x of the run covering bit 0 of each word
  int32_t* labels;       // == bufs.labels
  int32_t H, W, wd;
  ComponentBuffers bufs;
};

// Word rects of a batch of masks (detection.rs:41-62 semantics): outer contours of 8-connected
// components that are not nested inside holes, RDP(eps), min-area rect, expanded by 2*expand_dist,
// kept when area >= min_area.  Results per page: bufs.rects / bufs.rect_root, count in
// bufs.counters[3]; unordered -- sort by rect_root for discovery order.  `d_pages` is the device copy
// of `h_pages`.  Four launches for the whole batch.
void find_component_rects_batch(const CclPage* d_pages, const CclPage* h_pages, int n_pages, float eps, float expand_dist,
                                float min_area, cudaStream_t st);

// ---- line crops ------------------------------------------------------------------------------
struct LineDesc {
  int32_t poly_off, poly_n;       // vertices in the shared polygon array (x, y int32 pairs)
  int32_t top, left, lh, lw;      // polygon bounding rect (canvas = lh x lw)
  int32_t resized_width;          // w'
  int32_t group_width;            // padded width Wg of the destination row
  int64_t dst_off;                // element offset of this line's [out_h, group_width] image in the batch
  int64_t cross_off;              // element offset of this line's per-row crossing table
  int32_t max_cross;              // table width (number of non-horizontal edges)
  int32_t page;                   // page index (for batched pages)
};

// Fills each line's [out_h, group_width] image: polygon-masked copy of the page, bilinear
// resize to [out_h, resized_width], right-padded with BLACK_VALUE.
// pages: array of device pointers to grey pages, page_h/page_w per page.
void crop_lines(const float* const* pages, const int* page_h, const int* page_w, const LineDesc* lines, int n_lines,
                const int32_t* poly_xy, int32_t* cross_scratch, float* dst, int out_h, int max_group_width,
                int max_rows, cudaStream_t st);

// ---- CTC greedy -------------------------------------------------------------------------------
// logits: [T, B, C] (the network's native layout).  excluded: [C] u8 mask or null.
// labels_out/pos_out: [B, T]; counts_out: [B].
void ctc_greedy(const float* logits, int T, int B, int C, const uint8_t* excluded, int32_t* scratch_labels,
                int32_t* labels_out, int32_t* pos_out, int32_t* counts_out, cudaStream_t st);

// Packed variant: one argmax launch over all rows of all width groups, one collapse launch over
// all lines.  logits: [rows, C]; line i reads rows base + t*stride, t < T.
struct CtcLine {
  int64_t base, stride;
  int32_t T, pad;
  int64_t lab_off, pos_off, cnt_off;  // element offsets into `out`
  int64_t node_off;                   // beam search only: first trie node of this line in `nodes`
};
void ctc_greedy_packed(const float* logits, int64_t rows, int C, const uint8_t* excluded, int32_t* row_labels,
                       const CtcLine* lines, int n_lines, int32_t* out, cudaStream_t st);

// CTC prefix beam search (`DecodeMethod::BeamSearch { width }`, ocrs/src/recognition.rs:199-205,
// 512-514), one thread block per line.  Same line descriptors and output layout as the greedy
// decoder; `nodes` is scratch of 6 int32 per trie node, ctc_beam_nodes_per_line(T, width) nodes per
// line starting at CtcLine::node_off.
constexpr int kMaxBeamWidth = 1024;
int64_t ctc_beam_nodes_per_line(int T, int width);
void ctc_beam_search(const float* logits, int C, const uint8_t* excluded, const CtcLine* lines, int n_lines, int width,
                     int32_t* nodes, int32_t* out, cudaStream_t st);

}  // namespace img
}  // namespace ocrs
