// Bit-exact pixel / index kernels (compiled with -fmad=false).  See image_kernels.h.
#include "image_kernels.h"

#include <climits>

#include "common.h"

namespace ocrs {
namespace img {

namespace {
constexpr int kThreads = 256;
inline unsigned grid1d(int64_t n, int threads = kThreads) { return (unsigned)ceil_div(n, threads); }

// ---------------------------------------------------------------------------------------------
// prepare_image (preprocess.rs:201-248): out = ((-0.5 + c0*w0) + c1*w1) + c2*w2
// ---------------------------------------------------------------------------------------------
struct PrepWeights {
  float w[3];
  int n;
};

template <typename T, bool kHwc>
__global__ void prepare_image_kernel(const T* __restrict__ px, float* __restrict__ out, int64_t hw, int C,
                                     PrepWeights pw) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  float v = kBlackValue;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (c < pw.n) {
      float s = kHwc ? (float)px[i * C + c] : (float)px[(int64_t)c * hw + i];
      v += s * pw.w[c];
    }
  }
  out[i] = v;
}

// Fast path: u8 HWC RGB, 4 pixels (12 bytes) per thread, float4 store.
__global__ void prepare_image_rgb8_x4_kernel(const uint32_t* __restrict__ px, float4* __restrict__ out,
                                             int64_t quads, PrepWeights pw) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= quads) return;
  uint32_t a = px[q * 3 + 0], b = px[q * 3 + 1], c = px[q * 3 + 2];
  // bytes: a = R0 G0 B0 R1 | b = G1 B1 R2 G2 | c = B2 R3 G3 B3   (little endian)
  float r0 = (float)(a & 0xFF), g0 = (float)((a >> 8) & 0xFF), b0 = (float)((a >> 16) & 0xFF);
  float r1 = (float)(a >> 24), g1 = (float)(b & 0xFF), b1 = (float)((b >> 8) & 0xFF);
  float r2 = (float)((b >> 16) & 0xFF), g2 = (float)(b >> 24), b2 = (float)(c & 0xFF);
  float r3 = (float)((c >> 8) & 0xFF), g3 = (float)((c >> 16) & 0xFF), b3 = (float)(c >> 24);
  float4 o;
  o.x = ((kBlackValue + r0 * pw.w[0]) + g0 * pw.w[1]) + b0 * pw.w[2];
  o.y = ((kBlackValue + r1 * pw.w[0]) + g1 * pw.w[1]) + b1 * pw.w[2];
  o.z = ((kBlackValue + r2 * pw.w[0]) + g2 * pw.w[1]) + b2 * pw.w[2];
  o.w = ((kBlackValue + r3 * pw.w[0]) + g3 * pw.w[1]) + b3 * pw.w[2];
  out[q] = o;
}

// ---------------------------------------------------------------------------------------------
// Bilinear resize, half-pixel centres (rten resize_image; oracle/imageops.py:resize_bilinear)
// ---------------------------------------------------------------------------------------------
struct AxisTap {
  int i0, i1;
  float w;
};
__device__ __forceinline__ AxisTap axis_tap(int d, int n_in, int n_out) {
  float scale = (float)n_in / (float)n_out;
  float src = scale * ((float)d + 0.5f) - 0.5f;
  src = fminf(fmaxf(src, 0.0f), (float)(n_in - 1));
  AxisTap t;
  t.i0 = (int)src;
  t.i1 = min(t.i0 + 1, n_in - 1);
  t.w = src - (float)t.i0;
  return t;
}
__device__ __forceinline__ float lerp2(float tl, float tr, float bl, float br, float wx, float wy) {
  float top = (1.0f - wx) * tl + wx * tr;
  float bot = (1.0f - wx) * bl + wx * br;
  return (1.0f - wy) * top + wy * bot;
}

__global__ void resize_padded_kernel(const float* __restrict__ src, int H, int W, int padH, int padW,
                                     float pad_value, float* __restrict__ dst, int OH, int OW, int64_t src_stride,
                                     int64_t dst_stride) {
  int ox = blockIdx.x * blockDim.x + threadIdx.x;
  int oy = blockIdx.y;
  int img = blockIdx.z;
  if (ox >= OW) return;
  const float* s = src + (int64_t)img * src_stride;
  AxisTap ty = axis_tap(oy, padH, OH);
  AxisTap tx = axis_tap(ox, padW, OW);
  auto at = [&](int y, int x) -> float { return (y < H && x < W) ? s[(int64_t)y * W + x] : pad_value; };
  float v = lerp2(at(ty.i0, tx.i0), at(ty.i0, tx.i1), at(ty.i1, tx.i0), at(ty.i1, tx.i1), tx.w, ty.w);
  dst[(int64_t)img * dst_stride + (int64_t)oy * OW + ox] = v;
}

__global__ void resize_threshold_kernel(const float* __restrict__ net, int inW, int sliceH, int sliceW,
                                        float* __restrict__ prob, uint8_t* __restrict__ mask, int H, int W,
                                        float thr) {
  int ox = blockIdx.x * blockDim.x + threadIdx.x;
  int oy = blockIdx.y;
  if (ox >= W) return;
  AxisTap ty = axis_tap(oy, sliceH, H);
  AxisTap tx = axis_tap(ox, sliceW, W);
  const float* r0 = net + (int64_t)ty.i0 * inW;
  const float* r1 = net + (int64_t)ty.i1 * inW;
  float v = lerp2(r0[tx.i0], r0[tx.i1], r1[tx.i0], r1[tx.i1], tx.w, ty.w);
  int64_t o = (int64_t)oy * W + ox;
  if (prob) prob[o] = v;
  mask[o] = v > thr ? 1 : 0;
}

// ---- batched over pages: one launch for the whole detection batch (page = blockIdx.z) ----
__global__ void resize_padded_batch_kernel(const PageResizeIn* __restrict__ tab, float pad_value, float* __restrict__ dst,
                                           int OH, int OW, int64_t dst_stride) {
  const PageResizeIn pg = tab[blockIdx.z];
  int ox = blockIdx.x * blockDim.x + threadIdx.x;
  int oy = blockIdx.y;
  if (ox >= OW) return;
  const float* s = pg.src;
  const int H = pg.H, W = pg.W;
  AxisTap ty = axis_tap(oy, pg.padH, OH);
  AxisTap tx = axis_tap(ox, pg.padW, OW);
  auto at = [&](int y, int x) -> float { return (y < H && x < W) ? __ldg(s + (int64_t)y * W + x) : pad_value; };
  float v = lerp2(at(ty.i0, tx.i0), at(ty.i0, tx.i1), at(ty.i1, tx.i0), at(ty.i1, tx.i1), tx.w, ty.w);
  dst[(int64_t)blockIdx.z * dst_stride + (int64_t)oy * OW + ox] = v;
}

__global__ void resize_threshold_batch_kernel(const PageResizeOut* __restrict__ tab, int inW, float thr) {
  const PageResizeOut pg = tab[blockIdx.z];
  int ox = blockIdx.x * blockDim.x + threadIdx.x;
  int oy = blockIdx.y;
  if (ox >= pg.W || oy >= pg.H) return;
  AxisTap ty = axis_tap(oy, pg.sliceH, pg.H);
  AxisTap tx = axis_tap(ox, pg.sliceW, pg.W);
  const float* r0 = pg.net + (int64_t)ty.i0 * inW;
  const float* r1 = pg.net + (int64_t)ty.i1 * inW;
  float v = lerp2(__ldg(r0 + tx.i0), __ldg(r0 + tx.i1), __ldg(r1 + tx.i0), __ldg(r1 + tx.i1), tx.w, ty.w);
  int64_t o = (int64_t)oy * pg.W + ox;
  if (pg.prob) pg.prob[o] = v;
  pg.mask[o] = v > thr ? 1 : 0;
}

// u8 HWC RGB pages of one shape, 4 pixels (12 bytes) per thread, page = blockIdx.y
__global__ void prepare_image_rgb8_x4_batch_kernel(const PagePrepare* __restrict__ tab, int64_t quads, PrepWeights pw) {
  const PagePrepare pg = tab[blockIdx.y];
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= quads) return;
  const uint32_t* px = reinterpret_cast<const uint32_t*>(pg.src);
  uint32_t a = __ldg(px + q * 3 + 0), b = __ldg(px + q * 3 + 1), c = __ldg(px + q * 3 + 2);
  float r0 = (float)(a & 0xFF), g0 = (float)((a >> 8) & 0xFF), b0 = (float)((a >> 16) & 0xFF);
  float r1 = (float)(a >> 24), g1 = (float)(b & 0xFF), b1 = (float)((b >> 8) & 0xFF);
  float r2 = (float)((b >> 16) & 0xFF), g2 = (float)(b >> 24), b2 = (float)(c & 0xFF);
  float r3 = (float)((c >> 8) & 0xFF), g3 = (float)((c >> 16) & 0xFF), b3 = (float)(c >> 24);
  float4 o;
  o.x = ((kBlackValue + r0 * pw.w[0]) + g0 * pw.w[1]) + b0 * pw.w[2];
  o.y = ((kBlackValue + r1 * pw.w[0]) + g1 * pw.w[1]) + b1 * pw.w[2];
  o.z = ((kBlackValue + r2 * pw.w[0]) + g2 * pw.w[1]) + b2 * pw.w[2];
  o.w = ((kBlackValue + r3 * pw.w[0]) + g3 * pw.w[1]) + b3 * pw.w[2];
  reinterpret_cast<float4*>(pg.dst)[q] = o;
}

__global__ void threshold_kernel(const float* __restrict__ p, uint8_t* __restrict__ m, int64_t n, float thr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) m[i] = p[i] > thr ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Connected components: union-find with atomicMin (root = smallest pixel index of the set).
// Foreground is 8-connected, background 4-connected (so that holes are what Suzuki-Abe's
// outermost-border rule sees); background pixels on the image border are united with the
// virtual frame node at index H*W.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int32_t* L, int i) {
  int p = L[i];
  while (p != i) {
    i = p;
    p = L[i];
  }
  return i;
}
__device__ void uf_union(int32_t* L, int a, int b) {
  bool done = false;
  while (!done) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) {
      int old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  }
}

__global__ void ccl_frame_init_kernel(int32_t* L, int64_t n) { L[n] = (int32_t)n; }

// ---- two-level labelling: tile-local union-find in shared memory, then unions across tile borders ----
constexpr int kTile = 32;

__device__ __forceinline__ int suf_find(const int* L, int i) {
  int p = L[i];
  while (p != i) {
    i = p;
    p = L[i];
  }
  return i;
}
__device__ __forceinline__ void suf_union(int* L, int a, int b) {
  bool done = false;
  while (!done) {
    a = suf_find(L, a);
    b = suf_find(L, b);
    if (a < b) {
      int old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  }
}

__global__ void __launch_bounds__(256) ccl_tile_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ L, int H,
                                                       int W) {
  __shared__ int sl[kTile * kTile];
  __shared__ uint8_t sm[kTile * kTile];
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;  // 8 rows per pass
  for (int ty = ty0; ty < kTile; ty += 8) {
    int x = x0 + tx, y = y0 + ty, li = ty * kTile + tx;
    sl[li] = li;
    sm[li] = (x < W && y < H) ? (mask[(int64_t)y * W + x] ? 1 : 0) : 2;  // 2 = outside the image
  }
  __syncthreads();
  for (int ty = ty0; ty < kTile; ty += 8) {
    int li = ty * kTile + tx;
    uint8_t v = sm[li];
    if (v == 2) continue;
    if (v) {
      if (tx > 0 && sm[li - 1] == 1) suf_union(sl, li, li - 1);
      if (ty > 0) {
        if (sm[li - kTile] == 1) suf_union(sl, li, li - kTile);
        if (tx > 0 && sm[li - kTile - 1] == 1) suf_union(sl, li, li - kTile - 1);
        if (tx + 1 < kTile && sm[li - kTile + 1] == 1) suf_union(sl, li, li - kTile + 1);
      }
    } else {
      if (tx > 0 && sm[li - 1] == 0) suf_union(sl, li, li - 1);
      if (ty > 0 && sm[li - kTile] == 0) suf_union(sl, li, li - kTile);
    }
  }
  __syncthreads();
  for (int ty = ty0; ty < kTile; ty += 8) {
    int x = x0 + tx, y = y0 + ty, li = ty * kTile + tx;
    if (x >= W || y >= H) continue;
    int r = suf_find(sl, li);
    L[(int64_t)y * W + x] = (y0 + r / kTile) * W + x0 + (r % kTile);  // global index of the tile-local root
  }
}

// unions across tile borders + the virtual frame node (index H*W) for background on the image border
__global__ void ccl_border_kernel(const uint8_t* __restrict__ mask, int32_t* L, int H, int W) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= W) return;
  const bool on_v = (x % kTile) == 0, on_h = (y % kTile) == 0;
  const bool img_border = (x == 0 || y == 0 || x == W - 1 || y == H - 1);
  const bool right_edge = (x % kTile) == kTile - 1;  // NE neighbour lies in the next tile column
  if (!on_v && !on_h && !img_border && !right_edge) return;
  int p = y * W + x;
  uint8_t v = mask[p];
  if (v) {
    if (on_v && x > 0 && mask[p - 1]) uf_union(L, p, p - 1);
    if (y > 0) {
      if (on_h && mask[p - W]) uf_union(L, p, p - W);
      if ((on_h || on_v) && x > 0 && mask[p - W - 1]) uf_union(L, p, p - W - 1);
      if ((on_h || right_edge) && x + 1 < W && mask[p - W + 1]) uf_union(L, p, p - W + 1);
    }
  } else {
    if (on_v && x > 0 && !mask[p - 1]) uf_union(L, p, p - 1);
    if (on_h && y > 0 && !mask[p - W]) uf_union(L, p, p - W);
    if (img_border) uf_union(L, p, H * W);
  }
}

__global__ void ccl_flatten_kernel(const uint8_t* __restrict__ mask, int32_t* L, int64_t n, int32_t* comp_roots,
                                   int32_t* counters, int max_comps) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int r = uf_find(L, (int)i);
  L[i] = r;
  if (i < n && r == (int)i && mask[i]) {
    int slot = atomicAdd(&counters[0], 1);
    if (slot < max_comps) comp_roots[slot] = (int)i;
    else atomicExch(&counters[2], 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Per component: Suzuki-Abe outer border following -> RDP -> hull -> min-area rect.
// ---------------------------------------------------------------------------------------------
__constant__ int kDy[8] = {0, 1, 1, 1, 0, -1, -1, -1};  // clockwise from E (image coords)
__constant__ int kDx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
__constant__ int kDirOf[9] = {5, 6, 7, 4, -1, 0, 3, 2, 1};  // [(dy+1)*3 + (dx+1)]

struct MaskView {
  const uint8_t* m;
  int H, W;
  __device__ __forceinline__ int at(int y, int x) const {
    return (y >= 0 && y < H && x >= 0 && x < W) ? m[y * W + x] : 0;
  }
};

// Follows the outer border that starts at (i, j) (the component's first pixel in raster
// order).  When `out` is non-null writes (x, y) pairs.  Returns the number of points, or -1 if
// `limit` would be exceeded.
__device__ int trace_border(const MaskView& mv, int i, int j, int16_t* out, int64_t limit) {
  int i1 = 0, j1 = 0;
  bool found = false;
  for (int k = 0; k < 8; ++k) {  // (3.1) clockwise from W around (i, j)
    int d = (4 + k) & 7;
    if (mv.at(i + kDy[d], j + kDx[d])) {
      i1 = i + kDy[d];
      j1 = j + kDx[d];
      found = true;
      break;
    }
  }
  if (!found) {  // isolated pixel
    if (limit < 1) return -1;
    if (out) { out[0] = (int16_t)j; out[1] = (int16_t)i; }
    return 1;
  }
  int i2 = i1, j2 = j1, i3 = i, j3 = j;
  int n = 0;
  while (true) {
    int d0 = kDirOf[(i2 - i3 + 1) * 3 + (j2 - j3 + 1)];
    int i4 = i3, j4 = j3;
    for (int k = 1; k <= 8; ++k) {
      int d = (d0 - k) & 7;
      if (mv.at(i3 + kDy[d], j3 + kDx[d])) {
        i4 = i3 + kDy[d];
        j4 = j3 + kDx[d];
        break;
      }
    }
    if (n >= limit) return -1;
    if (out) { out[2 * n] = (int16_t)j3; out[2 * n + 1] = (int16_t)i3; }
    ++n;
    if (i4 == i && j4 == j && i3 == i1 && j3 == j1) break;
    i2 = i3; j2 = j3; i3 = i4; j3 = j4;
  }
  return n;
}

__global__ void component_rects_kernel(const uint8_t* __restrict__ mask, int H, int W, float eps, float expand,
                                       float min_area, ComponentBuffers b) {
  int ci = blockIdx.x * blockDim.x + threadIdx.x;
  int n_comps = min(b.counters[0], b.max_comps);
  if (ci >= n_comps) return;
  int root = b.comp_roots[ci];
  int i = root / W, j = root - i * W;
  // outermost-border rule (Suzuki-Abe App. II): the 0-pixel left of the first pixel must belong
  // to the background component that touches the frame.
  if (j > 0 && b.labels[root - 1] != b.labels[(int64_t)H * W]) return;
  MaskView mv{mask, H, W};
  int n = trace_border(mv, i, j, nullptr, b.pool_cap);
  if (n < 0) { atomicExch(&b.counters[2], 2); return; }
  int64_t off = (int64_t)atomicAdd((unsigned long long*)(void*)&b.counters[4], (unsigned long long)(n + 2));
  if (off + n + 2 > b.pool_cap) { atomicExch(&b.counters[2], 2); return; }
  int16_t* pts = b.pts + 2 * off;
  trace_border(mv, i, j, pts, n);

  // ---- RDP on the closed polyline P[0..n], P[n] == P[0] (simplify_polygon) ----
  int32_t* simp = b.simp_idx + off;
  int32_t* stack = b.stack + 3 * off;
  auto P = [&](int idx) -> geom::PointF {
    if (idx == n) idx = 0;
    return geom::PointF{(float)pts[2 * idx], (float)pts[2 * idx + 1]};
  };
  int sp = 0, m = 0;
  stack[0] = 0; stack[1] = n; stack[2] = 1;
  sp = 1;
  while (sp > 0) {
    --sp;
    int s = stack[3 * sp], e = stack[3 * sp + 1], keep = stack[3 * sp + 2];
    if (e - s + 1 <= 1) {
      simp[m++] = s;
      continue;
    }
    geom::LineF seg{P(s), P(e)};
    int max_i = s;  // "0" relative to the slice
    float max_d = 0.0f;
    for (int k = s + 1; k < e; ++k) {
      float d = geom::line_distance(seg, P(k));
      if (d >= max_d) { max_i = k; max_d = d; }
    }
    if (max_d > eps) {
      // left half first: push right, then left
      stack[3 * sp] = max_i; stack[3 * sp + 1] = e; stack[3 * sp + 2] = keep; ++sp;
      stack[3 * sp] = s; stack[3 * sp + 1] = max_i; stack[3 * sp + 2] = 0; ++sp;
    } else {
      simp[m++] = s;
      if (keep) simp[m++] = e;
    }
  }
  m -= 1;  // drop the duplicated closing point
  if (m < 1) return;
  float* fp = b.fpts + 2 * off;
  for (int k = 0; k < m; ++k) {
    geom::PointF q = P(simp[k]);
    fp[2 * k] = q.x;
    fp[2 * k + 1] = q.y;
  }
  geom::PointF* fpts = reinterpret_cast<geom::PointF*>(fp);
  geom::PointF* hull = reinterpret_cast<geom::PointF*>(b.hull + 2 * off);
  int hm = geom::convex_hull(fpts, m, hull);
  geom::RotatedRect rr;
  if (!geom::min_area_rect_of_hull(hull, hm, &rr)) return;
  rr.w = rr.w + 2.0f * expand;  // detection.rs:53-57
  rr.h = rr.h + 2.0f * expand;
  if (!(geom::rr_area(rr) >= min_area)) return;  // detection.rs:60
  int slot = atomicAdd(&b.counters[3], 1);
  b.rects[slot] = rr;
  b.rect_root[slot] = root;
}

// ---------------------------------------------------------------------------------------------
// Line crops (recognition.rs:91-158)
// ---------------------------------------------------------------------------------------------
__global__ void line_crossings_kernel(const LineDesc* __restrict__ lines, int n_lines,
                                      const int32_t* __restrict__ poly_xy, int32_t* __restrict__ cross) {
  int li = blockIdx.y;
  if (li >= n_lines) return;
  const LineDesc L = lines[li];
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= L.lh) return;
  int y = L.top + r;
  int32_t* row = cross + L.cross_off + (int64_t)r * (L.max_cross + 1);
  const int32_t* v = poly_xy + 2 * (int64_t)L.poly_off;
  int cnt = 0;
  for (int k = 0; k < L.poly_n; ++k) {
    int k2 = (k + 1 == L.poly_n) ? 0 : k + 1;
    int sx = v[2 * k], sy = v[2 * k + 1], ex = v[2 * k2], ey = v[2 * k2 + 1];
    if (sy == ey) continue;
    if (sy > ey) { int t = sx; sx = ex; ex = t; t = sy; sy = ey; ey = t; }  // downwards
    if (!(sy <= y && y < ey)) continue;
    geom::LineF lf{geom::PointF{(float)sx, (float)sy}, geom::PointF{(float)ex, (float)ey}};
    float xf = 0.0f;
    geom::line_x_for_y(lf, (float)y, &xf);
    if (cnt < L.max_cross) row[1 + cnt] = geom::f2i(roundf(xf));
    ++cnt;
  }
  row[0] = cnt;
}

__global__ void crop_resize_kernel(const float* const* __restrict__ pages, const int* __restrict__ page_h,
                                   const int* __restrict__ page_w, const LineDesc* __restrict__ lines, int n_lines,
                                   const int32_t* __restrict__ cross, float* __restrict__ dst, int out_h) {
  int li = blockIdx.z;
  if (li >= n_lines) return;
  const LineDesc L = lines[li];
  int ox = blockIdx.x * blockDim.x + threadIdx.x;
  int oy = blockIdx.y;
  if (ox >= L.group_width) return;
  float v = kBlackValue;
  if (ox < L.resized_width && L.lh > 0 && L.lw > 0) {
    const float* page = pages[L.page];
    const int PH = page_h[L.page], PW = page_w[L.page];
    AxisTap ty = axis_tap(oy, L.lh, out_h);
    AxisTap tx = axis_tap(ox, L.lw, L.resized_width);
    auto canvas = [&](int cy, int cx) -> float {
      int py = cy + L.top, px = cx + L.left;
      // both points are tested against the page index rect (recognition.rs:112)
      if (py < 0 || py > PH - 1 || px < 0 || px > PW - 1) return kBlackValue;
      if (cy < 0 || cy > PH - 1 || cx < 0 || cx > PW - 1) return kBlackValue;
      const int32_t* row = cross + L.cross_off + (int64_t)cy * (L.max_cross + 1);
      int cnt = row[0], inside = 0;
      for (int k = 0; k < cnt; ++k) inside += (row[1 + k] <= px) ? 1 : 0;
      return (inside & 1) ? page[(int64_t)py * PW + px] : kBlackValue;
    };
    v = lerp2(canvas(ty.i0, tx.i0), canvas(ty.i0, tx.i1), canvas(ty.i1, tx.i0), canvas(ty.i1, tx.i1), tx.w, ty.w);
  }
  dst[L.dst_off + (int64_t)oy * L.group_width + ox] = v;
}

// ---------------------------------------------------------------------------------------------
// CTC greedy: warp per (t, b) argmax (first maximum), then one thread per line collapses.
// ---------------------------------------------------------------------------------------------
__global__ void ctc_argmax_kernel(const float* __restrict__ logits, int T, int B, int C,
                                  const uint8_t* __restrict__ excluded, int32_t* __restrict__ labels) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x & 31;
  if (warp >= (int64_t)T * B) return;
  int t = warp / B, b = warp - (int64_t)t * B;
  const float* row = logits + warp * C;
  float best = -INFINITY;
  int bi = INT_MAX;
  for (int c = lane; c < C; c += 32) {
    float v = (excluded && excluded[c]) ? -INFINITY : row[c];
    if (bi == INT_MAX || v > best) { best = v; bi = c; }
  }
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != INT_MAX && (bi == INT_MAX || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  if (lane == 0) labels[(int64_t)b * T + t] = bi;
}

__global__ void ctc_collapse_kernel(const int32_t* __restrict__ labels, int T, int B, int32_t* __restrict__ out_l,
                                    int32_t* __restrict__ out_p, int32_t* __restrict__ counts) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int32_t* l = labels + (int64_t)b * T;
  int last = 0, n = 0;
  for (int pos = 0; pos < T; ++pos) {
    int v = l[pos];
    if (v == last) continue;
    last = v;
    if (v > 0) {
      out_l[(int64_t)b * T + n] = v;
      out_p[(int64_t)b * T + n] = pos;
      ++n;
    }
  }
  counts[b] = n;
}

__global__ void ctc_argmax_rows_kernel(const float* __restrict__ logits, int64_t rows, int C,
                                       const uint8_t* __restrict__ excluded, int32_t* __restrict__ labels) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* row = logits + warp * C;
  float best = -INFINITY;
  int bi = INT_MAX;
  for (int c = lane; c < C; c += 32) {
    float v = (excluded && excluded[c]) ? -INFINITY : row[c];
    if (bi == INT_MAX || v > best) { best = v; bi = c; }
  }
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != INT_MAX && (bi == INT_MAX || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  if (lane == 0) labels[warp] = bi;
}

// one warp per line: a label is emitted where it differs from its predecessor and is not blank (equivalent to
// the sequential `last` rule of decode_greedy); order-preserving compaction through ballots
__global__ void ctc_collapse_lines_kernel(const int32_t* __restrict__ labels, const CtcLine* __restrict__ lines,
                                          int n_lines, int32_t* __restrict__ out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_lines) return;
  const CtcLine L = lines[i];
  int n = 0, carry = 0;  // carry = label at the last position of the previous chunk (0 before the line starts)
  for (int p0 = 0; p0 < L.T; p0 += 32) {
    const int pos = p0 + lane;
    const int v = pos < L.T ? labels[L.base + (int64_t)pos * L.stride] : 0;
    int prev = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) prev = carry;
    const bool emit = pos < L.T && v != prev && v > 0;
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    if (emit) {
      const int k = n + __popc(m & ((1u << lane) - 1u));
      out[L.lab_off + k] = v;
      out[L.pos_off + k] = pos;
    }
    n += __popc(m);
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) out[L.cnt_off] = n;
}

}  // namespace

// ================================== host launchers ===========================================
void prepare_image(const void* pixels, int dtype, int order, int H, int W, int C, float* out, cudaStream_t st) {
  int64_t hw = (int64_t)H * W;
  if (hw == 0) return;
  const float itu[3] = {0.299f, 0.587f, 0.114f};  // preprocess.rs:171
  PrepWeights pw;
  pw.n = (C == 1) ? 1 : 3;
  for (int c = 0; c < 3; ++c) {
    if (dtype == 0) pw.w[c] = (C == 1) ? (1.0f / 255.0f) : (itu[c] / 255.0f);  // preprocess.rs:182,184
    else pw.w[c] = (C == 1) ? 1.0f : itu[c];
  }
  if (dtype == 0) {
    const uint8_t* p = static_cast<const uint8_t*>(pixels);
    if (order == 0 && C == 3 && hw % 4 == 0 && (reinterpret_cast<uintptr_t>(p) % 4 == 0) &&
        (reinterpret_cast<uintptr_t>(out) % 16 == 0)) {
      prepare_image_rgb8_x4_kernel<<<grid1d(hw / 4), kThreads, 0, st>>>(reinterpret_cast<const uint32_t*>(p),
                                                                         reinterpret_cast<float4*>(out), hw / 4, pw);
  count_launch();
    } else if (order == 0) {
      prepare_image_kernel<uint8_t, true><<<grid1d(hw), kThreads, 0, st>>>(p, out, hw, C, pw);
  count_launch();
    } else {
      prepare_image_kernel<uint8_t, false><<<grid1d(hw), kThreads, 0, st>>>(p, out, hw, C, pw);
  count_launch();
    }
  } else {
    const float* p = static_cast<const float*>(pixels);
    if (order == 0) {
      prepare_image_kernel<float, true><<<grid1d(hw), kThreads, 0, st>>>(p, out, hw, C, pw);
    } else {
      prepare_image_kernel<float, false><<<grid1d(hw), kThreads, 0, st>>>(p, out, hw, C, pw);
    }
    count_launch();
  }
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void resize_padded(const float* src, int H, int W, int padH, int padW, float pad_value, float* dst, int OH, int OW,
                   int n, int64_t src_stride, int64_t dst_stride, cudaStream_t st) {
  if (OH == 0 || OW == 0 || n == 0) return;
  dim3 grid(grid1d(OW, 128), OH, n);
  resize_padded_kernel<<<grid, 128, 0, st>>>(src, H, W, padH, padW, pad_value, dst, OH, OW, src_stride, dst_stride);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void resize_threshold(const float* net_out, int inH, int inW, int sliceH, int sliceW, float* prob, uint8_t* mask,
                      int H, int W, float thr, cudaStream_t st) {
  (void)inH;
  if (H == 0 || W == 0) return;
  dim3 grid(grid1d(W, 128), H);
  resize_threshold_kernel<<<grid, 128, 0, st>>>(net_out, inW, sliceH, sliceW, prob, mask, H, W, thr);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void resize_padded_batch(const PageResizeIn* d_tab, int n, float pad_value, float* dst, int OH, int OW, int64_t dst_stride,
                         cudaStream_t st) {
  if (OH == 0 || OW == 0 || n == 0) return;
  dim3 grid(grid1d(OW, 128), OH, n);
  resize_padded_batch_kernel<<<grid, 128, 0, st>>>(d_tab, pad_value, dst, OH, OW, dst_stride);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void resize_threshold_batch(const PageResizeOut* d_tab, int n, int inW, int maxH, int maxW, float thr, cudaStream_t st) {
  if (maxH == 0 || maxW == 0 || n == 0) return;
  dim3 grid(grid1d(maxW, 128), maxH, n);
  resize_threshold_batch_kernel<<<grid, 128, 0, st>>>(d_tab, inW, thr);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

bool prepare_image_batchable(const void* pixels, int dtype, int order, int H, int W, int C, const float* out) {
  const int64_t hw = (int64_t)H * W;
  return dtype == 0 && order == 0 && C == 3 && hw > 0 && hw % 4 == 0 && (reinterpret_cast<uintptr_t>(pixels) % 4 == 0) &&
         (reinterpret_cast<uintptr_t>(out) % 16 == 0);
}

void prepare_image_rgb8_batch(const PagePrepare* d_tab, int n, int H, int W, cudaStream_t st) {
  const int64_t hw = (int64_t)H * W;
  if (hw == 0 || n == 0) return;
  const float itu[3] = {0.299f, 0.587f, 0.114f};  // preprocess.rs:171
  PrepWeights pw;
  pw.n = 3;
  for (int c = 0; c < 3; ++c) pw.w[c] = itu[c] / 255.0f;  // preprocess.rs:184
  dim3 grid(grid1d(hw / 4), (unsigned)n);
  prepare_image_rgb8_x4_batch_kernel<<<grid, kThreads, 0, st>>>(d_tab, hw / 4, pw);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void threshold(const float* prob, uint8_t* mask, int64_t n, float thr, cudaStream_t st) {
  if (!n) return;
  threshold_kernel<<<grid1d(n), kThreads, 0, st>>>(prob, mask, n, thr);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void find_component_rects(const uint8_t* mask, int H, int W, float eps, float expand_dist, float min_area,
                          const ComponentBuffers& b, cudaStream_t st) {
  int64_t n = (int64_t)H * W;
  OCRS_CUDA_CHECK(cudaMemsetAsync(b.counters, 0, 8 * sizeof(int32_t), st));
  if (n == 0) return;
  OCRS_CHECK(H < 32768 && W < 32768, kInvalidArg, "image dimensions exceed 32767");
  {
    dim3 tg((unsigned)ceil_div(W, kTile), (unsigned)ceil_div(H, kTile));
    ccl_tile_kernel<<<tg, 256, 0, st>>>(mask, b.labels, H, W);
    count_launch();
    ccl_frame_init_kernel<<<1, 1, 0, st>>>(b.labels, n);
    count_launch();
    dim3 grid(grid1d(W, 128), H);
    ccl_border_kernel<<<grid, 128, 0, st>>>(mask, b.labels, H, W);
    count_launch();
  }
  ccl_flatten_kernel<<<grid1d(n + 1), kThreads, 0, st>>>(mask, b.labels, n, b.comp_roots, b.counters, b.max_comps);
  count_launch();
  // one thread per component; the count is only known on the device, so launch for the
  // theoretical maximum in chunks guarded by counters[0] (cheap: threads beyond n_comps exit).
  int max_c = b.max_comps;
  component_rects_kernel<<<grid1d(max_c, 64), 64, 0, st>>>(mask, H, W, eps, expand_dist, min_area, b);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void crop_lines(const float* const* pages, const int* page_h, const int* page_w, const LineDesc* lines, int n_lines,
                const int32_t* poly_xy, int32_t* cross_scratch, float* dst, int out_h, int max_group_width,
                int max_rows, cudaStream_t st) {
  if (n_lines == 0) return;
  if (max_rows > 0) {
    dim3 g1(grid1d(max_rows, 64), n_lines);
    line_crossings_kernel<<<g1, 64, 0, st>>>(lines, n_lines, poly_xy, cross_scratch);
  count_launch();
  }
  dim3 g2(grid1d(max_group_width, 128), out_h, n_lines);
  crop_resize_kernel<<<g2, 128, 0, st>>>(pages, page_h, page_w, lines, n_lines, cross_scratch, dst, out_h);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void ctc_greedy(const float* logits, int T, int B, int C, const uint8_t* excluded, int32_t* scratch_labels,
                int32_t* labels_out, int32_t* pos_out, int32_t* counts_out, cudaStream_t st) {
  if (B == 0) return;
  if (T > 0) ctc_argmax_kernel<<<grid1d((int64_t)T * B * 32), kThreads, 0, st>>>(logits, T, B, C, excluded, scratch_labels);
  count_launch();
  ctc_collapse_kernel<<<grid1d(B, 64), 64, 0, st>>>(scratch_labels, T, B, labels_out, pos_out, counts_out);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void ctc_greedy_packed(const float* logits, int64_t rows, int C, const uint8_t* excluded, int32_t* row_labels,
                       const CtcLine* lines, int n_lines, int32_t* out, cudaStream_t st) {
  if (n_lines == 0) return;
  if (rows > 0) {
    ctc_argmax_rows_kernel<<<grid1d(rows * 32), kThreads, 0, st>>>(logits, rows, C, excluded, row_labels);
    count_launch();
  }
  ctc_collapse_lines_kernel<<<grid1d((int64_t)n_lines * 32, 128), 128, 0, st>>>(row_labels, lines, n_lines, out);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

}  // namespace img
}  // namespace ocrs
