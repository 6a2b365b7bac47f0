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
// same with the scale (float)n_in / (float)n_out computed once on the host (one IEEE division, same bits)
__device__ __forceinline__ AxisTap axis_tap_s(int d, int n_in, float scale) {
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
  AxisTap ty = axis_tap_s(oy, pg.padH, pg.sy);
  AxisTap tx = axis_tap_s(ox, pg.padW, pg.sx);
  auto at = [&](int y, int x) -> float { return (y < H && x < W) ? __ldg(s + (int64_t)y * W + x) : pad_value; };
  float v = lerp2(at(ty.i0, tx.i0), at(ty.i0, tx.i1), at(ty.i1, tx.i0), at(ty.i1, tx.i1), tx.w, ty.w);
  dst[(int64_t)blockIdx.z * dst_stride + (int64_t)oy * OW + ox] = v;
}

__global__ void resize_threshold_batch_kernel(const PageResizeOut* __restrict__ tab, int inW, float thr) {
  const PageResizeOut pg = tab[blockIdx.z];
  int ox = blockIdx.x * blockDim.x + threadIdx.x;  // a warp covers 32 consecutive x starting at a multiple of 32
  int oy = blockIdx.y;
  if (oy >= pg.H) return;
  const bool in = ox < pg.W;
  bool fg = false;
  if (in) {
    AxisTap ty = axis_tap_s(oy, pg.sliceH, pg.sy);
    AxisTap tx = axis_tap_s(ox, pg.sliceW, pg.sx);
    const float* r0 = pg.net + (int64_t)ty.i0 * inW;
    const float* r1 = pg.net + (int64_t)ty.i1 * inW;
    float v = lerp2(__ldg(r0 + tx.i0), __ldg(r0 + tx.i1), __ldg(r1 + tx.i0), __ldg(r1 + tx.i1), tx.w, ty.w);
    int64_t o = (int64_t)oy * pg.W + ox;
    if (pg.prob) pg.prob[o] = v;
    fg = v > thr;
    pg.mask[o] = fg ? 1 : 0;
  }
  // bit-packed copy of the mask (bit x & 31 of word x >> 5; bits beyond W are 0) for labelling / contours
  const unsigned word = __ballot_sync(0xffffffffu, fg);
  const int wd = (pg.W + 31) >> 5;
  if ((threadIdx.x & 31) == 0 && pg.bits != nullptr && (ox >> 5) < wd) pg.bits[(int64_t)oy * wd + (ox >> 5)] = word;
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
// Connected components on the bit-packed mask, run based.
// Foreground is 8-connected, background 4-connected (so that holes are what Suzuki-Abe's
// outermost-border rule sees); background pixels on the image border are united with a virtual
// frame node at index H*W.  Nodes of the union-find are the START pixels of horizontal runs of
// equal value (index y*W + x; root = smallest index of the set, so the root of a foreground
// component is its first pixel in raster order).  One warp per row, lane = 32-pixel word:
//   ccl_init   : parent[start] = start for every run start; wstart[y][w] = start x of the run that
//                covers bit 0 of word w (warp scan over "uniform" words);
//   ccl_union  : for the three 8-neighbour offsets d (foreground) / d = 0 (background) the pair masks
//                P_d = cur & shift(up, d); one union per maximal run of P_d (a run of pairs joins the
//                same two runs) -- no per-pixel work, no shared-memory atomics;
//   ccl_flatten: parent[start] = root; foreground roots are appended to the component list.
// All pages of a batch run in one launch each (blockIdx.y = page).
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

// start x of the run (of equal mask value) that contains pixel x of row y
__device__ __forceinline__ int run_start(const uint32_t* __restrict__ bits, const uint16_t* __restrict__ wstart, int wd,
                                         int y, int x) {
  const int w = x >> 5, b = x & 31;
  const uint32_t word = bits[(int64_t)y * wd + w];
  const uint32_t same = ((word >> b) & 1u) ? word : ~word;  // bits equal to pixel x's value
  const uint32_t z = ~same & ((1u << b) - 1u);              // different-valued bits below b
  if (z) return (w << 5) + (32 - __clz(z));
  return wstart[(int64_t)y * wd + w];
}

constexpr int kCclWarps = 4;

__global__ void __launch_bounds__(kCclWarps * 32) ccl_init_kernel(const CclPage* __restrict__ pages) {
  const CclPage pg = pages[blockIdx.y];
  const int y = blockIdx.x * kCclWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) pg.labels[(int64_t)pg.H * pg.W] = pg.H * pg.W;  // frame node
  if (y >= pg.H) return;
  const int wd = pg.wd, W = pg.W;
  uint32_t prev_last = 0;       // last pixel value of the previous chunk's last word
  int carry_start = 0;          // start of the run reaching the end of the previous chunk
  for (int w0 = 0; w0 < wd; w0 += 32) {
    const int w = w0 + lane;
    const uint32_t m = w < wd ? pg.bits[(int64_t)y * wd + w] : 0u;
    uint32_t p31 = __shfl_up_sync(0xffffffffu, m >> 31, 1);
    if (lane == 0) p31 = w0 == 0 ? ((~m) & 1u) : prev_last;  // the first pixel of a row always starts a run
    const uint32_t c = m ^ ((m << 1) | p31);                  // bit x: value differs from pixel x-1 -> run start
    if (w < wd) {
      uint32_t cc = c;
      while (cc) {
        const int bpos = __ffs(cc) - 1;
        cc &= cc - 1;
        const int x = (w << 5) + bpos;
        if (x < W) pg.labels[(int64_t)y * W + x] = y * W + x;
      }
    }
    // start of the run that covers the LAST bit of each word; words without a change are transparent
    const int last_start = c ? (w << 5) + (31 - __clz(c)) : 0;
    const unsigned nt = __ballot_sync(0xffffffffu, c != 0u && w < wd);
    const unsigned below = nt & (0xffffffffu >> (31 - lane));  // non-transparent lanes <= lane
    const int src = below ? 31 - __clz(below) : 0;
    int resolved = __shfl_sync(0xffffffffu, last_start, src);
    if (!below) resolved = carry_start;
    int prev_resolved = __shfl_up_sync(0xffffffffu, resolved, 1);
    if (lane == 0) prev_resolved = carry_start;
    if (w < wd) pg.wstart[(int64_t)y * wd + w] = (uint16_t)((c & 1u) ? (w << 5) : prev_resolved);
    carry_start = __shfl_sync(0xffffffffu, resolved, 31);
    prev_last = __shfl_sync(0xffffffffu, m >> 31, 31);
  }
}

__global__ void __launch_bounds__(kCclWarps * 32) ccl_union_kernel(const CclPage* __restrict__ pages) {
  const CclPage pg = pages[blockIdx.y];
  const int y = blockIdx.x * kCclWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y >= pg.H) return;
  const int wd = pg.wd, W = pg.W, H = pg.H;
  const int frame = H * W;
  int32_t* L = pg.labels;
  const uint32_t* bits = pg.bits;
  const uint16_t* ws = pg.wstart;
  // carries between chunks of 32 words: last bit of the pair masks of the previous chunk's last word
  uint32_t cm = 0, cu = 0, cP0 = 0, cPm = 0, cPp = 0, cQ = 0;
  for (int w0 = 0; w0 < wd; w0 += 32) {
    const int w = w0 + lane;
    const bool live = w < wd;
    const uint32_t m = live ? bits[(int64_t)y * wd + w] : 0u;
    const uint32_t u = (live && y > 0) ? bits[(int64_t)(y - 1) * wd + w] : 0u;
    uint32_t valid = 0u;
    if (live) valid = (W - (w << 5) >= 32) ? 0xffffffffu : ((1u << (W - (w << 5))) - 1u);
    // neighbour bits of the upper row across word boundaries
    uint32_t u_prev31 = __shfl_up_sync(0xffffffffu, u >> 31, 1);
    if (lane == 0) u_prev31 = cu;
    uint32_t u_next0 = __shfl_down_sync(0xffffffffu, u & 1u, 1);
    if (lane == 31) u_next0 = (w + 1 < wd && y > 0) ? (bits[(int64_t)(y - 1) * wd + w + 1] & 1u) : 0u;
    const uint32_t up_m1 = (u << 1) | u_prev31;          // upper pixel at x - 1
    const uint32_t up_p1 = (u >> 1) | (u_next0 << 31);   // upper pixel at x + 1
    const uint32_t P0 = m & u, Pm = m & up_m1, Pp = m & up_p1;
    const uint32_t bgm = ~m & valid, bgu = (y > 0) ? (~u & valid) : 0u;
    const uint32_t Q = bgm & bgu;
    auto prev31 = [&](uint32_t v, uint32_t carry) {
      uint32_t p = __shfl_up_sync(0xffffffffu, v >> 31, 1);
      return lane == 0 ? carry : p;
    };
    const uint32_t E0 = P0 & ~((P0 << 1) | prev31(P0, cP0));
    const uint32_t Em = Pm & ~((Pm << 1) | prev31(Pm, cPm));
    const uint32_t Ep = Pp & ~((Pp << 1) | prev31(Pp, cPp));
    const uint32_t EQ = Q & ~((Q << 1) | prev31(Q, cQ));
    if (live && y > 0) {
      for (int d = -1; d <= 1; ++d) {
        uint32_t e = d == 0 ? E0 : (d < 0 ? Em : Ep);
        while (e) {
          const int bpos = __ffs(e) - 1;
          e &= e - 1;
          const int x = (w << 5) + bpos;
          uf_union(L, y * W + run_start(bits, ws, wd, y, x), (y - 1) * W + run_start(bits, ws, wd, y - 1, x + d));
        }
      }
      uint32_t e = EQ;
      while (e) {
        const int bpos = __ffs(e) - 1;
        e &= e - 1;
        const int x = (w << 5) + bpos;
        uf_union(L, y * W + run_start(bits, ws, wd, y, x), (y - 1) * W + run_start(bits, ws, wd, y - 1, x));
      }
    }
    // background on the image border joins the frame node
    if (live) {
      uint32_t border = 0u;
      if (w == 0) border |= bgm & 1u;
      if (w == ((W - 1) >> 5)) border |= bgm & (1u << ((W - 1) & 31));
      while (border) {
        const int bpos = __ffs(border) - 1;
        border &= border - 1;
        uf_union(L, y * W + run_start(bits, ws, wd, y, (w << 5) + bpos), frame);
      }
    }
    if (y == 0 || y == H - 1) {
      // run starts of background in this word (value changes where the new value is background)
      uint32_t p31 = __shfl_up_sync(0xffffffffu, m >> 31, 1);
      if (lane == 0) p31 = w0 == 0 ? ((~m) & 1u) : cm;
      uint32_t sb = (m ^ ((m << 1) | p31)) & bgm;
      while (live && sb) {
        const int bpos = __ffs(sb) - 1;
        sb &= sb - 1;
        uf_union(L, y * W + (w << 5) + bpos, frame);
      }
    }
    cm = __shfl_sync(0xffffffffu, m >> 31, 31);
    cu = __shfl_sync(0xffffffffu, u >> 31, 31);
    cP0 = __shfl_sync(0xffffffffu, P0 >> 31, 31);
    cPm = __shfl_sync(0xffffffffu, Pm >> 31, 31);
    cPp = __shfl_sync(0xffffffffu, Pp >> 31, 31);
    cQ = __shfl_sync(0xffffffffu, Q >> 31, 31);
  }
}

__global__ void __launch_bounds__(kCclWarps * 32) ccl_flatten_kernel(const CclPage* __restrict__ pages) {
  const CclPage pg = pages[blockIdx.y];
  const int y = blockIdx.x * kCclWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int W = pg.W, wd = pg.wd;
  int32_t* L = pg.labels;
  if (blockIdx.x == 0 && threadIdx.x == 0) L[(int64_t)pg.H * W] = uf_find(L, pg.H * W);
  if (y >= pg.H) return;
  uint32_t prev_last = 0;
  for (int w0 = 0; w0 < wd; w0 += 32) {
    const int w = w0 + lane;
    const uint32_t m = w < wd ? pg.bits[(int64_t)y * wd + w] : 0u;
    uint32_t p31 = __shfl_up_sync(0xffffffffu, m >> 31, 1);
    if (lane == 0) p31 = w0 == 0 ? ((~m) & 1u) : prev_last;
    uint32_t c = m ^ ((m << 1) | p31);
    while (w < wd && c) {
      const int bpos = __ffs(c) - 1;
      c &= c - 1;
      const int x = (w << 5) + bpos;
      if (x >= W) break;
      const int node = y * W + x;
      const int r = uf_find(L, node);
      L[node] = r;
      if (((m >> bpos) & 1u) && r == node) {
        const int slot = atomicAdd(&pg.bufs.counters[0], 1);
        if (slot < pg.bufs.max_comps) pg.bufs.comp_roots[slot] = node;
        else atomicExch(&pg.bufs.counters[2], 1);
      }
    }
    prev_last = __shfl_sync(0xffffffffu, m >> 31, 31);
  }
}

// ---------------------------------------------------------------------------------------------
// Per component: Suzuki-Abe outer border following -> RDP -> hull -> min-area rect.
// ---------------------------------------------------------------------------------------------
// direction tables, packed into immediates (a __constant__ table lookup with a computed index costs a
// dependent constant-cache access on the critical path of every border step):
//   d:  0 E, 1 SE, 2 S, 3 SW, 4 W, 5 NW, 6 N, 7 NE  (clockwise in image coordinates)
//   dy = {0, 1, 1, 1, 0, -1, -1, -1}, dx = {1, 1, 0, -1, -1, -1, 0, 1}, two bits each (+1)
//   dir_of[(dy+1)*3 + (dx+1)] = {5, 6, 7, 4, -, 0, 3, 2, 1}, four bits each
__device__ __forceinline__ int dir_dy(int d) { return (int)((0x01A9u >> (2 * d)) & 3u) - 1; }
__device__ __forceinline__ int dir_dx(int d) { return (int)((0x901Au >> (2 * d)) & 3u) - 1; }
__device__ __forceinline__ int dir_of(int dy, int dx) {
  return (int)((0x1230F4765ull >> (4 * ((dy + 1) * 3 + (dx + 1)))) & 15ull);
}

struct BitView {
  const uint32_t* bits;
  int H, W, wd;
  // the three mask bits (x-1, x, x+1) of row y as bits 0..2; 0 outside the image
  __device__ __forceinline__ uint32_t row3(int y, int x) const {
    if (y < 0 || y >= H) return 0u;
    const uint32_t* r = bits + (int64_t)y * wd;
    const int x0 = x - 1;  // may be -1
    const int w = x0 >> 5;  // arithmetic shift: -1 for x0 = -1
    const uint64_t lo = w >= 0 ? r[w] : 0u;
    const uint64_t hi = (w + 1 < wd) ? r[w + 1] : 0u;
    const uint64_t both = lo | (hi << 32);
    return (uint32_t)(both >> (x0 & 31)) & 7u;  // bits beyond W are 0 in the packed mask
  }
  // 8-neighbourhood of (y, x): bit d = neighbour in direction d
  __device__ __forceinline__ uint32_t nb8(int y, int x) const {
    const uint32_t a = row3(y - 1, x), b = row3(y, x), c = row3(y + 1, x);
    return ((b >> 2) & 1u) | (((c >> 2) & 1u) << 1) | (((c >> 1) & 1u) << 2) | ((c & 1u) << 3) | ((b & 1u) << 4) |
           ((a & 1u) << 5) | (((a >> 1) & 1u) << 6) | (((a >> 2) & 1u) << 7);
  }
};

// Follows the outer border that starts at (i, j) (the component's first pixel in raster
// order).  When `out` is non-null writes (x, y) pairs.  Returns the number of points, or -1 if
// `limit` would be exceeded.  One neighbourhood fetch (three rows of the packed mask) per step.
__device__ int trace_border(const BitView& mv, int i, int j, int16_t* out, int64_t limit) {
  int i1 = 0, j1 = 0;
  {  // (3.1) clockwise from W around (i, j): d = 4, 5, 6, 7, 0, 1, 2, 3
    const uint32_t nb = mv.nb8(i, j);
    const uint32_t t = nb | (nb << 8);
    const uint32_t w = (t >> 4) & 0xFFu;
    if (!w) {  // isolated pixel
      if (limit < 1) return -1;
      if (out) { out[0] = (int16_t)j; out[1] = (int16_t)i; }
      return 1;
    }
    const int d = (4 + (__ffs(w) - 1)) & 7;
    i1 = i + dir_dy(d);
    j1 = j + dir_dx(d);
  }
  int i2 = i1, j2 = j1, i3 = i, j3 = j;
  int n = 0;
  while (true) {
    const int d0 = dir_of(i2 - i3, j2 - j3);
    // counter-clockwise from the direction after d0: d = d0-1, d0-2, ..., d0-8
    const uint32_t nb = mv.nb8(i3, j3);
    const uint32_t t = nb | (nb << 8);
    const uint32_t w = (t >> d0) & 0xFFu;  // bit q = direction (d0 + q) & 7; scan q = 7 down to 0
    int i4 = i3, j4 = j3;
    if (w) {
      const int d = (d0 + (31 - __clz(w))) & 7;
      i4 = i3 + dir_dy(d);
      j4 = j3 + dir_dx(d);
    }
    if (n >= limit) return -1;
    if (out) { out[2 * n] = (int16_t)j3; out[2 * n + 1] = (int16_t)i3; }
    ++n;
    if (i4 == i && j4 == j && i3 == i1 && j3 == j1) break;
    i2 = i3; j2 = j3; i3 = i4; j3 = j4;
  }
  return n;
}

// One WARP per component, components drawn from a per-page counter (the component count is only known on the
// device, so a fixed grid of warps pulls work until the page's list is exhausted):
//   lane 0 follows the border ONCE, into a shared-memory point buffer (borders longer than the buffer are counted
//   first and traced into the global pool, the old two-pass form);
//   the Ramer-Douglas-Peucker scan for the farthest point of a span runs over the 32 lanes (the span order, the
//   distance function and the "last of the farthest" tie rule are those of the sequential form, so the kept points
//   are identical); lane 0 owns the explicit stack;
//   hull + min-area rectangle (a few dozen points) stay on lane 0, on shared-memory copies of the points.
constexpr int kRectWarps = 4;            // warps per block
constexpr int kRectBlocksPerPage = 74;   // x pages x 4 warps: 8 pages put 16 warps on every SM
constexpr int kRectSmemPts = 1024;       // border points per warp kept in shared memory (4 KB)
constexpr int kRectSmemSimp = 128;       // simplified points per warp kept in shared memory (2 x 1 KB)

__global__ void __launch_bounds__(kRectWarps * 32)
component_rects_kernel(const CclPage* __restrict__ pages, float eps, float expand, float min_area) {
  __shared__ int16_t s_pts[kRectWarps][2 * kRectSmemPts];
  __shared__ geom::PointF s_fp[kRectWarps][kRectSmemSimp];
  __shared__ geom::PointF s_hull[kRectWarps][kRectSmemSimp];
  const CclPage pg = pages[blockIdx.y];
  const ComponentBuffers& b = pg.bufs;
  const int H = pg.H, W = pg.W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_comps = min(b.counters[0], b.max_comps);
  const BitView mv{pg.bits, H, W, pg.wd};
  const int bg_label = pg.labels[(int64_t)H * W];

  while (true) {
    int ci = 0;
    if (lane == 0) ci = atomicAdd(&b.counters[6], 1);
    ci = __shfl_sync(0xffffffffu, ci, 0);
    if (ci >= n_comps) return;
    const int root = b.comp_roots[ci];
    const int i = root / W, j = root - i * W;

    // ---- border following (lane 0) ----
    int n = 0;
    long long off = -1;
    bool in_smem = true;
    if (lane == 0) {
      // outermost-border rule (Suzuki-Abe App. II): the 0-pixel left of the first pixel must belong
      // to the background component that touches the frame.
      if (j > 0 && pg.labels[(int64_t)i * W + run_start(pg.bits, pg.wstart, pg.wd, i, j - 1)] != bg_label) {
        n = 0;
      } else {
        n = trace_border(mv, i, j, s_pts[warp], kRectSmemPts);
        if (n < 0) {  // longer than the shared buffer: count, then trace into the pool
          in_smem = false;
          n = trace_border(mv, i, j, nullptr, b.pool_cap);
          if (n < 0) atomicExch(&b.counters[2], 2);
        }
        if (n > 0) {
          off = (long long)atomicAdd((unsigned long long*)(void*)&b.counters[4], (unsigned long long)(n + 2));
          if (off + n + 2 > b.pool_cap) {
            atomicExch(&b.counters[2], 2);
            n = -1;
          } else if (!in_smem) {
            trace_border(mv, i, j, b.pts + 2 * off, n);
          }
        }
      }
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    if (n <= 0) continue;
    off = __shfl_sync(0xffffffffu, off, 0);
    in_smem = __shfl_sync(0xffffffffu, (int)in_smem, 0) != 0;
    __syncwarp();
    const int16_t* pts = in_smem ? s_pts[warp] : (b.pts + 2 * off);

    // ---- RDP on the closed polyline P[0..n], P[n] == P[0] (simplify_polygon) ----
    int32_t* simp = b.simp_idx + off;
    int32_t* stack = b.stack + 3 * off;
    auto P = [&](int idx) -> geom::PointF {
      if (idx == n) idx = 0;
      return geom::PointF{(float)pts[2 * idx], (float)pts[2 * idx + 1]};
    };
    int sp = 1, m = 0;  // warp-uniform
    if (lane == 0) { stack[0] = 0; stack[1] = n; stack[2] = 1; }
    while (sp > 0) {
      --sp;
      int s = 0, e = 0, keep = 0;
      if (lane == 0) { s = stack[3 * sp]; e = stack[3 * sp + 1]; keep = stack[3 * sp + 2]; }
      s = __shfl_sync(0xffffffffu, s, 0);
      e = __shfl_sync(0xffffffffu, e, 0);
      keep = __shfl_sync(0xffffffffu, keep, 0);
      if (e - s + 1 <= 1) {
        if (lane == 0) simp[m] = s;
        ++m;
        continue;
      }
      const geom::LineF seg{P(s), P(e)};
      // sequential rule: max_i = s, max_d = 0; for k in (s, e): if (d >= max_d) take k  ==  the LAST k among the
      // farthest points.  Per lane the same rule over k = s+1+lane, +32, ...; across lanes (larger d, then larger k).
      int max_i = s;
      float max_d = 0.0f;
      for (int k = s + 1 + lane; k < e; k += 32) {
        const float d = geom::line_distance(seg, P(k));
        if (d >= max_d) { max_i = k; max_d = d; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, max_d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, max_i, o);
        if (od > max_d || (od == max_d && oi > max_i)) { max_d = od; max_i = oi; }
      }
      if (max_d > eps) {
        if (lane == 0) {  // left half first: push right, then left
          stack[3 * sp] = max_i; stack[3 * sp + 1] = e; stack[3 * sp + 2] = keep;
          stack[3 * sp + 3] = s; stack[3 * sp + 4] = max_i; stack[3 * sp + 5] = 0;
        }
        sp += 2;
      } else {
        if (lane == 0) {
          simp[m] = s;
          if (keep) simp[m + 1] = e;
        }
        m += keep ? 2 : 1;
      }
    }
    m -= 1;  // drop the duplicated closing point
    if (m < 1) continue;
    __syncwarp();  // simp[] (written by lane 0) is read by every lane below
    const bool small = m <= kRectSmemSimp;
    geom::PointF* fpts = small ? s_fp[warp] : reinterpret_cast<geom::PointF*>(b.fpts + 2 * off);
    geom::PointF* hull = small ? s_hull[warp] : reinterpret_cast<geom::PointF*>(b.hull + 2 * off);
    for (int k = lane; k < m; k += 32) fpts[k] = P(simp[k]);
    __syncwarp();
    if (lane == 0) {
      const int hm = geom::convex_hull(fpts, m, hull);
      geom::RotatedRect rr;
      if (geom::min_area_rect_of_hull(hull, hm, &rr)) {
        rr.w = rr.w + 2.0f * expand;  // detection.rs:53-57
        rr.h = rr.h + 2.0f * expand;
        if (geom::rr_area(rr) >= min_area) {  // detection.rs:60
          const int slot = atomicAdd(&b.counters[3], 1);
          b.rects[slot] = rr;
          b.rect_root[slot] = root;
        }
      }
    }
    __syncwarp();  // the shared buffers are reused by the next component
  }
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

// One block = 128 output columns of one line, ALL output rows: the column taps, the page bounds and the line
// record are set up once per thread instead of once per output pixel, and the line's whole crossing table
// ([lh][max_cross + 1] ints, a few KB) is staged in shared memory once per block (from global memory when it
// does not fit).  Per output pixel: the row taps from shared memory, four inside tests, four loads, one store.
constexpr int kCropSmemInts = 6144;  // 24 KB of crossings
constexpr int kCropRowTaps = 128;    // row taps kept in shared memory (the recognition height is 64)

__global__ void __launch_bounds__(128) crop_resize_kernel(const float* const* __restrict__ pages, const int* __restrict__ page_h,
                                                          const int* __restrict__ page_w, const LineDesc* __restrict__ lines,
                                                          int n_lines, const int32_t* __restrict__ cross, float* __restrict__ dst,
                                                          int out_h) {
  __shared__ int32_t s_cross[kCropSmemInts];
  __shared__ AxisTap s_ty[kCropRowTaps];
  const int li = blockIdx.y;
  if (li >= n_lines) return;
  const LineDesc L = lines[li];
  if ((int)(blockIdx.x * blockDim.x) >= L.group_width) return;  // block-uniform
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const bool has_src = L.lh > 0 && L.lw > 0;
  const int stride = L.max_cross + 1;
  const int32_t* ctab = cross + L.cross_off;
  if (has_src) {
    const int n_ints = L.lh * stride;
    if (n_ints <= kCropSmemInts) {
      for (int k = threadIdx.x; k < n_ints; k += blockDim.x) s_cross[k] = ctab[k];
      ctab = s_cross;
    }
    for (int r = threadIdx.x; r < min(out_h, kCropRowTaps); r += blockDim.x) s_ty[r] = axis_tap(r, L.lh, out_h);
  }
  __syncthreads();
  if (ox >= L.group_width) return;
  float* out = dst + L.dst_off + ox;
  if (!(ox < L.resized_width && has_src)) {
    for (int oy = 0; oy < out_h; ++oy) out[(int64_t)oy * L.group_width] = kBlackValue;
    return;
  }
  const float* page = pages[L.page];
  const int PH = page_h[L.page], PW = page_w[L.page];
  const AxisTap tx = axis_tap(ox, L.lw, L.resized_width);
  // both points (canvas and page coordinates) are tested against the page index rect (recognition.rs:112)
  const int px0 = tx.i0 + L.left, px1 = tx.i1 + L.left;
  const bool xin0 = !(px0 < 0 || px0 > PW - 1 || tx.i0 < 0 || tx.i0 > PW - 1);
  const bool xin1 = !(px1 < 0 || px1 > PW - 1 || tx.i1 < 0 || tx.i1 > PW - 1);
  auto canvas = [&](int cy, int px, bool xin) -> float {
    const int py = cy + L.top;
    if (!xin || py < 0 || py > PH - 1 || cy < 0 || cy > PH - 1) return kBlackValue;
    const int32_t* row = ctab + cy * stride;
    const int cnt = row[0];
    int inside = 0;
    for (int k = 0; k < cnt; ++k) inside += (row[1 + k] <= px) ? 1 : 0;
    return (inside & 1) ? __ldg(page + (int64_t)py * PW + px) : kBlackValue;
  };
  // the line is usually magnified: consecutive output rows share their source rows, whose four samples are reused
  int r0 = -1, r1 = -1;
  float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
  for (int oy = 0; oy < out_h; ++oy) {
    const AxisTap ty = oy < kCropRowTaps ? s_ty[oy] : axis_tap(oy, L.lh, out_h);
    if (ty.i0 != r0) {
      if (ty.i0 == r1) {
        v00 = v10;
        v01 = v11;
      } else {
        v00 = canvas(ty.i0, px0, xin0);
        v01 = canvas(ty.i0, px1, xin1);
      }
      r0 = ty.i0;
    }
    if (ty.i1 != r1) {
      if (ty.i1 == r0) {
        v10 = v00;
        v11 = v01;
      } else {
        v10 = canvas(ty.i1, px0, xin0);
        v11 = canvas(ty.i1, px1, xin1);
      }
      r1 = ty.i1;
    }
    out[(int64_t)oy * L.group_width] = lerp2(v00, v01, v10, v11, tx.w, ty.w);
  }
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

void find_component_rects_batch(const CclPage* d_pages, const CclPage* h_pages, int n_pages, float eps, float expand_dist,
                                float min_area, cudaStream_t st) {
  if (n_pages == 0) return;
  int max_h = 0, max_c = 0;
  for (int i = 0; i < n_pages; ++i) {
    OCRS_CHECK(h_pages[i].H < 32768 && h_pages[i].W < 32768, kInvalidArg, "image dimensions exceed 32767");
    OCRS_CHECK((int64_t)h_pages[i].H * h_pages[i].W < ((int64_t)1 << 31) - 1, kInvalidArg, "image too large");
    OCRS_CUDA_CHECK(cudaMemsetAsync(h_pages[i].bufs.counters, 0, 8 * sizeof(int32_t), st));
    max_h = std::max(max_h, h_pages[i].H);
    max_c = std::max(max_c, h_pages[i].bufs.max_comps);
  }
  if (max_h == 0) return;
  dim3 rows((unsigned)ceil_div(max_h, kCclWarps), (unsigned)n_pages);
  ccl_init_kernel<<<rows, kCclWarps * 32, 0, st>>>(d_pages);
  count_launch();
  ccl_union_kernel<<<rows, kCclWarps * 32, 0, st>>>(d_pages);
  count_launch();
  ccl_flatten_kernel<<<rows, kCclWarps * 32, 0, st>>>(d_pages);
  count_launch();
  // one warp per component, pulled from a per-page counter (counters[6]); the count is only known on the device,
  // so a fixed grid of warps per page loops until the page's list is exhausted
  (void)max_c;
  dim3 comps((unsigned)kRectBlocksPerPage, (unsigned)n_pages);
  component_rects_kernel<<<comps, kRectWarps * 32, 0, st>>>(d_pages, eps, expand_dist, min_area);
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
  dim3 g2(grid1d(max_group_width, 128), n_lines);
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
