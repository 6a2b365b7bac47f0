// fp32 CUDA-core operator kernels (sm_100a).  First-generation implementations: correct and
// coalesced; the dense 3x3 convolutions and the GRU are superseded by the tcgen05 kernels in
// conv_tc.cu / gru_tc.cu when those are enabled (DESIGN.md "Kernels").
#include "nn_kernels.h"

#include <cstdlib>

#include <cfloat>

#include "common.h"

namespace ocrs {
namespace nn {

namespace {

constexpr int kThreads = 256;
inline int grid1d(int64_t n, int threads = kThreads) { return (int)ceil_div(n, threads); }

// ---------------------------------------------------------------------------------------------
// Generic direct convolution (any groups): one thread per output element.
// ---------------------------------------------------------------------------------------------
__global__ void conv_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ b, float* __restrict__ y, ConvParams p) {
  int64_t total = (int64_t)p.N * p.K * p.OH * p.OW;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int ow = idx % p.OW;
  int oh = (idx / p.OW) % p.OH;
  int k = (idx / ((int64_t)p.OW * p.OH)) % p.K;
  int n = idx / ((int64_t)p.OW * p.OH * p.K);
  int cpg = p.C / p.groups;  // in channels per group
  int kpg = p.K / p.groups;
  int g = k / kpg;
  float acc = b ? b[k] : 0.f;
  const float* wk = w + (int64_t)k * cpg * p.R * p.S;
  for (int c = 0; c < cpg; ++c) {
    const float* xc = x + ((int64_t)n * p.C + g * cpg + c) * p.H * p.W;
    for (int r = 0; r < p.R; ++r) {
      int ih = oh * p.stride_h - p.pad_t + r * p.dil_h;
      if (ih < 0 || ih >= p.H) continue;
      for (int s = 0; s < p.S; ++s) {
        int iw = ow * p.stride_w - p.pad_l + s * p.dil_w;
        if (iw < 0 || iw >= p.W) continue;
        acc = fmaf(xc[(int64_t)ih * p.W + iw], wk[(c * p.R + r) * p.S + s], acc);
      }
    }
  }
  if (p.relu) acc = fmaxf(acc, 0.f);
  y[idx] = acc;
}

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM convolution, groups == 1.  M = K (out channels), N = pixels, Kdim = C*R*S.
// 64x64 tile, BK = 16, 256 threads, 4x4 register block.
// ---------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) conv_igemm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, float* __restrict__ y,
                                                         ConvParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t npix = (int64_t)p.N * p.OH * p.OW;
  const int64_t n0 = (int64_t)blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int RS = p.R * p.S;
  const int CRS = p.C * RS;
  const int64_t HW = (int64_t)p.H * p.W;

  // B-tile loader: fixed pixel column per thread
  const int ln = tid % BN;
  const int lk0 = tid / BN;  // 0..3
  const int64_t lpix = n0 + ln;
  const bool lvalid = lpix < npix;
  int l_img = 0, l_ih0 = 0, l_iw0 = 0;
  if (lvalid) {
    int ow = lpix % p.OW;
    int oh = (lpix / p.OW) % p.OH;
    l_img = lpix / ((int64_t)p.OW * p.OH);
    l_ih0 = oh * p.stride_h - p.pad_t;
    l_iw0 = ow * p.stride_w - p.pad_l;
  }
  const float* xin = x + (int64_t)l_img * p.C * HW;

  // A-tile loader
  const int ak = tid % BK;
  const int am0 = tid / BK;  // 0..15

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < CRS; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = am0 + i * 16;
      int kk = k0 + ak;
      float v = 0.f;
      if (m0 + m < p.K && kk < CRS) v = w[(int64_t)(m0 + m) * CRS + kk];
      As[ak][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int kk = lk0 + i * 4;
      int kidx = k0 + kk;
      float v = 0.f;
      if (lvalid && kidx < CRS) {
        int c = kidx / RS;
        int rs = kidx - c * RS;
        int r = rs / p.S;
        int s = rs - r * p.S;
        int ih = l_ih0 + r * p.dil_h;
        int iw = l_iw0 + s * p.dil_w;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v = xin[(int64_t)c * HW + (int64_t)ih * p.W + iw];
      }
      Bs[kk][ln] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int64_t OHW = (int64_t)p.OH * p.OW;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int64_t pix = n0 + tx * 4 + j;
    if (pix >= npix) continue;
    int64_t img = pix / OHW;
    int64_t rem = pix - img * OHW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = m0 + ty * 4 + i;
      if (k >= p.K) continue;
      float v = acc[i][j] + (b ? b[k] : 0.f);
      if (p.relu) v = fmaxf(v, 0.f);
      y[(img * p.K + k) * OHW + rem] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void conv_transpose_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ b, float* __restrict__ y, ConvTParams p) {
  int64_t total = (int64_t)p.N * p.K * p.OH * p.OW;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int ow = idx % p.OW;
  int oh = (idx / p.OW) % p.OH;
  int k = (idx / ((int64_t)p.OW * p.OH)) % p.K;
  int n = idx / ((int64_t)p.OW * p.OH * p.K);
  int kpg = p.K / p.groups, cpg = p.C / p.groups;
  int g = k / kpg, kk = k - g * kpg;
  float acc = b ? b[k] : 0.f;
  for (int r = 0; r < p.R; ++r) {
    int th = oh + p.pad_t - r;
    if (th < 0 || th % p.stride_h) continue;
    int ih = th / p.stride_h;
    if (ih >= p.H) continue;
    for (int s = 0; s < p.S; ++s) {
      int tw = ow + p.pad_l - s;
      if (tw < 0 || tw % p.stride_w) continue;
      int iw = tw / p.stride_w;
      if (iw >= p.W) continue;
      for (int c = 0; c < cpg; ++c) {
        int ci = g * cpg + c;
        float xv = x[(((int64_t)n * p.C + ci) * p.H + ih) * p.W + iw];
        float wv = w[(((int64_t)ci * kpg + kk) * p.R + r) * p.S + s];
        acc = fmaf(xv, wv, acc);
      }
    }
  }
  if (p.relu) acc = fmaxf(acc, 0.f);
  y[idx] = acc;
}

// ---------------------------------------------------------------------------------------------
// Fused depthwise 3x3 + pointwise 1x1 (+ReLU).  One thread per output pixel, KOUT accumulators in
// registers, all weights in shared memory.  Removes the depthwise intermediate (write + read).
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Separable-convolution family of the detection U-Net (detection.rs:131-184 runs it through rten):
//   phase 1  per (pixel, channel): depthwise 3x3 (pad 1, stride 1 | 2) over a VIRTUAL input -- the channel
//            concatenation of up to two tensors, each optionally end-padded with a constant (the decoder's
//            Pad + Concat never materialise) -- or, for ConvTranspose, the input value itself; result in
//            shared memory as [C][P];
//   phase 2  pointwise: [P x C] x [C x KO] from shared memory, KC = 32 output channels at a time, PPT pixels x
//            KPT outputs per thread; epilogue bias (+ ReLU) and either a plain NCHW store or the 2x2
//            pixel shuffle of ConvTranspose(k = 2, s = 2) (KO = 4 K, column kk = k*4 + a*2 + b).
// Two shapes: P = 128 pixels per block for the wide, shallow levels; P = 32 for the deep levels (few pixels,
// C up to 256) so that they still fill SMs.
// ---------------------------------------------------------------------------------------------
struct SepSrc {
  const float* x;  // [N, C, H, W]
  int C, H, W;     // valid extent; reads at y >= H or x >= W (inside the virtual grid) return `pad`
  float pad;
};
struct SepParams {
  SepSrc src[2];
  int n_src;
  int N, C, Hv, Wv;        // virtual input: C = sum of source channels, Hv x Wv
  int OH, OW, stride;      // grid of phase-1 pixels (depthwise output; for ConvTranspose = input grid)
  int KO;                  // pointwise outputs (ConvTranspose: 4 * K)
  int relu, mode;          // mode 0: depthwise + pointwise, NCHW store; 1: ConvTranspose 2x2 s2 (no depthwise)
  int w_ck;                // pointwise weights already laid out [C][KO] (ConvTranspose); else [KO][C]
  const float *dw_w, *dw_b, *pw_w, *pw_b;
  float* y;
};

constexpr int kSepKC = 32;

// Shallow levels (few channels, many pixels): one thread per output pixel, KOUT accumulators in registers,
// all weights in shared memory; the depthwise value of each channel is consumed as soon as it is formed.
template <int KOUT>
__global__ void __launch_bounds__(128) dwpw_pixel_kernel(SepParams p) {
  extern __shared__ float sm[];
  float* s_dw = sm;                   // [C][9]
  float* s_db = s_dw + p.C * 9;       // [C]
  float* s_pw = s_db + p.C;           // [C][KOUT]  (transposed for broadcast reads)
  float* s_pb = s_pw + p.C * KOUT;    // [KOUT]
  for (int i = threadIdx.x; i < p.C * 9; i += blockDim.x) s_dw[i] = p.dw_w[i];
  for (int i = threadIdx.x; i < p.C; i += blockDim.x) s_db[i] = p.dw_b ? p.dw_b[i] : 0.f;
  for (int i = threadIdx.x; i < p.C * KOUT; i += blockDim.x) {
    int c = i / KOUT, k = i - c * KOUT;
    s_pw[i] = p.pw_w[k * p.C + c];
  }
  for (int i = threadIdx.x; i < KOUT; i += blockDim.x) s_pb[i] = p.pw_b ? p.pw_b[i] : 0.f;
  __syncthreads();
  const int ohw = p.OH * p.OW;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (idx >= ohw) return;
  const int oy = idx / p.OW, ox = idx - oy * p.OW;
  float acc[KOUT];
#pragma unroll
  for (int k = 0; k < KOUT; ++k) acc[k] = s_pb[k];
  const int iy0 = oy * p.stride - 1, ix0 = ox * p.stride - 1;
  int c = 0;
#pragma unroll 1
  for (int si = 0; si < p.n_src; ++si) {
    const SepSrc s = p.src[si];
    const float* xs = s.x + (int64_t)n * s.C * s.H * s.W;
    // tap validity is the same for every channel of the source
    bool inside[9], real[9];
    int off[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int iy = iy0 + r, ix = ix0 + q;
        inside[r * 3 + q] = iy >= 0 && iy < p.Hv && ix >= 0 && ix < p.Wv;
        real[r * 3 + q] = inside[r * 3 + q] && iy < s.H && ix < s.W;
        off[r * 3 + q] = iy * s.W + ix;
      }
    for (int cs = 0; cs < s.C; ++cs, ++c) {
      const float* xc = xs + (int64_t)cs * s.H * s.W;
      const float* wd = s_dw + c * 9;
      float v = s_db[c];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float xv = real[t] ? __ldg(xc + off[t]) : (inside[t] ? s.pad : 0.f);
        v = fmaf(xv, wd[t], v);
      }
      const float* wp = s_pw + c * KOUT;
#pragma unroll
      for (int k = 0; k < KOUT; ++k) acc[k] = fmaf(v, wp[k], acc[k]);
    }
  }
  float* yn = p.y + (int64_t)n * KOUT * ohw + idx;
#pragma unroll
  for (int k = 0; k < KOUT; ++k) {
    float v = acc[k];
    if (p.relu) v = fmaxf(v, 0.f);
    yn[(int64_t)k * ohw] = v;
  }
}

// Two vertically adjacent output pixels per thread: their 3x3 windows share rows (S + 3 input rows instead of 6), so a
// channel costs 12 (stride 1) or 15 (stride 2) loads instead of 18, and the depthwise / pointwise weights are read
// from shared memory once for both pixels.  Consecutive threads are consecutive pixels of a row, as in the
// one-pixel kernel, so loads and stores stay coalesced.  Per pixel the arithmetic (order of the FMAs) is unchanged.
template <int KOUT, int S>
__global__ void __launch_bounds__(128) dwpw_pixel2_kernel(SepParams p) {
  constexpr int R = S + 3;  // input rows of the pair
  extern __shared__ float sm[];
  float* s_dw = sm;                   // [C][9]
  float* s_db = s_dw + p.C * 9;       // [C]
  float* s_pw = s_db + p.C;           // [C][KOUT]  (transposed for broadcast reads)
  float* s_pb = s_pw + p.C * KOUT;    // [KOUT]
  for (int i = threadIdx.x; i < p.C * 9; i += blockDim.x) s_dw[i] = p.dw_w[i];
  for (int i = threadIdx.x; i < p.C; i += blockDim.x) s_db[i] = p.dw_b ? p.dw_b[i] : 0.f;
  for (int i = threadIdx.x; i < p.C * KOUT; i += blockDim.x) {
    int c = i / KOUT, k = i - c * KOUT;
    s_pw[i] = p.pw_w[k * p.C + c];
  }
  for (int i = threadIdx.x; i < KOUT; i += blockDim.x) s_pb[i] = p.pw_b ? p.pw_b[i] : 0.f;
  __syncthreads();
  const int ohw = p.OH * p.OW;
  const int pairs = ((p.OH + 1) / 2) * p.OW;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (idx >= pairs) return;
  const int oy = 2 * (idx / p.OW), ox = idx % p.OW;
  const bool second = oy + 1 < p.OH;
  float acc0[KOUT], acc1[KOUT];
#pragma unroll
  for (int k = 0; k < KOUT; ++k) acc0[k] = acc1[k] = s_pb[k];
  const int iy0 = oy * S - 1, ix0 = ox * S - 1;
  int c = 0;
#pragma unroll 1
  for (int si = 0; si < p.n_src; ++si) {
    const SepSrc s = p.src[si];
    const float* xs = s.x + (int64_t)n * s.C * s.H * s.W;
    // tap validity is the same for every channel of the source
    bool inside[R * 3], real[R * 3];
    int off[R * 3];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int iy = iy0 + r, ix = ix0 + q;
        inside[r * 3 + q] = iy >= 0 && iy < p.Hv && ix >= 0 && ix < p.Wv;
        real[r * 3 + q] = inside[r * 3 + q] && iy < s.H && ix < s.W;
        off[r * 3 + q] = iy * s.W + ix;
      }
    for (int cs = 0; cs < s.C; ++cs, ++c) {
      const float* xc = xs + (int64_t)cs * s.H * s.W;
      const float* wd = s_dw + c * 9;
      float x[R * 3];
#pragma unroll
      for (int t = 0; t < R * 3; ++t) x[t] = real[t] ? __ldg(xc + off[t]) : (inside[t] ? s.pad : 0.f);
      float w9[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) w9[t] = wd[t];
      float v0 = s_db[c], v1 = v0;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        v0 = fmaf(x[t], w9[t], v0);
        v1 = fmaf(x[S * 3 + t], w9[t], v1);
      }
      const float* wp = s_pw + c * KOUT;
#pragma unroll
      for (int k = 0; k < KOUT; ++k) {
        const float w = wp[k];
        acc0[k] = fmaf(v0, w, acc0[k]);
        acc1[k] = fmaf(v1, w, acc1[k]);
      }
    }
  }
  float* yn = p.y + (int64_t)n * KOUT * ohw + (int64_t)oy * p.OW + ox;
#pragma unroll
  for (int k = 0; k < KOUT; ++k) {
    float v = acc0[k];
    if (p.relu) v = fmaxf(v, 0.f);
    yn[(int64_t)k * ohw] = v;
    if (second) {
      float u = acc1[k];
      if (p.relu) u = fmaxf(u, 0.f);
      yn[(int64_t)k * ohw + p.OW] = u;
    }
  }
}

template <int P, int PPT, int KPT>
__global__ void __launch_bounds__(256) sepconv_kernel(SepParams p) {
  extern __shared__ float sm[];
  float* s_dwv = sm;                          // [C][P]
  float* s_pw = s_dwv + (size_t)p.C * P;      // [C][kSepKC]
  float* s_dw = s_pw + (size_t)p.C * kSepKC;  // [C][9] + [C] (mode 0)
  const int tid = threadIdx.x;
  const int ohw = p.OH * p.OW;
  const int npix = p.N * ohw;
  const int pix0 = blockIdx.x * P;
  const int k0 = blockIdx.y * kSepKC;         // this block's chunk of pointwise outputs
  if (p.mode == 0) {
    for (int i = tid; i < p.C * 9; i += 256) s_dw[i] = p.dw_w[i];
    for (int i = tid; i < p.C; i += 256) s_dw[p.C * 9 + i] = p.dw_b ? p.dw_b[i] : 0.f;
  }
  for (int i = tid; i < p.C * kSepKC; i += 256) {
    const int c = i / kSepKC, kk = i - c * kSepKC;
    float w = 0.f;
    if (k0 + kk < p.KO) w = p.w_ck ? p.pw_w[(int64_t)c * p.KO + k0 + kk] : p.pw_w[(int64_t)(k0 + kk) * p.C + c];
    s_pw[i] = w;
  }
  __syncthreads();
  // ---- phase 1: every (channel, pixel) of the block; pixel decode once per thread (P divides 256 or vice versa) ----
  {
    constexpr int CSTEP = 256 / P;            // channels advanced per iteration (P = 32: 8; P = 128: 2)
    const int pp = tid % P, c_first = tid / P;
    const int pix = pix0 + pp;
    const bool live = pix < npix;
    int n = 0, oy = 0, ox = 0;
    if (live) {
      n = pix / ohw;
      const int rem = pix - n * ohw;
      oy = rem / p.OW;
      ox = rem - oy * p.OW;
    }
    const int iy0 = oy * p.stride - 1, ix0 = ox * p.stride - 1;
    for (int c = c_first; c < p.C; c += CSTEP) {
      float v = 0.f;
      if (live) {
        const bool second = p.n_src > 1 && c >= p.src[0].C;
        const SepSrc& s = second ? p.src[1] : p.src[0];
        const int cs = second ? c - p.src[0].C : c;
        const float* xc = s.x + ((int64_t)n * s.C + cs) * s.H * s.W;
        if (p.mode == 0) {
          const float* wd = s_dw + c * 9;
          v = s_dw[p.C * 9 + c];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const int iy = iy0 + r;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const int ix = ix0 + q;
              float xv = 0.f;  // conv zero padding outside the virtual grid
              if (iy >= 0 && iy < p.Hv && ix >= 0 && ix < p.Wv) xv = (iy < s.H && ix < s.W) ? __ldg(xc + iy * s.W + ix) : s.pad;
              v = fmaf(xv, wd[r * 3 + q], v);
            }
          }
        } else {
          v = __ldg(xc + oy * s.W + ox);
        }
      }
      s_dwv[c * P + pp] = v;
    }
  }
  __syncthreads();
  // ---- phase 2 ----
  constexpr int PL = P / PPT;            // pixel lanes
  constexpr int KG = kSepKC / KPT;       // output groups
  static_assert(PL * KG == 256, "thread mapping");
  const int pl = tid % PL, kq = tid / PL;
  float acc[PPT][KPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j)
#pragma unroll
    for (int k = 0; k < KPT; ++k) acc[j][k] = 0.f;
  const float* wrow = s_pw + kq * KPT;
#pragma unroll 4
  for (int c = 0; c < p.C; ++c) {
    float a[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) a[j] = s_dwv[c * P + pl + j * PL];
    float w[KPT];
#pragma unroll
    for (int k = 0; k < KPT; k += 4) {
      const float4 t = *reinterpret_cast<const float4*>(wrow + c * kSepKC + k);
      w[k] = t.x; w[k + 1] = t.y; w[k + 2] = t.z; w[k + 3] = t.w;
    }
#pragma unroll
    for (int j = 0; j < PPT; ++j)
#pragma unroll
      for (int k = 0; k < KPT; ++k) acc[j][k] = fmaf(a[j], w[k], acc[j][k]);
  }
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int pix = pix0 + pl + j * PL;
    if (pix >= npix) continue;
    const int n = pix / ohw;
    const int rem = pix - n * ohw;
    const int oy = rem / p.OW, ox = rem - oy * p.OW;
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      const int kk = k0 + kq * KPT + k;
      if (kk >= p.KO) continue;
      if (p.mode == 0) {
        float v = acc[j][k] + (p.pw_b ? __ldg(p.pw_b + kk) : 0.f);
        if (p.relu) v = fmaxf(v, 0.f);
        p.y[((int64_t)n * p.KO + kk) * ohw + rem] = v;
      } else {
        const int ko = kk >> 2, t = kk & 3;
        float v = acc[j][k] + (p.pw_b ? __ldg(p.pw_b + ko) : 0.f);
        if (p.relu) v = fmaxf(v, 0.f);
        p.y[(((int64_t)n * (p.KO >> 2) + ko) * (2 * p.OH) + 2 * oy + (t >> 1)) * (2 * p.OW) + 2 * ox + (t & 1)] = v;
      }
    }
  }
}

// ConvTranspose 2x2 stride 2: one thread per input pixel and chunk of 8 output channels.
constexpr int CT_KB = 8;
__global__ void __launch_bounds__(128)
convt2x2_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ y, int N, int C, int H, int W, int K, int relu) {
  extern __shared__ float sw[];  // [C][CT_KB][4]
  const int k0 = blockIdx.y * CT_KB;
  for (int i = threadIdx.x; i < C * CT_KB * 4; i += blockDim.x) {
    int c = i / (CT_KB * 4), r = i - c * CT_KB * 4, kk = r / 4, t = r - kk * 4;
    sw[i] = (k0 + kk < K) ? w[((int64_t)c * K + k0 + kk) * 4 + t] : 0.f;
  }
  __syncthreads();
  const int64_t HW = (int64_t)H * W;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * HW) return;
  const int n = (int)(idx / HW);
  const int64_t rem = idx - (int64_t)n * HW;
  const int iy = (int)(rem / W), ix = (int)(rem - (int64_t)iy * W);
  float acc[CT_KB][4];
#pragma unroll
  for (int kk = 0; kk < CT_KB; ++kk) {
    const float bb = (b && k0 + kk < K) ? b[k0 + kk] : 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[kk][t] = bb;
  }
  const float* xp = x + (int64_t)n * C * HW + rem;
  for (int c = 0; c < C; ++c) {
    const float v = __ldg(xp + (int64_t)c * HW);
    const float* wc = sw + c * CT_KB * 4;
#pragma unroll
    for (int kk = 0; kk < CT_KB; ++kk)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[kk][t] = fmaf(v, wc[kk * 4 + t], acc[kk][t]);
  }
  const int OW = 2 * W;
  const int64_t OHW = 4 * HW;
#pragma unroll
  for (int kk = 0; kk < CT_KB; ++kk) {
    if (k0 + kk >= K) break;
    float* yo = y + ((int64_t)n * K + k0 + kk) * OHW + (int64_t)(2 * iy) * OW + 2 * ix;
    float2 top = make_float2(acc[kk][0], acc[kk][1]), bot = make_float2(acc[kk][2], acc[kk][3]);
    if (relu) {
      top.x = fmaxf(top.x, 0.f); top.y = fmaxf(top.y, 0.f);
      bot.x = fmaxf(bot.x, 0.f); bot.y = fmaxf(bot.y, 0.f);
    }
    *reinterpret_cast<float2*>(yo) = top;
    *reinterpret_cast<float2*>(yo + OW) = bot;
  }
}

// ConvT 2x2 s2 (C -> Km) + ReLU + Conv1x1 (Km -> 1) + Sigmoid; one thread per input pixel.
__global__ void __launch_bounds__(128)
convt2x2_head_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                     const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ y, int N, int C,
                     int H, int W, int Km) {
  __shared__ float sw[16 * 16 * 4 + 16 + 16 + 1];
  float* s_b = sw + C * Km * 4;
  float* s_w2 = s_b + Km;
  for (int i = threadIdx.x; i < C * Km * 4; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Km; i += blockDim.x) {
    s_b[i] = b ? b[i] : 0.f;
    s_w2[i] = w2[i];
  }
  if (threadIdx.x == 0) s_w2[Km] = b2 ? b2[0] : 0.f;
  __syncthreads();
  const int64_t HW = (int64_t)H * W;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * HW) return;
  const int n = (int)(idx / HW);
  const int64_t rem = idx - (int64_t)n * HW;
  const int iy = (int)(rem / W), ix = (int)(rem - (int64_t)iy * W);
  float xin[16];
  const float* xp = x + (int64_t)n * C * HW + rem;
  for (int c = 0; c < C; ++c) xin[c] = __ldg(xp + (int64_t)c * HW);
  float o[4] = {s_w2[Km], s_w2[Km], s_w2[Km], s_w2[Km]};
  for (int k = 0; k < Km; ++k) {
    float a[4] = {s_b[k], s_b[k], s_b[k], s_b[k]};
    for (int c = 0; c < C; ++c) {
      const float* wc = sw + (c * Km + k) * 4;
#pragma unroll
      for (int t = 0; t < 4; ++t) a[t] = fmaf(xin[c], wc[t], a[t]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) o[t] = fmaf(fmaxf(a[t], 0.f), s_w2[k], o[t]);
  }
  const int OW = 2 * W;
  float* yo = y + (int64_t)n * 4 * HW + (int64_t)(2 * iy) * OW + 2 * ix;
  float2 top = make_float2(1.f / (1.f + expf(-o[0])), 1.f / (1.f + expf(-o[1])));
  float2 bot = make_float2(1.f / (1.f + expf(-o[2])), 1.f / (1.f + expf(-o[3])));
  *reinterpret_cast<float2*>(yo) = top;
  *reinterpret_cast<float2*>(yo + OW) = bot;
}

template <bool kMax>
__global__ void pool_kernel(const float* __restrict__ x, float* __restrict__ y, PoolParams p) {
  int64_t total = (int64_t)p.NC * p.OH * p.OW;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int ow = idx % p.OW;
  int oh = (idx / p.OW) % p.OH;
  int64_t nc = idx / ((int64_t)p.OW * p.OH);
  const float* xc = x + nc * p.H * p.W;
  float acc = kMax ? -FLT_MAX : 0.f;
  if (kMax) acc = -INFINITY;
  int cnt = 0;
  for (int r = 0; r < p.R; ++r) {
    int ih = oh * p.stride_h - p.pad_t + r;
    if (ih < 0 || ih >= p.H) continue;
    for (int s = 0; s < p.S; ++s) {
      int iw = ow * p.stride_w - p.pad_l + s;
      if (iw < 0 || iw >= p.W) continue;
      float v = xc[(int64_t)ih * p.W + iw];
      if (kMax) acc = fmaxf(acc, v);
      else acc += v;
      ++cnt;
    }
  }
  if (!kMax) acc = acc / (float)(p.count_include_pad ? p.R * p.S : cnt);
  y[idx] = acc;
}

enum class Unary { kRelu, kSigmoid, kTanh };
template <Unary U>
__global__ void unary_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (U == Unary::kRelu) v = fmaxf(v, 0.f);
  else if (U == Unary::kSigmoid) v = 1.f / (1.f + expf(-v));
  else v = tanhf(v);
  y[i] = v;
}

__global__ void add_bcast_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                 int64_t n, int64_t nb) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = a[i] + b[i % nb];
}

__global__ void fill_kernel(float* y, float v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = v;
}

// ---------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * B[N,K]^T + bias.  64x64x16 tile, 4x4 per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       const float* __restrict__ bias, float* __restrict__ C, int M,
                                                       int N, int K, int relu) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lk = tid % BK, lr0 = tid / BK;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r = lr0 + i * 16;
      int kk = k0 + lk;
      As[lk][r] = (m0 + r < M && kk < K) ? A[(int64_t)(m0 + r) * K + kk] : 0.f;
      Bs[lk][r] = (n0 + r < N && kk < K) ? B[(int64_t)(n0 + r) * K + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      C[(int64_t)m * N + n] = v;
    }
  }
}

struct PermuteArgs {
  int64_t out_shape[6];
  int64_t in_stride[6];  // stride of the input dim that feeds output dim i
  int ndim;
};
__global__ void permute_kernel(const float* __restrict__ x, float* __restrict__ y, PermuteArgs a, int64_t total) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int64_t rem = idx, src = 0;
  for (int d = a.ndim - 1; d >= 0; --d) {
    int64_t c = rem % a.out_shape[d];
    rem /= a.out_shape[d];
    src += c * a.in_stride[d];
  }
  y[idx] = x[src];
}

__global__ void concat_copy_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t outer,
                                   int64_t len_src, int64_t len_dst, int64_t dst_off, int64_t inner) {
  int64_t total = outer * len_src * inner;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int64_t in_i = idx % inner;
  int64_t l = (idx / inner) % len_src;
  int64_t o = idx / (inner * len_src);
  y[(o * len_dst + dst_off + l) * inner + in_i] = x[idx];
}

struct Pad4Args {
  int64_t in_shape[4], begin[4], out_shape[4];
};
__global__ void pad4d_kernel(const float* __restrict__ x, float* __restrict__ y, Pad4Args a, float value,
                             int64_t total) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int64_t rem = idx;
  int64_t c[4];
  for (int d = 3; d >= 0; --d) {
    c[d] = rem % a.out_shape[d];
    rem /= a.out_shape[d];
  }
  int64_t src = 0;
  bool inside = true;
  for (int d = 0; d < 4; ++d) {
    int64_t s = c[d] - a.begin[d];
    if (s < 0 || s >= a.in_shape[d]) inside = false;
    src = src * a.in_shape[d] + s;
  }
  y[idx] = inside ? x[src] : value;
}

// one warp per row
__global__ void log_softmax_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t rows, int cols) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, xr[c]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += expf(xr[c] - m);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float lse = m + logf(s);
  for (int c = lane; c < cols; c += 32) y[row * cols + c] = xr[c] - lse;
}

// ---------------------------------------------------------------------------------------------
// GRU step: tile = 16 lines x 64 hidden units, K chunks of 32.
// ---------------------------------------------------------------------------------------------
constexpr int GL = 16, GJ = 64, GK = 32;
__global__ void __launch_bounds__(256) gru_step_kernel(const float* __restrict__ xw, const float* __restrict__ R,
                                                       const float* __restrict__ Rb, const float* __restrict__ h_in,
                                                       float* __restrict__ h_out, float* __restrict__ Y, int D, int T,
                                                       int N, int H, int t0, int t1, int lbr) {
  __shared__ float hs[GL][GK];
  __shared__ float Rs[3][GJ][GK + 1];
  const int d = blockIdx.z;
  const int t = d == 0 ? t0 : t1;
  const int l0 = blockIdx.x * GL, j0 = blockIdx.y * GJ;
  const int tid = threadIdx.x;
  const int tj = tid % GJ, tl = tid / GJ;  // tl in 0..3 -> lines tl*4 .. tl*4+3
  const float* Rd = R + (int64_t)d * 3 * H * H;
  const float* hd = h_in + (int64_t)d * N * H;
  float acc[3][4];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[g][i] = 0.f;
  for (int k0 = 0; k0 < H; k0 += GK) {
    // h tile: 16 x 32 = 512 elements
    for (int e = tid; e < GL * GK; e += 256) {
      int l = e / GK, kk = e % GK;
      hs[l][kk] = (l0 + l < N && k0 + kk < H) ? hd[(int64_t)(l0 + l) * H + k0 + kk] : 0.f;
    }
    // R tile: 3 x 64 x 32 = 6144 elements
    for (int e = tid; e < 3 * GJ * GK; e += 256) {
      int kk = e % GK;
      int j = (e / GK) % GJ;
      int g = e / (GK * GJ);
      Rs[g][j][kk] = (j0 + j < H && k0 + kk < H) ? Rd[((int64_t)g * H + j0 + j) * H + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < GK; ++kk) {
      float rz = Rs[0][tj][kk], rr = Rs[1][tj][kk], rn = Rs[2][tj][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float hv = hs[tl * 4 + i][kk];
        acc[0][i] = fmaf(hv, rz, acc[0][i]);
        acc[1][i] = fmaf(hv, rr, acc[1][i]);
        acc[2][i] = fmaf(hv, rn, acc[2][i]);
      }
    }
    __syncthreads();
  }
  const int j = j0 + tj;
  if (j >= H) return;
  const float* rb = Rb + (int64_t)d * 3 * H;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = l0 + tl * 4 + i;
    if (n >= N) continue;
    const float* xg = xw + (((int64_t)d * T + t) * N + n) * 3 * H;
    float hp = hd[(int64_t)n * H + j];
    float z = 1.f / (1.f + expf(-(xg[j] + acc[0][i] + rb[j])));
    float r = 1.f / (1.f + expf(-(xg[H + j] + acc[1][i] + rb[H + j])));
    float nn_;
    if (lbr) nn_ = tanhf(xg[2 * H + j] + r * (acc[2][i] + rb[2 * H + j]));
    else nn_ = tanhf(xg[2 * H + j] + acc[2][i] + rb[2 * H + j]);  // (r*h) R^T handled by caller when !lbr
    float hnew = (1.f - z) * nn_ + z * hp;
    h_out[((int64_t)d * N + n) * H + j] = hnew;
    Y[(((int64_t)t * D + d) * N + n) * H + j] = hnew;
  }
}

}  // namespace

void conv2d(const float* x, const float* w, const float* b, float* y, const ConvParams& p, cudaStream_t st) {
  int64_t total = (int64_t)p.N * p.K * p.OH * p.OW;
  if (total == 0) return;
  if (p.groups == 1 && p.C * p.R * p.S >= 16 && p.K >= 16) {
    int64_t npix = (int64_t)p.N * p.OH * p.OW;
    dim3 grid((unsigned)ceil_div(npix, BN), (unsigned)ceil_div(p.K, BM));
    conv_igemm_kernel<<<grid, 256, 0, st>>>(x, w, b, y, p);
  count_launch();
  } else {
    conv_direct_kernel<<<grid1d(total), kThreads, 0, st>>>(x, w, b, y, p);
  count_launch();
  }
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void conv_transpose2d(const float* x, const float* w, const float* b, float* y, const ConvTParams& p,
                      cudaStream_t st) {
  int64_t total = (int64_t)p.N * p.K * p.OH * p.OW;
  if (total == 0) return;
  conv_transpose_kernel<<<grid1d(total), kThreads, 0, st>>>(x, w, b, y, p);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

namespace {
void launch_sep(const SepParams& p, cudaStream_t st) {
  const int64_t npix = (int64_t)p.N * p.OH * p.OW;
  if (npix == 0 || p.KO == 0) return;
  OCRS_CHECK(npix < ((int64_t)1 << 30), kInternal, "separable conv: too many pixels for one launch");
  // the choice depends on the PAGE size only, never on the batch size: a page's result is the same
  // bits whether it runs alone or in a batch of eight
  if (p.mode == 0 && p.C <= 64 && (p.KO == 8 || p.KO == 16 || p.KO == 32) && (int64_t)p.OH * p.OW >= 4096) {
    // shallow levels: thread per pixel
    const size_t smem = (size_t)(p.C * 9 + p.C + p.C * p.KO + p.KO) * sizeof(float);
    static const bool one_px = std::getenv("OCRS_B200_DWPW_ONE_PIXEL") != nullptr;  // the one-pixel-per-thread form
    if (!one_px && (p.stride == 1 || p.stride == 2)) {
      dim3 grid((unsigned)ceil_div((int64_t)((p.OH + 1) / 2) * p.OW, 128), (unsigned)p.N);
      if (p.stride == 1) {
        if (p.KO == 8) dwpw_pixel2_kernel<8, 1><<<grid, 128, smem, st>>>(p);
        else if (p.KO == 16) dwpw_pixel2_kernel<16, 1><<<grid, 128, smem, st>>>(p);
        else dwpw_pixel2_kernel<32, 1><<<grid, 128, smem, st>>>(p);
      } else {
        if (p.KO == 8) dwpw_pixel2_kernel<8, 2><<<grid, 128, smem, st>>>(p);
        else if (p.KO == 16) dwpw_pixel2_kernel<16, 2><<<grid, 128, smem, st>>>(p);
        else dwpw_pixel2_kernel<32, 2><<<grid, 128, smem, st>>>(p);
      }
    } else {
      dim3 grid((unsigned)ceil_div((int64_t)p.OH * p.OW, 128), (unsigned)p.N);
      if (p.KO == 8) dwpw_pixel_kernel<8><<<grid, 128, smem, st>>>(p);
      else if (p.KO == 16) dwpw_pixel_kernel<16><<<grid, 128, smem, st>>>(p);
      else dwpw_pixel_kernel<32><<<grid, 128, smem, st>>>(p);
    }
    count_launch();
    OCRS_CHECK(cudaGetLastError() == cudaSuccess, kCuda, "dwpw_pixel_kernel launch failed");
    return;
  }
  // deep levels: [C][32 pixels] tile in shared memory, one 32-wide chunk of outputs per block (grid.y)
  const size_t smem = ((size_t)p.C * 32 + (size_t)p.C * kSepKC + (size_t)p.C * 10) * sizeof(float);
  OCRS_CHECK(smem <= 200 * 1024, kInternal, "separable conv: channel count too large for shared memory");
  if (smem > 48 * 1024)
    OCRS_CUDA_CHECK(cudaFuncSetAttribute(sepconv_kernel<32, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((unsigned)ceil_div(npix, 32), (unsigned)ceil_div(p.KO, kSepKC));
  sepconv_kernel<32, 1, 4><<<grid, 256, smem, st>>>(p);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}
}  // namespace

bool dwpw_supported(int C, int K) { return C >= 1 && C <= 1024 && K >= 1 && K <= 4096 && (size_t)C * (32 + kSepKC + 10) * 4 <= 200 * 1024; }

void dwpw_conv2(const SepInput* srcs, int n_src, const float* dw_w, const float* dw_b, const float* pw_w, const float* pw_b,
                float* y, int N, int Hv, int Wv, int K, int stride, int relu, cudaStream_t st) {
  SepParams p{};
  p.n_src = n_src;
  p.C = 0;
  for (int i = 0; i < n_src; ++i) {
    p.src[i] = SepSrc{srcs[i].x, srcs[i].C, srcs[i].H, srcs[i].W, srcs[i].pad};
    p.C += srcs[i].C;
  }
  p.N = N; p.Hv = Hv; p.Wv = Wv; p.stride = stride;
  p.OH = (Hv + 2 - 3) / stride + 1;
  p.OW = (Wv + 2 - 3) / stride + 1;
  if (N == 0 || p.OH <= 0 || p.OW <= 0) return;
  p.KO = K; p.relu = relu; p.mode = 0; p.w_ck = 0;
  p.dw_w = dw_w; p.dw_b = dw_b; p.pw_w = pw_w; p.pw_b = pw_b; p.y = y;
  launch_sep(p, st);
}

void dwpw_conv(const float* x, const float* dw_w, const float* dw_b, const float* pw_w, const float* pw_b, float* y,
               int N, int C, int H, int W, int K, int stride, int relu, cudaStream_t st) {
  SepInput s{x, C, H, W, 0.f};
  dwpw_conv2(&s, 1, dw_w, dw_b, pw_w, pw_b, y, N, H, W, K, stride, relu, st);
}

void conv_transpose_2x2s2(const float* x, const float* w, const float* b, float* y, int N, int C, int H, int W, int K,
                          int relu, cudaStream_t st) {
  if ((int64_t)N * H * W == 0 || !K) return;
  if (!dwpw_supported(C, 4 * K)) {  // very wide layers: the simple kernel
    static bool attr = false;
    size_t smem = (size_t)C * CT_KB * 4 * sizeof(float);
    if (!attr && smem > 48 * 1024) {
      OCRS_CUDA_CHECK(cudaFuncSetAttribute(convt2x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    dim3 grid((unsigned)ceil_div((int64_t)N * H * W, 128), (unsigned)ceil_div(K, CT_KB));
    convt2x2_kernel<<<grid, 128, smem, st>>>(x, w, b, y, N, C, H, W, K, relu);
    count_launch();
    OCRS_CUDA_CHECK(cudaGetLastError());
    return;
  }
  SepParams p{};
  p.n_src = 1;
  p.src[0] = SepSrc{x, C, H, W, 0.f};
  p.C = C; p.N = N; p.Hv = H; p.Wv = W; p.OH = H; p.OW = W; p.stride = 1;
  p.KO = 4 * K; p.relu = relu; p.mode = 1; p.w_ck = 1;
  p.pw_w = w; p.pw_b = b; p.y = y;
  launch_sep(p, st);
}

void conv_transpose_2x2s2_head(const float* x, const float* w, const float* b, const float* w2, const float* b2,
                               float* y, int N, int C, int H, int W, int Km, cudaStream_t st) {
  int64_t total = (int64_t)N * H * W;
  if (!total) return;
  OCRS_CHECK(C <= 16 && Km <= 16, kInternal, "conv_transpose_2x2s2_head: channel count > 16");
  convt2x2_head_kernel<<<(unsigned)ceil_div(total, 128), 128, 0, st>>>(x, w, b, w2, b2, y, N, C, H, W, Km);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void max_pool2d(const float* x, float* y, const PoolParams& p, cudaStream_t st) {
  int64_t total = (int64_t)p.NC * p.OH * p.OW;
  if (total == 0) return;
  pool_kernel<true><<<grid1d(total), kThreads, 0, st>>>(x, y, p);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}
void avg_pool2d(const float* x, float* y, const PoolParams& p, cudaStream_t st) {
  int64_t total = (int64_t)p.NC * p.OH * p.OW;
  if (total == 0) return;
  pool_kernel<false><<<grid1d(total), kThreads, 0, st>>>(x, y, p);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void relu(const float* x, float* y, int64_t n, cudaStream_t st) {
  if (!n) return;
  unary_kernel<Unary::kRelu><<<grid1d(n), kThreads, 0, st>>>(x, y, n);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}
void sigmoid(const float* x, float* y, int64_t n, cudaStream_t st) {
  if (!n) return;
  unary_kernel<Unary::kSigmoid><<<grid1d(n), kThreads, 0, st>>>(x, y, n);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}
void tanh_op(const float* x, float* y, int64_t n, cudaStream_t st) {
  if (!n) return;
  unary_kernel<Unary::kTanh><<<grid1d(n), kThreads, 0, st>>>(x, y, n);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}
void add_bcast_suffix(const float* a, const float* b, float* y, int64_t n, int64_t nb, cudaStream_t st) {
  if (!n) return;
  add_bcast_kernel<<<grid1d(n), kThreads, 0, st>>>(a, b, y, n, nb);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}
void fill(float* y, float v, int64_t n, cudaStream_t st) {
  if (!n) return;
  fill_kernel<<<grid1d(n), kThreads, 0, st>>>(y, v, n);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void sgemm_nt(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int relu,
              cudaStream_t st) {
  if (M == 0 || N == 0) return;
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
  sgemm_nt_kernel<<<grid, 256, 0, st>>>(A, B, bias, C, M, N, K, relu);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void permute(const float* x, float* y, const int64_t* shape, const int* perm, int ndim, cudaStream_t st) {
  OCRS_CHECK(ndim <= 6, kInternal, "permute: rank > 6");
  int64_t in_stride[6];
  int64_t s = 1;
  for (int d = ndim - 1; d >= 0; --d) {
    in_stride[d] = s;
    s *= shape[d];
  }
  if (s == 0) return;
  PermuteArgs a;
  a.ndim = ndim;
  for (int d = 0; d < ndim; ++d) {
    a.out_shape[d] = shape[perm[d]];
    a.in_stride[d] = in_stride[perm[d]];
  }
  permute_kernel<<<grid1d(s), kThreads, 0, st>>>(x, y, a, s);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void concat_copy(const float* x, float* y, int64_t outer, int64_t len_src, int64_t len_dst, int64_t dst_off,
                 int64_t inner, cudaStream_t st) {
  int64_t total = outer * len_src * inner;
  if (!total) return;
  concat_copy_kernel<<<grid1d(total), kThreads, 0, st>>>(x, y, outer, len_src, len_dst, dst_off, inner);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void pad4d(const float* x, float* y, const int64_t in_shape[4], const int64_t begin[4], const int64_t out_shape[4],
           float value, cudaStream_t st) {
  Pad4Args a;
  int64_t total = 1;
  for (int d = 0; d < 4; ++d) {
    a.in_shape[d] = in_shape[d];
    a.begin[d] = begin[d];
    a.out_shape[d] = out_shape[d];
    total *= out_shape[d];
  }
  if (!total) return;
  pad4d_kernel<<<grid1d(total), kThreads, 0, st>>>(x, y, a, value, total);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void log_softmax_lastdim(const float* x, float* y, int64_t rows, int cols, cudaStream_t st) {
  log_softmax_rows(x, cols, y, rows, cols, st);
}

void log_softmax_rows(const float* x, int64_t ldx, float* y, int64_t rows, int cols, cudaStream_t st) {
  if (!rows) return;
  log_softmax_kernel<<<grid1d(rows * 32), kThreads, 0, st>>>(x, ldx, y, rows, cols);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void gru_step(const float* xw, const float* R, const float* Rb, const float* h_in, float* h_out, float* Y, int D,
              int T, int N, int H, int step, const int* dir_reverse_host, int lbr, cudaStream_t st) {
  if (N == 0) return;
  int t0 = dir_reverse_host[0] ? T - 1 - step : step;
  int t1 = (D > 1) ? (dir_reverse_host[1] ? T - 1 - step : step) : 0;
  dim3 grid((unsigned)ceil_div(N, GL), (unsigned)ceil_div(H, GJ), (unsigned)D);
  gru_step_kernel<<<grid, 256, 0, st>>>(xw, R, Rb, h_in, h_out, Y, D, T, N, H, t0, t1, lbr);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nn
}  // namespace ocrs
