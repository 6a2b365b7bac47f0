// Tensor-core GRU: input-projection GEMM + persistent recurrent kernel (sm_100a).  See gru_tc.h.
#include "gru_tc.h"

#include "conv_tc.h"

#include <cuda_bf16.h>

#include <cstdlib>
#include <vector>

#include "tc_host.h"
#include "tc_ptx.cuh"

namespace ocrs {
namespace tc {

namespace {

using namespace ptx;

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// f32 -> split bf16 planes, 8 elements per thread
__global__ void split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                             int64_t n8) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(v[j], h0, l0);
    split_bf16(v[j + 1], h1, l1);
    ph[j / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    pl[j / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  reinterpret_cast<uint4*>(hi)[i] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  reinterpret_cast<uint4*>(lo)[i] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// ------------------------------------------------------------------------------------------
// C[M, Ntot] (f32) = A[M, K] * B[Ntot, K]^T + bias[Ntot]; A, B split bf16, K-major.
// Persistent, warp-specialised: one CTA per SM walks 128 x 128 tiles (the N tiles of one M tile are adjacent in
// the walk, so the A rows come from DRAM once and from L2 for the other N tiles).
//   warp 0   TMA producer: K chunks of 32 (64-byte-swizzled rows), 5 stages of 32 KB;
//   warp 1   MMA issuer: hi*hi + hi*lo + lo*hi into one of TWO tensor-memory accumulators (128 columns each), so
//            the MMAs of tile i+1 run while tile i is drained;
//   warps 2-9 epilogue (lane quadrant = warp % 4, column half = (warp - 2) / 4; with four warps the K = 128
//            projection was bound by the drain of its 64 KB tiles): tcgen05.ld 32 columns at a time, + bias, through a
//            128B-swizzled shared-memory staging
//            tile so that every global store instruction writes whole 128-byte lines (a thread owns a ROW of the
//            accumulator; storing straight from registers writes 16-byte pieces 6 KB apart).
// (The first form -- one tile per CTA, two CTAs per SM, the same four warps loading, issuing and storing in turn --
// ran the layer-1 / layer-2 input projections at 20% / 47% of the tensor peak: profiles/r02i_batch_ncu_table.md.)
// ------------------------------------------------------------------------------------------
// persistent grid: one CTA per SM of the current device
inline int gemm_ctas() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  OCRS_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    OCRS_CUDA_CHECK(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
    cached_dev = dev;
  }
  return cached;
}

constexpr int kGemmStages = 5;
constexpr int kGemmBK = 32;
constexpr int kGemmTile = 128 * kGemmBK * 2;            // one operand plane of one stage (8 KB)
constexpr int kGemmStageBytes = 4 * kGemmTile;          // A_hi, A_lo, B_hi, B_lo
constexpr int kGemmEpiWarps = 8;                        // two per TMEM lane quadrant, 64 columns each
constexpr int kGemmStagingBytes = 32 * 128;             // per epilogue warp: 32 rows x 32 floats
constexpr int kGemmThreads = 32 * (2 + kGemmEpiWarps);
constexpr int kGemmSmem = kGemmStages * kGemmStageBytes + kGemmEpiWarps * kGemmStagingBytes + 1024 + 256;

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const float* __restrict__ bias, float* __restrict__ Cout, int M, int Ntot, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging0 = base + kGemmStages * kGemmStageBytes;
  const uint32_t bar_base = staging0 + kGemmEpiWarps * kGemmStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kGemmStages + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * kGemmStages + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * kGemmStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kGemmStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), kGemmEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const int nkb = K / kGemmBK;
  const int n_tiles_n = Ntot / 128;
  const int n_tiles = ((M + 127) / 128) * n_tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles_n) * 128, n0 = (t % n_tiles_n) * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t st = base + s * kGemmStageBytes;
          mbar_expect_tx(full_bar(s), kGemmStageBytes);
          tma_load_2d(st, &tm_a_hi, kb * kGemmBK, m0, full_bar(s));
          tma_load_2d(st + kGemmTile, &tm_a_lo, kb * kGemmBK, m0, full_bar(s));
          tma_load_2d(st + 2 * kGemmTile, &tm_b_hi, kb * kGemmBK, n0, full_bar(s));
          tma_load_2d(st + 3 * kGemmTile, &tm_b_lo, kb * kGemmBK, n0, full_bar(s));
          if (++s == kGemmStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t s = 0, ph = 0, it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t a = it & 1;
        mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + a * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t st = base + s * kGemmStageBytes;
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            const uint32_t koff = k * 32;
            const uint64_t da_hi = make_desc<32>(st + koff), da_lo = make_desc<32>(st + kGemmTile + koff);
            const uint64_t db_hi = make_desc<32>(st + 2 * kGemmTile + koff), db_lo = make_desc<32>(st + 3 * kGemmTile + koff);
            umma_bf16(d, da_hi, db_hi, idesc, (kb | k) ? 1u : 0u);
            umma_bf16(d, da_hi, db_lo, idesc, 1u);
            umma_bf16(d, da_lo, db_hi, idesc, 1u);
          }
          umma_commit(empty_bar(s));
          if (++s == kGemmStages) { s = 0; ph ^= 1; }
        }
        umma_commit(acc_full(a));
      }
    }
  } else {
    // epilogue warp: TMEM lane quadrant q = warp % 4 (rows 32q .. 32q+31 of the tile), columns [64 * half, 64 * half + 64)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t stg = staging0 + (uint32_t)(warp - 2) * kGemmStagingBytes;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int m0 = (t / n_tiles_n) * 128, n0 = (t % n_tiles_n) * 128;
      const uint32_t a = it & 1;
      mbar_wait(acc_full(a), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 64 * half; c0 < 64 * half + 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * 128 + (uint32_t)c0, r);
        if (c0 == 64 * half + 32) {  // this warp's part of the accumulator is in registers: hand it back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(a));
        }
        // row `lane` of the warp's 32 x 32 block -> staging, 16-byte chunk c at (c ^ (lane & 7))
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + 4 * c));
          const float4 v = make_float4(__uint_as_float(r[4 * c]) + b4.x, __uint_as_float(r[4 * c + 1]) + b4.y,
                                       __uint_as_float(r[4 * c + 2]) + b4.z, __uint_as_float(r[4 * c + 3]) + b4.w);
          const uint32_t addr = stg + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
        __syncwarp();
        // 8 lanes per row, 4 rows per instruction: every row segment is one full 128-byte line
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3), c = lane & 7;
          const uint32_t addr = stg + (uint32_t)rr * 128u + (uint32_t)((c ^ (rr & 7)) << 4);
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
          const int row = m0 + q * 32 + rr;
          if (row < M) *reinterpret_cast<float4*>(Cout + (size_t)row * Ntot + n0 + c0 + 4 * c) = v;
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------
// Cluster-resident recurrent kernel.  A cluster of 8 CTAs owns one tile of <= 32 lines of one
// direction for all timesteps.  CTA rank r keeps the R rows of hidden units [32r, 32r+32) (3 gates
// x 32 rows, split bf16) resident in TENSOR MEMORY as the MMA A operand, so nothing is streamed per
// step; after each step every CTA writes its 32-unit slice of h_t (split bf16, MMA B-operand layout)
// into the shared memory of all 8 CTAs (DSMEM) and signals their mbarriers.
// Lines are described individually (ragged): own length, own row strides for xw and Y.
// ------------------------------------------------------------------------------------------

constexpr int NL = 32;                               // lines per cluster tile (= MMA N)
constexpr int kHSub = NL * 128;                      // one 64-wide K sub-tile of h (bytes)
constexpr int kHPlane = 4 * kHSub;                   // h plane (hi or lo): [NL][256] bf16

constexpr int kCl = 8;                                 // CTAs per cluster
constexpr int kHBuf = 2 * kHPlane;                     // one h buffer (hi + lo) = 32 KB
constexpr int kExBytes = 3 * NL * 32 * 4;              // gate pre-activation exchange [3][NL][32] f32
constexpr int kGateWarps = 16;                         // 4 line groups x 4 TMEM lane quarters
constexpr int kGateThreads = kGateWarps * 32;
constexpr int kStageBf = kGateWarps * 256;             // per warp: [hi|lo][2 lines][32 units] bf16
constexpr int kClThreads = kGateThreads + 64;          // + TMA/init warp + MMA warp
constexpr int kClSmem = 2 * kHBuf + kExBytes + kStageBf + 1024 + 256;

// fast gate functions for the latency-critical recurrence: MUFU exp + approximate divide
// (|error| ~1e-7, far inside the 1e-3 log-prob budget; saturate correctly for large |x|)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void reg_fence8(uint32_t (&r)[8]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
}

__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kClThreads, 1)
gru_cluster_kernel(const __nv_bfloat16* __restrict__ r_hi_g, const __nv_bfloat16* __restrict__ r_lo_g,
                   const float* __restrict__ xw, const float* __restrict__ rb, const float* __restrict__ h0,
                   float* __restrict__ Y, float* __restrict__ Yh, const SeqLine* __restrict__ lines, int n_lines,
                   int D, int64_t y_dstride, int rev0, int rev1, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t base = (smem0 + 1023u) & ~1023u;
  const uint32_t hbuf0 = base;                         // buffer b at hbuf0 + b*kHBuf: [hi plane | lo plane]
  const uint32_t ex = hbuf0 + 2 * kHBuf;
  const uint32_t stage = ex + kExBytes;
  const uint32_t bar_base = stage + kStageBf;
  const uint32_t d_full_bar = bar_base + 8;
  const uint32_t h_ready_bar0 = bar_base + 16;         // [2]
  const uint32_t tmem_slot = bar_base + 32;
  float* exf = reinterpret_cast<float*>(smem_raw + (ex - smem0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x / kCl;
  const int d = blockIdx.y;
  const int rev = d == 0 ? rev0 : rev1;
  const SeqLine* tl = lines + (size_t)tile * NL;
  const int n_in_tile = min(NL, n_lines - tile * NL);
  const int steps = tl[0].T;                           // lines are sorted by T descending inside a tile

  if (threadIdx.x == 0) {
    mbar_init(d_full_bar, 1);
    mbar_init(h_ready_bar0, kCl);
    mbar_init(h_ready_bar0 + 8, kCl);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);  // D: cols [0,NL); A_hi: [64,192); A_lo: [192,320)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    // ---- one-time: this CTA's R slice becomes the MMA A operand in TENSOR MEMORY.  TMEM lane = row
    // (gate g = warp, unit = lane; lanes 96..127 are zero padding), columns = packed bf16 pairs along K.
    const uint4* src_hi = nullptr;
    const uint4* src_lo = nullptr;
    if (warp < 3) {
      const size_t row = (size_t)d * 768 + (size_t)warp * 256 + (size_t)rank * 32 + lane;
      src_hi = reinterpret_cast<const uint4*>(r_hi_g + row * 256);
      src_lo = reinterpret_cast<const uint4*>(r_lo_g + row * 256);
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int plane = 0; plane < 2; ++plane) {
      const uint4* src = plane == 0 ? src_hi : src_lo;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {  // 4 x 32 columns = 128 words = 256 bf16
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 v = src ? __ldg(src + c * 8 + j) : make_uint4(0, 0, 0, 0);
          r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
        }
        tmem_st32(lane_base + (uint32_t)(64 + plane * 128 + c * 32), r);
      }
    }
    tmem_wait_st();
  }
  if (warp < kGateWarps) {
    // ---- h_{-1}: every CTA fills its own copy of buffer 0 (all 256 units) ----
    for (int e = threadIdx.x; e < NL * 256; e += kGateThreads) {
      const int l = e >> 8, u = e & 255;
      float v = 0.f;
      if (h0 != nullptr && l < n_in_tile && tl[l].valid) v = h0[((size_t)d * n_lines + tile * NL + l) * 256 + u];
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      const int kk = u >> 6, col = u & 63;
      const uint32_t off = (uint32_t)(kk * kHSub + (l >> 3) * 1024 + (l & 7) * 128 + (((col >> 3) ^ (l & 7)) << 4) + (col & 7) * 2);
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(hbuf0 + off), "h"(__bfloat16_as_ushort(hi)) : "memory");
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(hbuf0 + kHPlane + off), "h"(__bfloat16_as_ushort(lo)) : "memory");
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers are initialised before anyone arrives remotely

  if (warp == kGateWarps + 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NL >> 3) << 17) | ((128u >> 4) << 24);
      tc_fence_after();
      long long m_wait = 0, m_issue = 0;
      for (int step = 0; step < steps; ++step) {
        const int b = step & 1;
        const long long c0 = clock64();
        if (step > 0) {
          mbar_wait(h_ready_bar0 + 8 * b, ((step - 1) >> 1) & 1);  // relaxed spin ...
          fence_acq_rel_cluster();                                 // ... then one cluster-scope acquire
        }
        tc_fence_after();
        const long long c1 = clock64();
        m_wait += c1 - c0;
        const uint32_t hb = hbuf0 + b * kHBuf;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t koff = k * 32;
            const uint32_t ta_hi = tmem_base + 64u + (uint32_t)((kk * 4 + k) * 8), ta_lo = ta_hi + 128u;
            const uint64_t db_hi = make_desc<64>(hb + kk * kHSub + koff), db_lo = make_desc<64>(hb + kHPlane + kk * kHSub + koff);
            umma_bf16_ts(tmem_base, ta_hi, db_hi, idesc, (kk | k) ? 1u : 0u);
            umma_bf16_ts(tmem_base, ta_hi, db_lo, idesc, 1u);
            umma_bf16_ts(tmem_base, ta_lo, db_hi, idesc, 1u);
          }
        }
        umma_commit(d_full_bar);
        m_issue += clock64() - c1;
      }
      if (dbg && blockIdx.x == 0 && blockIdx.y == 0) { dbg[0] = m_wait; dbg[1] = m_issue; dbg[7] = steps; }
    }
  } else if (warp < kGateWarps) {
    // ---------------- gate math: 16 warps ----------------
    // phase A: warp (q, g) = (warp / 4, warp % 4), g < 3, drains gate g of lines 8q..8q+7 (TMEM lanes 32g.. = units)
    // phase B: warp w owns lines 2w, 2w+1 for unit = lane of this CTA's slice
    const int q = warp >> 2, g = warp & 3;
    const int unit = (int)rank * 32 + lane;
    const int l0 = 2 * warp;
    float h[2];
    int lT[2];
    int64_t lx[2], lxs[2], ly[2], lys[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int l = l0 + i;
      const bool ok = l < n_in_tile && tl[l].valid;
      h[i] = (h0 != nullptr && ok) ? h0[((size_t)d * n_lines + tile * NL + l) * 256 + unit] : 0.f;
      lT[i] = ok ? tl[l].T : 0;
      lx[i] = ok ? tl[l].xw_base : 0;
      lxs[i] = ok ? tl[l].xw_tstride : 0;
      ly[i] = ok ? tl[l].y_base : 0;
      lys[i] = ok ? tl[l].y_tstride : 0;
    }
    const float rbz = rb[(size_t)d * 768 + unit], rbr = rb[(size_t)d * 768 + 256 + unit], rbn = rb[(size_t)d * 768 + 512 + unit];
    __nv_bfloat16* stg = reinterpret_cast<__nv_bfloat16*>(smem_raw + (stage - smem0)) + warp * 128;  // [hi: 2x32][lo: 2x32]
    const uint32_t my_stage = stage + warp * 256;
    long long e_wait = 0, e_drain = 0, e_math = 0, e_xchg = 0, e_sig = 0;
    for (int step = 0; step < steps; ++step) {
      const int b = step & 1, nb = b ^ 1;
      const long long t0 = clock64();
      // prefetch the input projections of this step (overlaps the MMAs)
      float xz[2], xr[2], xn[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        xz[i] = xr[i] = xn[i] = 0.f;
        if (step < lT[i]) {
          const int t = rev ? (lT[i] - 1 - step) : step;
          const float* xg = xw + lx[i] + (int64_t)t * lxs[i] + (int64_t)d * 768 + unit;
          xz[i] = __ldg(xg);
          xr[i] = __ldg(xg + 256);
          xn[i] = __ldg(xg + 512);
        }
      }
      mbar_wait(d_full_bar, step & 1);
      tc_fence_after();
      const long long t1 = clock64();
      e_wait += t1 - t0;
      if (g < 3) {
        uint32_t acc[8];
        tmem_ld8(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(q * 8), acc);
#pragma unroll
        for (int j = 0; j < 8; ++j) exf[(g * NL + q * 8 + j) * 32 + lane] = __uint_as_float(acc[j]);
      }
      tc_fence_before();
      named_bar_sync(1, kGateThreads);
      const long long t2 = clock64();
      e_drain += t2 - t1;
      {
        float pz[2], pr[2], pn[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          pz[i] = exf[(0 * NL + l0 + i) * 32 + lane];
          pr[i] = exf[(1 * NL + l0 + i) * 32 + lane];
          pn[i] = exf[(2 * NL + l0 + i) * 32 + lane];
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const bool act = step < lT[i];
          const float z = fast_sigmoid(xz[i] + pz[i] + rbz);
          const float r = fast_sigmoid(xr[i] + pr[i] + rbr);
          const float nn_ = fast_tanh(xn[i] + r * (pn[i] + rbn));
          const float hn = act ? ((1.f - z) * nn_ + z * h[i]) : h[i];
          h[i] = hn;
          if (act) {
            const int t = rev ? (lT[i] - 1 - step) : step;
            Y[ly[i] + (int64_t)t * lys[i] + (int64_t)d * y_dstride + unit] = hn;
          }
          __nv_bfloat16 hi, lo;
          split_bf16(hn, hi, lo);
          stg[i * 32 + lane] = hi;
          stg[64 + i * 32 + lane] = lo;
        }
      }
      __syncwarp();
      const long long t3 = clock64();
      e_math += t3 - t2;
      if (step + 1 < steps) {
        if (lane < 16) {
          // lane -> (plane p, line i, 16-byte chunk c = units 8c..8c+7 of my slice)
          const int c = lane & 3, i = (lane >> 2) & 1, p = lane >> 3;
          const int l = l0 + i;
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(my_stage + (uint32_t)(p * 128 + i * 64 + c * 16)));
          const int kk = (int)rank >> 1;
          const int chunk = (((int)rank & 1) * 4 + c) ^ (l & 7);
          const uint32_t dst = hbuf0 + nb * kHBuf + p * kHPlane +
                               (uint32_t)(kk * kHSub + (l >> 3) * 1024 + (l & 7) * 128 + chunk * 16);
#pragma unroll
          for (int pc = 0; pc < kCl; ++pc) st_cluster_v4(mapa(dst, (uint32_t)pc), v);
          fence_proxy_async_all();
        }
      }
      named_bar_sync(1, kGateThreads);
      const long long t4 = clock64();
      e_xchg += t4 - t3;
      if (step + 1 < steps && threadIdx.x == 0) {
        fence_acq_rel_cluster();  // one cluster-scope release for the whole CTA's DSMEM writes
#pragma unroll
        for (int pc = 0; pc < kCl; ++pc) mbar_arrive_remote_relaxed(mapa(h_ready_bar0 + 8 * nb, (uint32_t)pc));
      }
      e_sig += clock64() - t4;
    }
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
      dbg[2] = e_wait; dbg[3] = e_drain; dbg[4] = e_math; dbg[5] = e_xchg; dbg[6] = e_sig;
    }
    if (Yh != nullptr) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int l = l0 + i;
        if (l < n_in_tile && tl[l].valid) Yh[((size_t)d * n_lines + tile * NL + l) * 256 + unit] = h[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody exits while peers may still write into its shared memory
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// v3 of the cluster-resident recurrent kernel.  Same decomposition as above (8 CTAs per tile of <= 32
// lines, R slice resident in tensor memory), re-cut around what the micro-benchmarks measured
// (profiles/r02a_ubench_*.log):
//   * a single thread issues a tcgen05.mma every ~53 cycles whatever its size; four issuing warps get the
//     48 MMAs of a step through in ~940 cycles instead of ~2550.  Warp w owns the K range of the hidden
//     units of CTAs 2w and 2w+1 and accumulates into its own TMEM accumulator; the gate warps add the four
//     partial sums;
//   * the all-gather of h_t is DSMEM-bandwidth bound (~20 B/clk/SM): it cannot be made cheap, but it can
//     be overlapped.  h lives in shared memory as 8 mini tiles [32 lines][32 units] (64-byte-swizzled
//     K-major, hi | lo planes adjacent = 4 KB contiguous per source CTA); every CTA pushes its tile to the
//     7 peers with ONE cp.async.bulk (shared::cta -> shared::cluster) each, completing on the DESTINATION's
//     per-slice mbarrier, and the issuing warps start the MMAs of a slice as soon as that slice has landed;
//   * no fences or remote arrives on the critical path: complete_tx on the peer's barrier is the signal.
// Protocol per step t (buffer b = t & 1 holds h_{t-1}):  MMA warp w, for s in {2w, 2w+1}: wait
// slice_bar[b][s] -> 6 MMAs;  commit -> d_full (4 arrivals).  Gate warps: drain + add the 4 accumulators,
// gate math, write h_t (hi, lo) into their own slice of buffer b^1, fence.proxy.async, barrier; one thread
// arrives on its own slice barrier and issues the 7 bulk copies.  Buffer b^1 was last read by the MMAs of
// step t-1, which every peer has finished before it could produce the h_{t-1} this CTA consumed.
// ------------------------------------------------------------------------------------------
constexpr int kSliceTile = NL * 64;            // one plane of one slice: [NL][32] bf16, 64-byte rows = 2 KB
constexpr int kSliceBoth = 2 * kSliceTile;     // hi | lo
constexpr int kHBuf3 = kCl * kSliceBoth;       // 32 KB
constexpr int kMmaWarps3 = 4;
constexpr int kGruRegsDefault = 96;
constexpr int kCl3Threads = kGateThreads + 32 * kMmaWarps3;
constexpr int kCl3Smem = 2 * kHBuf3 + kExBytes + 1024 + 512;

// byte offset of (line l, unit u in [0, 32)) inside a 64-byte-swizzled [NL][32] bf16 tile
__device__ __forceinline__ uint32_t sw64_off(int l, int u) {
  return (uint32_t)(l * 64 + ((((u >> 3) ^ (l >> 1)) & 3) << 4) + (u & 7) * 2);
}

// Register cap as a template parameter (OCRS_B200_GRU_REGS = 64 | 72 | 80 | 96, default 96).  The idea behind the
// smaller caps -- 640 threads x 96 registers leave no room on the SM for the light kernels of the other batches in
// flight, while the recurrence keeps its 112 SMs mostly idle -- did not pay: each step got slower (3.85 k -> 4.3 k
// cycles at 64 registers) and the step time of the pipeline did not improve (7.74 / 7.71 / 7.75 / 7.82 ms per step at
// 96 / 80 / 72 / 64 registers with three batches in flight, 7.34 / 7.36 / 7.44 / 7.51 with four).  With ten batches in
// flight the capped variants showed occasional timed regions 1.5-5x longer than the rest (light kernels of other
// batches sharing the recurrence's SMs), the 96-register kernel none: profiles/r02aa_bench_r*.json.
template <int kRegs>
__global__ void __cluster_dims__(kCl, 1, 1) __maxnreg__(kRegs)
gru_cluster3_kernel(const __nv_bfloat16* __restrict__ r_hi_g, const __nv_bfloat16* __restrict__ r_lo_g,
                    const float* __restrict__ xw, const float* __restrict__ rb, const float* __restrict__ h0,
                    float* __restrict__ Y, float* __restrict__ Yh, const SeqLine* __restrict__ lines, int n_lines,
                    int D, int64_t y_dstride, int rev0, int rev1, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t base = (smem0 + 1023u) & ~1023u;
  const uint32_t hbuf0 = base;                         // buffer b at hbuf0 + b*kHBuf3; slice s at + s*kSliceBoth: [hi | lo]
  const uint32_t ex = hbuf0 + 2 * kHBuf3;
  const uint32_t bar_base = ex + kExBytes;
  const uint32_t d_full_bar = bar_base;
  auto slice_bar = [&](int b, int sl) { return bar_base + 16u + 8u * (uint32_t)(b * kCl + sl); };
  const uint32_t tmem_slot = bar_base + 16u + 8u * 2 * kCl;
  float* exf = reinterpret_cast<float*>(smem_raw + (ex - smem0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x / kCl;
  const int d = blockIdx.y;
  const int rev = d == 0 ? rev0 : rev1;
  const SeqLine* tl = lines + (size_t)tile * NL;
  const int n_in_tile = min(NL, n_lines - tile * NL);
  const int steps = tl[0].T;                           // lines are sorted by T descending inside a tile

  if (threadIdx.x == 0) {
    mbar_init(d_full_bar, kMmaWarps3);
    for (int b = 0; b < 2; ++b)
      for (int sl = 0; sl < kCl; ++sl) mbar_init(slice_bar(b, sl), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);  // D_w: cols [32w, 32w+32); A_hi: [128,256); A_lo: [256,384)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    // ---- one-time: this CTA's R slice becomes the MMA A operand in TENSOR MEMORY.  TMEM lane = row
    // (gate g = warp, unit = lane; lanes 96..127 are zero padding), columns = packed bf16 pairs along K.
    const uint4* src_hi = nullptr;
    const uint4* src_lo = nullptr;
    if (warp < 3) {
      const size_t row = (size_t)d * 768 + (size_t)warp * 256 + (size_t)rank * 32 + lane;
      src_hi = reinterpret_cast<const uint4*>(r_hi_g + row * 256);
      src_lo = reinterpret_cast<const uint4*>(r_lo_g + row * 256);
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int plane = 0; plane < 2; ++plane) {
      const uint4* src = plane == 0 ? src_hi : src_lo;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {  // 4 x 32 columns = 128 words = 256 bf16
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 v = src ? __ldg(src + c * 8 + j) : make_uint4(0, 0, 0, 0);
          r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
        }
        tmem_st32(lane_base + (uint32_t)(128 + plane * 128 + c * 32), r);
      }
    }
    tmem_wait_st();
  }
  if (warp < kGateWarps) {
    // ---- h_{-1}: every CTA fills its own copy of buffer 0 (all 256 units) ----
    for (int e = threadIdx.x; e < NL * 256; e += kGateThreads) {
      const int l = e >> 8, u = e & 255;
      float v = 0.f;
      if (h0 != nullptr && l < n_in_tile && tl[l].valid) v = h0[((size_t)d * n_lines + tile * NL + l) * 256 + u];
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      const uint32_t off = hbuf0 + (uint32_t)((u >> 5) * kSliceBoth) + sw64_off(l, u & 31);
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(off), "h"(__bfloat16_as_ushort(hi)) : "memory");
      asm volatile("st.shared.b16 [%0], %1;" ::"r"(off + kSliceTile), "h"(__bfloat16_as_ushort(lo)) : "memory");
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers are initialised before anyone's bulk copies complete on them

  if (warp >= kGateWarps) {
    // ---------------- MMA issuers: warp w handles the K range of slices 2w, 2w+1 ----------------
    const int w = warp - kGateWarps;
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NL >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t d_acc = tmem_base + (uint32_t)(w * NL);
      tc_fence_after();
      long long m_issue = 0;
      for (int step = 0; step < steps; ++step) {
        const int b = step & 1;
        const uint32_t hb = hbuf0 + b * kHBuf3;
        const long long c0 = clock64();
        // the own slice's barrier is arrived on after this CTA's gate warps have drained the accumulators of
        // the previous step: every issuing warp waits for it before it overwrites its accumulator (a warp
        // whose two slices are remote would otherwise depend on the peers' progress only)
        if (step > 0) mbar_wait_trap(slice_bar(b, (int)rank), ((step - 1) >> 1) & 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int sl = 2 * w + i;
          if (step > 0) {
            // remote slices arrive as bulk-copy bytes (arm the transaction count); the own slice as a plain arrive
            if (sl != (int)rank) mbar_expect_tx(slice_bar(b, sl), kSliceBoth);
            mbar_wait_trap(slice_bar(b, sl), ((step - 1) >> 1) & 1);
            tc_fence_after();
          }
          const uint64_t db0 = make_desc<32>(hb + sl * kSliceBoth);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t ta_hi = tmem_base + 128u + (uint32_t)((sl * 2 + k) * 8), ta_lo = ta_hi + 128u;
            const uint64_t db_hi = db0 + (uint64_t)(2 * k), db_lo = db_hi + (uint64_t)(kSliceTile >> 4);
            umma_bf16_ts(d_acc, ta_hi, db_hi, idesc, (i | k) ? 1u : 0u);
            umma_bf16_ts(d_acc, ta_hi, db_lo, idesc, 1u);
            umma_bf16_ts(d_acc, ta_lo, db_hi, idesc, 1u);
          }
        }
        umma_commit(d_full_bar);
        m_issue += clock64() - c0;
      }
      if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && w == 0) { dbg[0] = 0; dbg[1] = m_issue; dbg[7] = steps; }
    }
  } else {
    // ---------------- gate math: 16 warps ----------------
    // phase A: warp (q, g) = (warp / 4, warp % 4), g < 3, drains gate g of lines 8q..8q+7 (TMEM lanes 32g.. = units),
    //          adding the four K-partial accumulators
    // phase B: warp w owns lines 2w, 2w+1 for unit = lane of this CTA's slice
    const int q = warp >> 2, g = warp & 3;
    const int unit = (int)rank * 32 + lane;
    const int l0 = 2 * warp;
    float h[2];
    int lT[2];
    int64_t lx[2], lxs[2], ly[2], lys[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int l = l0 + i;
      const bool ok = l < n_in_tile && tl[l].valid;
      h[i] = (h0 != nullptr && ok) ? h0[((size_t)d * n_lines + tile * NL + l) * 256 + unit] : 0.f;
      lT[i] = ok ? tl[l].T : 0;
      lx[i] = ok ? tl[l].xw_base : 0;
      lxs[i] = ok ? tl[l].xw_tstride : 0;
      ly[i] = ok ? tl[l].y_base : 0;
      lys[i] = ok ? tl[l].y_tstride : 0;
    }
    const float rbz = rb[(size_t)d * 768 + unit], rbr = rb[(size_t)d * 768 + 256 + unit], rbn = rb[(size_t)d * 768 + 512 + unit];
    long long e_wait = 0, e_drain = 0, e_math = 0, e_xchg = 0;
    for (int step = 0; step < steps; ++step) {
      const int b = step & 1, nb = b ^ 1;
      const long long t0 = clock64();
      // prefetch the input projections of this step (overlaps the MMAs)
      float xz[2], xr[2], xn[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        xz[i] = xr[i] = xn[i] = 0.f;
        if (step < lT[i]) {
          const int t = rev ? (lT[i] - 1 - step) : step;
          const float* xg = xw + lx[i] + (int64_t)t * lxs[i] + (int64_t)d * 768 + unit;
          xz[i] = __ldg(xg);
          xr[i] = __ldg(xg + 256);
          xn[i] = __ldg(xg + 512);
        }
      }
      mbar_wait_trap(d_full_bar, step & 1);
      tc_fence_after();
      const long long t1 = clock64();
      e_wait += t1 - t0;
      if (g < 3) {
        uint32_t a0[8], a1[8], a2[8], a3[8];
        const uint32_t ta = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(q * 8);
        tmem_ld8_issue(ta, a0);
        tmem_ld8_issue(ta + NL, a1);
        tmem_ld8_issue(ta + 2 * NL, a2);
        tmem_ld8_issue(ta + 3 * NL, a3);
        tmem_wait_ld();
        reg_fence8(a0); reg_fence8(a1); reg_fence8(a2); reg_fence8(a3);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          exf[(g * NL + q * 8 + j) * 32 + lane] =
              (__uint_as_float(a0[j]) + __uint_as_float(a1[j])) + (__uint_as_float(a2[j]) + __uint_as_float(a3[j]));
      }
      tc_fence_before();
      named_bar_sync(1, kGateThreads);
      const long long t2 = clock64();
      e_drain += t2 - t1;
      {
        float pz[2], pr[2], pn[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          pz[i] = exf[(0 * NL + l0 + i) * 32 + lane];
          pr[i] = exf[(1 * NL + l0 + i) * 32 + lane];
          pn[i] = exf[(2 * NL + l0 + i) * 32 + lane];
        }
        const uint32_t own = hbuf0 + nb * kHBuf3 + rank * kSliceBoth;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const bool act = step < lT[i];
          const float z = fast_sigmoid(xz[i] + pz[i] + rbz);
          const float r = fast_sigmoid(xr[i] + pr[i] + rbr);
          const float nn_ = fast_tanh(xn[i] + r * (pn[i] + rbn));
          const float hn = act ? ((1.f - z) * nn_ + z * h[i]) : h[i];
          h[i] = hn;
          if (act) {
            const int t = rev ? (lT[i] - 1 - step) : step;
            Y[ly[i] + (int64_t)t * lys[i] + (int64_t)d * y_dstride + unit] = hn;
          }
          if (step + 1 < steps) {
            __nv_bfloat16 hi, lo;
            split_bf16(hn, hi, lo);
            const uint32_t off = own + sw64_off(l0 + i, lane);
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(off), "h"(__bfloat16_as_ushort(hi)) : "memory");
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(off + kSliceTile), "h"(__bfloat16_as_ushort(lo)) : "memory");
          }
        }
      }
      const long long t3 = clock64();
      e_math += t3 - t2;
      if (step + 1 < steps) {
        fence_proxy_async();                 // generic-proxy writes of h_t -> visible to the bulk copies and the MMAs
        named_bar_sync(1, kGateThreads);
        if (threadIdx.x < kCl) {
          const uint32_t own = hbuf0 + nb * kHBuf3 + rank * kSliceBoth;
          const uint32_t pc = (rank + threadIdx.x) & (kCl - 1);   // thread 0: own CTA; 1..7: the peers, staggered
          if (pc == rank) {
            mbar_arrive(slice_bar(nb, (int)rank));
          } else {
            asm volatile(
                "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mapa(own, pc)),
                "r"(own), "r"((uint32_t)kSliceBoth), "r"(mapa(slice_bar(nb, (int)rank), pc))
                : "memory");
          }
        }
      }
      e_xchg += clock64() - t3;
    }
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
      dbg[2] = e_wait; dbg[3] = e_drain; dbg[4] = e_math; dbg[5] = e_xchg; dbg[6] = 0;
    }
    if (Yh != nullptr) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int l = l0 + i;
        if (l < n_in_tile && tl[l].valid) Yh[((size_t)d * n_lines + tile * NL + l) * 256 + unit] = h[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody exits while peers' bulk copies may still target its shared memory
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// launch of the recurrent kernel: v3 by default, OCRS_B200_GRU_V2=1 selects the previous kernel
void launch_recurrence(const GruWeightsTC& w, const float* xw, const float* h0, float* Y, float* Yh, const SeqLine* d_desc,
                       int n_lines, int n_tiles, int64_t y_dstride, const int* reverse, long long* d_dbg, cudaStream_t st) {
  static const bool v2 = [] { const char* e = std::getenv("OCRS_B200_GRU_V2"); return e != nullptr && e[0] == '1'; }();
  dim3 grid((unsigned)(n_tiles * kCl), (unsigned)w.D);
  if (v2) {
    OCRS_CUDA_CHECK(cudaFuncSetAttribute(gru_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kClSmem));
    gru_cluster_kernel<<<grid, kClThreads, kClSmem, st>>>(w.r_hi.as<__nv_bfloat16>(), w.r_lo.as<__nv_bfloat16>(), xw, w.rb.as<float>(),
                                                          h0, Y, Yh, d_desc, n_lines, w.D, y_dstride, reverse[0],
                                                          w.D > 1 ? reverse[1] : 0, d_dbg);
  } else {
    // register cap of the recurrent kernel (see its comment); OCRS_B200_GRU_REGS = 64 | 72 | 80 | 96 for experiments
    static const int regs = [] { const char* e = std::getenv("OCRS_B200_GRU_REGS"); return e ? std::atoi(e) : kGruRegsDefault; }();
    auto launch = [&](auto kern) {
      OCRS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kCl3Smem));
      kern<<<grid, kCl3Threads, kCl3Smem, st>>>(w.r_hi.as<__nv_bfloat16>(), w.r_lo.as<__nv_bfloat16>(), xw, w.rb.as<float>(), h0, Y, Yh,
                                               d_desc, n_lines, w.D, y_dstride, reverse[0], w.D > 1 ? reverse[1] : 0, d_dbg);
    };
    if (regs == 64) launch(gru_cluster3_kernel<64>);
    else if (regs == 72) launch(gru_cluster3_kernel<72>);
    else if (regs == 80) launch(gru_cluster3_kernel<80>);
    else launch(gru_cluster3_kernel<96>);
  }
  count_launch();
}

}  // namespace

bool gru_supported(int D, int H, int I) { return available() && H == 256 && (D == 1 || D == 2) && I % 64 == 0 && I >= 64; }

std::unique_ptr<GruWeightsTC> prepare_gru(const float* W, const float* R, const float* B, int D, int H, int I) {
  auto g = std::make_unique<GruWeightsTC>();
  g->D = D; g->H = H; g->I = I;
  auto split_upload = [](const float* src, size_t n, DeviceBuffer& hi, DeviceBuffer& lo) {
    std::vector<__nv_bfloat16> h(n), l(n);
    for (size_t i = 0; i < n; ++i) {
      h[i] = __float2bfloat16_rn(src[i]);
      l[i] = __float2bfloat16_rn(src[i] - __bfloat162float(h[i]));
    }
    hi.reserve(n * 2);
    lo.reserve(n * 2);
    OCRS_CUDA_CHECK(cudaMemcpy(hi.ptr, h.data(), n * 2, cudaMemcpyHostToDevice));
    OCRS_CUDA_CHECK(cudaMemcpy(lo.ptr, l.data(), n * 2, cudaMemcpyHostToDevice));
  };
  split_upload(W, (size_t)D * 3 * H * I, g->w_hi, g->w_lo);
  split_upload(R, (size_t)D * 3 * H * H, g->r_hi, g->r_lo);
  std::vector<float> wb((size_t)D * 3 * H, 0.f), rb((size_t)D * 3 * H, 0.f);
  if (B) {
    for (int d = 0; d < D; ++d) {
      std::copy(B + (size_t)d * 6 * H, B + (size_t)d * 6 * H + 3 * H, wb.begin() + (size_t)d * 3 * H);
      std::copy(B + (size_t)d * 6 * H + 3 * H, B + (size_t)(d + 1) * 6 * H, rb.begin() + (size_t)d * 3 * H);
    }
  }
  g->wb.reserve(wb.size() * 4);
  g->rb.reserve(rb.size() * 4);
  OCRS_CUDA_CHECK(cudaMemcpy(g->wb.ptr, wb.data(), wb.size() * 4, cudaMemcpyHostToDevice));
  OCRS_CUDA_CHECK(cudaMemcpy(g->rb.ptr, rb.data(), rb.size() * 4, cudaMemcpyHostToDevice));
  return g;
}

namespace {
// xw[M][Ntot] = A W^T + bias through gemm_tc_kernel; A given as f32 [M][K] (split here)
void gemm_split_bf16(const float* X, int64_t M, int K, const void* w_hi, const void* w_lo, const float* bias, int Ntot, float* out,
                     const ScratchAlloc& alloc, cudaStream_t st) {
  auto* x_hi = static_cast<__nv_bfloat16*>(alloc((size_t)M * K * 2));
  auto* x_lo = static_cast<__nv_bfloat16*>(alloc((size_t)M * K * 2));
  const int64_t n8 = M * K / 8;
  split_kernel<<<(unsigned)ceil_div(n8, 256), 256, 0, st>>>(X, x_hi, x_lo, n8);
  count_launch();
  uint64_t ad[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t as[1] = {(uint64_t)K * 2};
  uint32_t box[2] = {(uint32_t)kGemmBK, 128};
  CUtensorMap ta_hi = make_map(x_hi, 2, ad, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
  CUtensorMap ta_lo = make_map(x_lo, 2, ad, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
  uint64_t bd[2] = {(uint64_t)K, (uint64_t)Ntot};
  CUtensorMap tb_hi = make_map(w_hi, 2, bd, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
  CUtensorMap tb_lo = make_map(w_lo, 2, bd, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
  OCRS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(M, 128) * (Ntot / 128), gemm_ctas()));
  gemm_tc_kernel<<<grid, kGemmThreads, kGemmSmem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, bias, out, (int)M, Ntot, K);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}
}  // namespace

bool linear_supported(int N, int K) { return available() && N >= 1 && K >= 32 && K % 32 == 0; }

std::unique_ptr<LinearWeightsTC> prepare_linear(const float* Wt, const float* bias, int N, int K) {
  auto l = std::make_unique<LinearWeightsTC>();
  l->N = N; l->K = K; l->Npad = (int)round_up(N, 128);
  std::vector<__nv_bfloat16> h((size_t)l->Npad * K, __float2bfloat16_rn(0.f)), lo((size_t)l->Npad * K, __float2bfloat16_rn(0.f));
  for (size_t i = 0; i < (size_t)N * K; ++i) {
    h[i] = __float2bfloat16_rn(Wt[i]);
    lo[i] = __float2bfloat16_rn(Wt[i] - __bfloat162float(h[i]));
  }
  std::vector<float> b((size_t)l->Npad, 0.f);
  if (bias) std::copy(bias, bias + N, b.begin());
  l->w_hi.reserve(h.size() * 2);
  l->w_lo.reserve(lo.size() * 2);
  l->bias.reserve(b.size() * 4);
  OCRS_CUDA_CHECK(cudaMemcpy(l->w_hi.ptr, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  OCRS_CUDA_CHECK(cudaMemcpy(l->w_lo.ptr, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
  OCRS_CUDA_CHECK(cudaMemcpy(l->bias.ptr, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  return l;
}

void linear_forward(const float* X, int64_t rows, const LinearWeightsTC& w, float* Y, const ScratchAlloc& alloc,
                    cudaStream_t st) {
  if (rows == 0) return;
  gemm_split_bf16(X, rows, w.K, w.w_hi.ptr, w.w_lo.ptr, w.bias.as<float>(), w.Npad, Y, alloc, st);
}

void gru_forward(const float* X, const GruWeightsTC& w, const float* h0, float* Y, float* Yh, int T, int N,
                 const int* reverse, const ScratchAlloc& alloc, cudaStream_t st) {
  if (T == 0 || N == 0) return;
  const int D = w.D, H = w.H, I = w.I;
  const int64_t M = (int64_t)T * N;
  const int Ntot = D * 3 * H;
  // (1) split X
  auto* x_hi = static_cast<__nv_bfloat16*>(alloc((size_t)M * I * 2));
  auto* x_lo = static_cast<__nv_bfloat16*>(alloc((size_t)M * I * 2));
  const int64_t n8 = M * I / 8;
  split_kernel<<<(unsigned)ceil_div(n8, 256), 256, 0, st>>>(X, x_hi, x_lo, n8);
  count_launch();
  // (2) xw[M][D*3H] = X W^T + Wb
  float* xw = static_cast<float*>(alloc((size_t)M * Ntot * 4));
  {
    uint64_t ad[2] = {(uint64_t)I, (uint64_t)M};
    uint64_t as[1] = {(uint64_t)I * 2};
    uint32_t box[2] = {(uint32_t)kGemmBK, 128};
    CUtensorMap ta_hi = make_map(x_hi, 2, ad, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
    CUtensorMap ta_lo = make_map(x_lo, 2, ad, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
    uint64_t bd[2] = {(uint64_t)I, (uint64_t)Ntot};
    CUtensorMap tb_hi = make_map(w.w_hi.ptr, 2, bd, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
    CUtensorMap tb_lo = make_map(w.w_lo.ptr, 2, bd, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
    OCRS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(M, 128) * (Ntot / 128), gemm_ctas()));
    gemm_tc_kernel<<<grid, kGemmThreads, kGemmSmem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, w.wb.as<float>(), xw, (int)M, Ntot, I);
    count_launch();
  }
  // (3) recurrence
  {
    const int n_tiles = (int)ceil_div(N, NL);
    std::vector<SeqLine> desc((size_t)n_tiles * NL);
    for (int i = 0; i < n_tiles * NL; ++i) {
      SeqLine& L = desc[(size_t)i];
      L.valid = i < N ? 1 : 0;
      L.T = i < N ? T : 0;
      L.xw_base = (int64_t)i * Ntot;
      L.xw_tstride = (int64_t)N * Ntot;
      L.y_base = (int64_t)i * H;
      L.y_tstride = (int64_t)D * N * H;
    }
    auto* d_desc = static_cast<SeqLine*>(alloc(desc.size() * sizeof(SeqLine)));
    OCRS_CUDA_CHECK(cudaMemcpyAsync(d_desc, desc.data(), desc.size() * sizeof(SeqLine), cudaMemcpyHostToDevice, st));
    launch_recurrence(w, xw, h0, Y, Yh, d_desc, N, n_tiles, (int64_t)N * H, reverse, nullptr, st);
  }
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void gru_forward_lines(const float* X, int64_t rows, const GruWeightsTC& w, const SeqLine* lines_host, int n_lines,
                       float* Y, int64_t y_dstride, const int* reverse, const ScratchAlloc& alloc, cudaStream_t st,
                       const std::function<void()>* after_projection) {
  if (rows == 0 || n_lines == 0) return;
  const int D = w.D, H = w.H, I = w.I;
  const int Ntot = D * 3 * H;
  auto* x_hi = static_cast<__nv_bfloat16*>(alloc((size_t)rows * I * 2));
  auto* x_lo = static_cast<__nv_bfloat16*>(alloc((size_t)rows * I * 2));
  const int64_t n8 = rows * I / 8;
  split_kernel<<<(unsigned)ceil_div(n8, 256), 256, 0, st>>>(X, x_hi, x_lo, n8);
  count_launch();
  float* xw = static_cast<float*>(alloc((size_t)rows * Ntot * 4));
  {
    uint64_t ad[2] = {(uint64_t)I, (uint64_t)rows};
    uint64_t as[1] = {(uint64_t)I * 2};
    uint32_t box[2] = {(uint32_t)kGemmBK, 128};
    CUtensorMap ta_hi = make_map(x_hi, 2, ad, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
    CUtensorMap ta_lo = make_map(x_lo, 2, ad, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
    uint64_t bd[2] = {(uint64_t)I, (uint64_t)Ntot};
    CUtensorMap tb_hi = make_map(w.w_hi.ptr, 2, bd, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
    CUtensorMap tb_lo = make_map(w.w_lo.ptr, 2, bd, as, box, CU_TENSOR_MAP_SWIZZLE_64B);
    OCRS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(rows, 128) * (Ntot / 128), gemm_ctas()));
    gemm_tc_kernel<<<grid, kGemmThreads, kGemmSmem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, w.wb.as<float>(), xw, (int)rows, Ntot, I);
    count_launch();
  }
  if (after_projection != nullptr && *after_projection) (*after_projection)();
  const int n_tiles = (int)ceil_div(n_lines, NL);
  std::vector<SeqLine> desc((size_t)n_tiles * NL);
  for (int i = 0; i < n_tiles * NL; ++i) {
    if (i < n_lines) desc[(size_t)i] = lines_host[i];
    else desc[(size_t)i] = SeqLine{0, 0, 0, 0, 0, 0};
  }
  auto* d_desc = static_cast<SeqLine*>(alloc(desc.size() * sizeof(SeqLine)));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(d_desc, desc.data(), desc.size() * sizeof(SeqLine), cudaMemcpyHostToDevice, st));
  static const bool dbg_on = std::getenv("OCRS_B200_GRU_DEBUG") != nullptr;
  long long* d_dbg = nullptr;
  if (dbg_on) {
    d_dbg = static_cast<long long*>(alloc(8 * sizeof(long long)));
    OCRS_CUDA_CHECK(cudaMemsetAsync(d_dbg, 0, 8 * sizeof(long long), st));
  }
  launch_recurrence(w, xw, nullptr, Y, nullptr, d_desc, n_lines, n_tiles, y_dstride, reverse, d_dbg, st);
  if (dbg_on) {
    long long hdbg[8];
    OCRS_CUDA_CHECK(cudaMemcpyAsync(hdbg, d_dbg, sizeof(hdbg), cudaMemcpyDeviceToHost, st));
    OCRS_CUDA_CHECK(cudaStreamSynchronize(st));
    double n = (double)std::max<long long>(hdbg[7], 1);
    fprintf(stderr,
            "[gru dbg] steps %lld | MMA thread: wait h_ready %.0f, issue+commit %.0f | gate warp0: wait d_full %.0f, drain %.0f, "
            "math %.0f, exchange+fence %.0f, signal %.0f (cycles/step)\n",
            hdbg[7], hdbg[0] / n, hdbg[1] / n, hdbg[2] / n, hdbg[3] / n, hdbg[4] / n, hdbg[5] / n, hdbg[6] / n);
  }
  OCRS_CUDA_CHECK(cudaGetLastError());
}

}  // namespace tc
}  // namespace ocrs
