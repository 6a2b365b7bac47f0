// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution (sm_100a).  See conv_tc.h.
//
// Tile: 128 output pixels (8 rows x 16 columns of one image) x COUT channels; persistent CTAs
// (one per SM) walk the tiles.
// K loop: 9 taps x (Cin / KC) channel chunks; per k-block TMA loads
//   A_hi, A_lo : [128 pixels][KC] fp16, box {KC, TW, TH, 1} of the NHWC activation at
//                (c0, w0 + kw - 1, h0 + kh - 1, n) -- the halo / zero padding is TMA out-of-bounds fill
//   B_hi, B_lo : [COUT][KC] fp16 from the [COUT][9*Cin] weight matrix
// into 128B- (KC = 64) or 64B- (KC = 32) swizzled shared memory; one elected thread issues
//   HH += A_hi*B_hi ; X += A_hi*B_lo ; X += A_lo*B_hi       (tcgen05.mma kind::f16, fp32 accumulate in TMEM)
// and eight warps (two per TMEM lane quadrant, half of the output channels each) promote HH to
// registers every 128 K-elements, then add X, bias, ReLU, optional max-pool, split to fp16 hi/lo and
// store NHWC (see the kernel comment for the numerics).  Eight rather than four because the MMA
// thread was measured waiting 25 % of its time for the promotion (experiments/README.md).
#include "conv_tc.h"

#include "tc_host.h"
#include "tc_ptx.cuh"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace ocrs {
namespace tc {

namespace {

using namespace ptx;

// ---- split-fp16 helpers: x = hi + lo, both fp16 (22 significant bits; |x| must stay below 65504:
// larger values raise the model's overflow flag and the executor reruns on the fp32 kernels) ----
__device__ __forceinline__ void split1(float v, uint16_t& hi, uint16_t& lo, int* ovf) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
  if (!(fabsf(v) <= 65504.f)) *ovf = 1;  // also catches NaN
}
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo, int* ovf) {
  uint16_t h0, l0, h1, l1;
  split1(v0, h0, l0, ovf);
  split1(v1, h1, l1, ovf);
  hi = (uint32_t)h0 | ((uint32_t)h1 << 16);
  lo = (uint32_t)l0 | ((uint32_t)l1 << 16);
}
// packed form for the conv epilogue: two cvt.rn.f16x2.f32 instead of four scalar conversions (the scalar
// F2F.F16.F32 issues at a quarter of the ALU rate).  Overflow = the hi half came out as inf / NaN.
// Returns non-zero when a value left the fp16 range (the caller ORs these and raises the flag once).
__device__ __forceinline__ uint32_t split2_packed(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
  const uint32_t t = hi & 0x7C007C00u;  // exponent fields; all ones <=> inf / NaN
  return (t + 0x04000400u) & 0x80008000u;
}
__device__ __forceinline__ float join_lo(uint32_t hi, uint32_t lo) {
  return __half2float(__ushort_as_half((uint16_t)(hi & 0xFFFFu))) + __half2float(__ushort_as_half((uint16_t)(lo & 0xFFFFu)));
}
__device__ __forceinline__ float join_hi(uint32_t hi, uint32_t lo) {
  return __half2float(__ushort_as_half((uint16_t)(hi >> 16))) + __half2float(__ushort_as_half((uint16_t)(lo >> 16)));
}

template <int KC, int COUT>
struct Cfg {
  static constexpr int kABytes = 128 * KC * 2;
  static constexpr int kBBytes = COUT * KC * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStagesRaw = (192 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kGroupKb = 128 / KC;      // k-blocks per promotion group (128 K-elements)
  static constexpr int kBarBytes = 512;          // mbarriers + TMEM slot + scheduler ring
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + kBarBytes;
  static constexpr int kTmemCols = 4 * COUT;     // HH[2] | X[2]
  // promotion + epilogue warps: one per (TMEM lane quadrant, 32 output channels) -- 16 for COUT = 128, 8 for 64.
  // (Eight warps of 64 channels each were measured falling behind the MMAs at every tile boundary: the MMA thread
  // then waits for an accumulator buffer, profiles/r02b_phase_timers.log.)
  static constexpr int kEpiWarps = 4 * (COUT / 32);
  static constexpr int kThreads = 32 * (kEpiWarps + 2);  // + TMA / scheduler warp + MMA warp
  static_assert(kStages > kGroupKb, "a promotion group must fit in the stage ring with room to prefetch");
};

constexpr int kTW = 16, kTH = 8;   // pixel tile: 8 rows x 16 columns = 128 TMEM lanes, lane = th*16 + tw
constexpr int kSched = 4;          // depth of the tile ring between the scheduler and the other roles

// One entry of the tile ring (written by the scheduler thread, read by the MMA thread and the
// eight promotion warps): which tile, and what the epilogue needs to know about its group.
struct TileEntry {
  int32_t g, n, h0, w0;       // group (-1 = no more tiles), image, tile origin
  int32_t H, W, OH, OW;       // input / pooled output dims of the group
  int64_t out_off;            // first output pixel of the group
  int64_t pad;
};
static_assert(sizeof(TileEntry) == 48, "TileEntry layout");

// In-register transpose of 16-byte chunks inside groups of G lanes (G = 4): x holds G chunks of 4 words;
// afterwards chunk q of lane r (r = lane % G) is what chunk r of lane (lane - r + q) was.
template <int G>
__device__ __forceinline__ void transpose_chunks(uint32_t (&x)[4 * G], int lane) {
#pragma unroll
  for (int m = 1; m < G; m <<= 1) {
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int c = 0; c < G; ++c) {
      if (c & m) continue;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t a = x[4 * c + w], b = x[4 * (c | m) + w];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, up ? a : b, m);
        if (up) x[4 * c + w] = recv;
        else x[4 * (c | m) + w] = recv;
      }
    }
  }
}

// Epilogue of one tile for a promotion warp (32 pixels x 32 channels per warp): bias, ReLU, max-pool by warp
// shuffles, fp16 split; the four 16-byte chunks of a thread are then transposed inside groups of 4 lanes so that
// lane r of a group holds chunk r of each of the group's 4 pixels -- a store instruction writes 64 contiguous
// bytes per pixel instead of 32 scattered 16-byte pieces.  (Measured with the in-kernel timers: the scattered
// stores cost ~2.5 cycles per lane, 7.7-10.6k cycles per tile, during which no accumulator is promoted and the
// MMA thread runs out of buffers -- profiles/r02g_conv_promotion_timers.log.)
template <int COUT>
__device__ __forceinline__ void epilogue_store(const float (&acc)[32], const TileEntry& e, int wq, int hsel, int lane, int relu,
                                               int ph, int pw, bool real_tile, const float* __restrict__ bias,
                                               act_t* __restrict__ out_hi, act_t* __restrict__ out_lo, int* __restrict__ ovf) {
  constexpr int CH = 32, kRowsPerWarp = 32 / kTW;
  // pooling factors are 1 or 2: divisions become shifts (an integer division costs ~20 issue slots, and the
  // epilogue is issue-bound: profiles/r02i_conv_sass_stalls.md)
  const int sh = ph >> 1, sw = pw >> 1;
  const int th = wq * kRowsPerWarp + (lane >> 4), tw = lane & (kTW - 1);
  const int oh = (e.h0 + th) >> sh, ow = (e.w0 + tw) >> sw;
  const bool writer = real_tile && (ph == 1 || (lane & kTW) == 0) && (pw == 1 || (lane & 1) == 0) && oh < e.OH && ow < e.OW;
  uint32_t hp[16], lp[16];
  uint32_t bad = 0u;  // branch-free: every lane splits its values, only writers' overflow counts
#pragma unroll
  for (int c0 = 0; c0 < CH; c0 += 8) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + hsel * CH + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + hsel * CH + c0 + 4));
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float v0 = acc[c0 + j] + bb[j];
      float v1 = acc[c0 + j + 1] + bb[j + 1];
      if (relu) {
        v0 = fmaxf(v0, 0.f);
        v1 = fmaxf(v1, 0.f);
      }
      if (ph == 2) {
        v0 = fmaxf(v0, __shfl_xor_sync(0xffffffffu, v0, kTW));
        v1 = fmaxf(v1, __shfl_xor_sync(0xffffffffu, v1, kTW));
      }
      if (pw == 2) {
        v0 = fmaxf(v0, __shfl_xor_sync(0xffffffffu, v0, 1));
        v1 = fmaxf(v1, __shfl_xor_sync(0xffffffffu, v1, 1));
      }
      const int k = (c0 + j) / 2;
      bad |= split2_packed(v0, v1, hp[k], lp[k]);
    }
  }
  if (writer && bad) *ovf = 1;
  transpose_chunks<4>(hp, lane);
  transpose_chunks<4>(lp, lane);
  const uint32_t wmask = __ballot_sync(0xffffffffu, writer);
  const int r = lane & 3;
  // the 4 pixels of this lane's group share a tile row; their output pixels differ only in the column
  const int w4 = e.w0 + (lane & (kTW - 4));
  const size_t row_pix = (size_t)e.out_off + ((size_t)e.n * e.OH + oh) * e.OW;
  act_t* const row_hi = out_hi + row_pix * COUT + hsel * CH + r * 8;
  act_t* const row_lo = out_lo + row_pix * COUT + hsel * CH + r * 8;
  const uint32_t wm4 = wmask >> (lane & ~3);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if ((wm4 >> q) & 1u) {  // lane (lane & ~3) + q writes: this lane holds chunk r of its pixel
      const int off = ((w4 + q) >> sw) * COUT;
      *reinterpret_cast<uint4*>(row_hi + off) = make_uint4(hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]);
      *reinterpret_cast<uint4*>(row_lo + off) = make_uint4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]);
    }
  }
}

__device__ __forceinline__ void fence_tensormap_acquire(const void* p) {
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(p) : "memory");
}

// Numerics.  tcgen05.mma truncates its fp32 accumulator toward zero after every instruction
// (measured: -0.5 ulp of the running sum per MMA, tools/diag_tc_rounding.py), so a K = 1152
// dot product issued as 216 MMAs into one accumulator ends ~100 ulp low.  Therefore:
//   * the cross terms (hi*lo + lo*hi, 2/3 of the MMAs, 2^-11 of the magnitude) get their own
//     accumulator X, whose truncation is negligible in absolute terms;
//   * the hi*hi partial sums are promoted to fp32 registers (round-to-nearest adds on the CUDA
//     cores) every 128 K-elements, double-buffered in TMEM so the tensor pipe never waits.
// The same warps then run the epilogue (bias, ReLU, optional 2x2 / 2x1 max-pool through warp
// shuffles, split to fp16 hi/lo, NHWC store) while the MMA warp is already on the next tile.
//
// Work distribution.  ONE launch covers every width group of a layer ("ragged"): the groups are
// separate NHWC tensors, each with its own pair of TMA tensor maps (hi, lo) in a device array, and
// the persistent CTAs draw 8x16-pixel tiles from a global counter (dynamic scheduling: a CTA that
// starts late -- its SM was busy with another kernel -- simply takes fewer tiles).  The scheduler is
// the TMA thread; it publishes each tile through a small shared-memory ring.
//
// Issue order inside a promotion group (128 K-elements = kGroupKb k-blocks): first ALL hi*hi MMAs of
// the group, commit -> the promotion warps start draining that accumulator while the tensor pipe
// works through the group's 2x as many cross-term MMAs.  The arithmetic (order of accumulation into
// either accumulator) is the same as issuing them interleaved.
//
// dbg (optional, OCRS_B200_CONV_DEBUG=1): per-launch sums over all CTAs of the MMA thread's phases
// [0] tiles, [1] k-blocks, [2] wait tile ring, [3] wait x_empty, [4] wait hh_empty, [5] wait operands
// (HH phase), [6] issue HH, [7] issue X + commits, [8] total cycles in the tile loop.
template <int KC, int COUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((Cfg<KC, COUT>::kThreads), 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                  const CUtensorMap* __restrict__ maps, const RaggedDesc* __restrict__ groups, int n_groups,
                  int n_tiles, int* __restrict__ counter, const float* __restrict__ bias, act_t* __restrict__ out_hi,
                  act_t* __restrict__ out_lo, int Cin, int relu, int ph, int pw, float promo_scale,
                  int* __restrict__ ovf, unsigned long long* __restrict__ dbg, int hh_first) {
  using C = Cfg<KC, COUT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t base = (smem0 + 1023u) & ~1023u;
  const uint32_t bar_base = base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t aux = bar_base + 8u * (2 * C::kStages);
  auto hh_full = [&](int b) { return aux + 8u * b; };
  auto hh_empty = [&](int b) { return aux + 8u * (2 + b); };
  auto x_full = [&](int b) { return aux + 8u * (4 + b); };
  auto x_empty = [&](int b) { return aux + 8u * (6 + b); };
  auto sched_full = [&](int s) { return aux + 8u * (8 + s); };
  auto sched_empty = [&](int s) { return aux + 8u * (8 + kSched + s); };
  const uint32_t tmem_slot = aux + 8u * (8 + 2 * kSched);
  const uint32_t ring = tmem_slot + 16u;  // kSched x TileEntry (16-byte aligned)
  static_assert(8 * (2 * 8 + 8 + 2 * kSched) + 16 + kSched * (int)sizeof(TileEntry) <= C::kBarBytes, "barrier area");
  TileEntry* ring_p = reinterpret_cast<TileEntry*>(smem_raw + (ring - smem0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(hh_full(b), 1);
      mbar_init(hh_empty(b), C::kEpiWarps);
      mbar_init(x_full(b), 1);
      mbar_init(x_empty(b), C::kEpiWarps);
    }
    for (int s = 0; s < kSched; ++s) {
      mbar_init(sched_full(s), 1);
      mbar_init(sched_empty(s), C::kEpiWarps + 1);  // MMA thread + the promotion warps
    }
    fence_barrier_init();
  }
  if (warp == C::kEpiWarps) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int chunks = Cin / KC;
  const int nkb = 9 * chunks;

  if (warp == C::kEpiWarps) {
    if (lane == 0) {
      // ---------------- tile scheduler + TMA producer ----------------
      uint32_t stage = 0, par = 0;
      int cur_g = 0, last_map_g = -1;
      RaggedDesc gd = groups[0];
      int t = atomicAdd(counter, 1);
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        mbar_wait(sched_empty(slot), ((ti / kSched) & 1) ^ 1);
        TileEntry e;
        e.g = -1;
        e.n = e.h0 = e.w0 = e.H = e.W = e.OH = e.OW = 0;
        e.out_off = 0;
        e.pad = 0;
        if (t < n_tiles) {
          while (cur_g + 1 < n_groups && __ldg(&groups[cur_g + 1].first) <= t) {
            ++cur_g;
            gd = groups[cur_g];
          }
          int l = t - gd.first;
          e.g = cur_g;
          e.w0 = (l % gd.tiles_w) * kTW;
          l /= gd.tiles_w;
          e.h0 = (l % gd.tiles_h) * kTH;
          e.n = l / gd.tiles_h;
          e.H = gd.H; e.W = gd.W; e.OH = gd.OH; e.OW = gd.OW;
          e.out_off = gd.out_off;
        }
        ring_p[slot] = e;
        mbar_arrive(sched_full(slot));  // release: the entry is visible to whoever sees the phase flip
        if (e.g < 0) break;
        const int t_next = atomicAdd(counter, 1);  // in flight while this tile's loads are issued
        const CUtensorMap* m_hi = maps + 2 * e.g;
        const CUtensorMap* m_lo = m_hi + 1;
        if (e.g != last_map_g) {
          fence_tensormap_acquire(m_hi);
          fence_tensormap_acquire(m_lo);
          last_map_g = e.g;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), par ^ 1);
          const uint32_t st = base + stage * C::kStageBytes;
          mbar_expect_tx(full_bar(stage), C::kStageBytes);
          const int tap = kb / chunks, c0 = (kb - tap * chunks) * KC;
          const int kh = tap / 3, kw = tap - kh * 3;
          tma_load_4d(st, m_hi, c0, e.w0 + kw - 1, e.h0 + kh - 1, e.n, full_bar(stage));
          tma_load_4d(st + C::kABytes, m_lo, c0, e.w0 + kw - 1, e.h0 + kh - 1, e.n, full_bar(stage));
          tma_load_2d(st + 2 * C::kABytes, &tm_w_hi, tap * Cin + c0, 0, full_bar(stage));
          tma_load_2d(st + 2 * C::kABytes + C::kBBytes, &tm_w_lo, tap * Cin + c0, 0, full_bar(stage));
          if (++stage == C::kStages) { stage = 0; par ^= 1; }
        }
        t = t_next;
      }
    }
  } else if (warp == C::kEpiWarps + 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      // c_format F32 (bit 4), a/b format F16 (0 at bits 7, 10), N >> 3 at 17, M >> 4 at 24
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t stage = 0, par = 0, gc = 0;
      unsigned long long c_sched = 0, c_xe = 0, c_hhe = 0, c_full = 0, c_hh = 0, c_x = 0, n_tiles_done = 0;
      const bool timing = dbg != nullptr;
      const long long t_begin = timing ? clock64() : 0;
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        long long c0 = timing ? clock64() : 0;
        mbar_wait(sched_full(slot), (ti / kSched) & 1);
        int g;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(g) : "r"(ring + slot * (uint32_t)sizeof(TileEntry)) : "memory");
        mbar_arrive(sched_empty(slot));
        if (g < 0) break;
        const uint32_t tp = ti & 1;
        const uint32_t d_x = tmem_base + 2 * COUT + tp * COUT;
        long long c1 = timing ? clock64() : 0;
        mbar_wait(x_empty(tp), ((ti >> 1) & 1) ^ 1);
        long long c2 = timing ? clock64() : 0;
        c_sched += (unsigned long long)(c1 - c0);
        c_xe += (unsigned long long)(c2 - c1);
        for (int kb = 0; kb < nkb; ++gc) {
          const uint32_t b = gc & 1;
          const uint32_t d_hh = tmem_base + b * COUT;
          const int nk = min(C::kGroupKb, nkb - kb);
          long long g0 = timing ? clock64() : 0;
          mbar_wait(hh_empty(b), ((gc >> 1) & 1) ^ 1);
          tc_fence_after();
          long long g1 = timing ? clock64() : 0;
          c_hhe += (unsigned long long)(g1 - g0);
          uint32_t s = stage, p = par;
          long long g2 = g1;
          if (hh_first) {
          // ---- hi*hi MMAs of the whole group ----
          for (int j = 0; j < nk; ++j) {
            long long w0 = timing ? clock64() : 0;
            mbar_wait(full_bar(s), p);
            tc_fence_after();
            if (timing) c_full += (unsigned long long)(clock64() - w0);
            // one descriptor per stage; the other operands and the K steps are byte offsets >> 4 added to its
            // 14-bit start-address field (shared memory ends below 256 KB, so the adds never carry out of it)
            const uint64_t d0 = make_desc<KC>(base + s * C::kStageBytes);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              const uint64_t da_hi = d0 + (uint64_t)(2 * k);
              const uint64_t db_hi = da_hi + (uint64_t)((2 * C::kABytes) >> 4);
              umma_bf16(d_hh, da_hi, db_hi, idesc, (j | k) ? 1u : 0u);
            }
            if (++s == C::kStages) { s = 0; p ^= 1; }
          }
          umma_commit(hh_full(b));
          g2 = timing ? clock64() : 0;
          // ---- cross terms of the group; each k-block's stage is released behind its last MMA ----
          s = stage;
          for (int j = 0; j < nk; ++j) {
            const uint64_t d0 = make_desc<KC>(base + s * C::kStageBytes);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              const uint64_t da_hi = d0 + (uint64_t)(2 * k), da_lo = da_hi + (uint64_t)(C::kABytes >> 4);
              const uint64_t db_hi = da_hi + (uint64_t)((2 * C::kABytes) >> 4), db_lo = db_hi + (uint64_t)(C::kBBytes >> 4);
              umma_bf16(d_x, da_hi, db_lo, idesc, (kb | j | k) ? 1u : 0u);
              umma_bf16(d_x, da_lo, db_hi, idesc, 1u);
            }
            umma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
            if (++s == C::kStages) s = 0;
          }
          } else {
          // ---- interleaved: per 16 K-elements hi*hi, hi*lo, lo*hi (consecutive MMAs share an operand) ----
          for (int j = 0; j < nk; ++j) {
            long long w0 = timing ? clock64() : 0;
            mbar_wait(full_bar(s), p);
            tc_fence_after();
            if (timing) c_full += (unsigned long long)(clock64() - w0);
            const uint64_t d0 = make_desc<KC>(base + s * C::kStageBytes);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              const uint64_t da_hi = d0 + (uint64_t)(2 * k), da_lo = da_hi + (uint64_t)(C::kABytes >> 4);
              const uint64_t db_hi = da_hi + (uint64_t)((2 * C::kABytes) >> 4), db_lo = db_hi + (uint64_t)(C::kBBytes >> 4);
              umma_bf16(d_hh, da_hi, db_hi, idesc, (j | k) ? 1u : 0u);
              umma_bf16(d_x, da_hi, db_lo, idesc, (kb | j | k) ? 1u : 0u);
              umma_bf16(d_x, da_lo, db_hi, idesc, 1u);
            }
            umma_commit(empty_bar(s));
            if (++s == C::kStages) { s = 0; p ^= 1; }
          }
          umma_commit(hh_full(b));
          g2 = timing ? clock64() : 0;
          }
          if (timing) {
            const long long g3 = clock64();
            c_hh += (unsigned long long)(g2 - g1);
            c_x += (unsigned long long)(g3 - g2);
          }
          stage = s;
          par = p;
          kb += nk;
        }
        umma_commit(x_full(tp));
        ++n_tiles_done;
      }
      if (timing) {
        atomicAdd(dbg + 0, n_tiles_done);
        atomicAdd(dbg + 1, n_tiles_done * (unsigned long long)nkb);
        atomicAdd(dbg + 2, c_sched);
        atomicAdd(dbg + 3, c_xe);
        atomicAdd(dbg + 4, c_hhe);
        atomicAdd(dbg + 5, c_full);
        atomicAdd(dbg + 6, c_hh - c_full);
        atomicAdd(dbg + 7, c_x);
        atomicAdd(dbg + 8, (unsigned long long)(clock64() - t_begin));
        atomicAdd(dbg + 9, 1ull);
      }
    }
  } else {
    // ---------------- promotion + epilogue: warps 0 .. kEpiWarps-1 ----------------
    // warp & 3 = TMEM lane quadrant (a warp may only touch lanes 32*(warp % 4)..+31), warp >> 2 = which
    // 32 output channels; each thread keeps 32 partial sums.
    constexpr int CH = 32;
    const int wq = warp & 3, hsel = warp >> 2;
    const uint32_t lane_base = ((uint32_t)(wq * 32) << 16) + (uint32_t)(hsel * CH);
    const int ngroups = (nkb + C::kGroupKb - 1) / C::kGroupKb;
    uint32_t gc = 0;
    const bool etime = dbg != nullptr && warp == 0 && lane == 0;
    unsigned long long e_ring = 0, e_hhf = 0, e_promo = 0, e_xf = 0, e_epi = 0;
    for (uint32_t ti = 0;; ++ti) {
      const uint32_t slot = ti & (kSched - 1);
      long long q0 = etime ? clock64() : 0;
      mbar_wait(sched_full(slot), (ti / kSched) & 1);
      if (etime) e_ring += (unsigned long long)(clock64() - q0);
      const TileEntry e = ring_p[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(sched_empty(slot));
      if (e.g < 0) break;
      float acc[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = 0.f;
      for (int g = 0; g < ngroups; ++g, ++gc) {
        const uint32_t b = gc & 1;
        long long q1 = etime ? clock64() : 0;
        mbar_wait(hh_full(b), (gc >> 1) & 1);
        tc_fence_after();
        long long q2 = etime ? clock64() : 0;
        if constexpr (CH == 64) {
          // both TMEM loads of the group in flight before the single wait
          uint32_t r0[32], r1[32];
          tmem_ld32_issue(tmem_base + lane_base + b * COUT, r0);
          tmem_ld32_issue(tmem_base + lane_base + b * COUT + 32u, r1);
          tmem_wait_ld();
          reg_fence32(r0);
          reg_fence32(r1);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(__uint_as_float(r0[j]), promo_scale, acc[j]);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[32 + j] = fmaf(__uint_as_float(r1[j]), promo_scale, acc[32 + j]);
        } else {
#pragma unroll
          for (int c0 = 0; c0 < CH; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + b * COUT + (uint32_t)c0, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c0 + j] = fmaf(__uint_as_float(r[j]), promo_scale, acc[c0 + j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(hh_empty(b));
        if (etime) {
          e_hhf += (unsigned long long)(q2 - q1);
          e_promo += (unsigned long long)(clock64() - q2);
        }
      }
      const uint32_t tp = ti & 1;
      long long q3 = etime ? clock64() : 0;
      mbar_wait(x_full(tp), (ti >> 1) & 1);
      tc_fence_after();
      long long q4 = etime ? clock64() : 0;
#pragma unroll
      for (int c0 = 0; c0 < CH; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + 2 * COUT + tp * COUT + (uint32_t)c0, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty(tp));

      epilogue_store<COUT>(acc, e, wq, hsel, lane, relu, ph, pw, true, bias, out_hi, out_lo, ovf);
      if (etime) {
        e_xf += (unsigned long long)(q4 - q3);
        e_epi += (unsigned long long)(clock64() - q4);
      }
    }
    if (etime) {
      atomicAdd(dbg + 10, e_ring);
      atomicAdd(dbg + 11, e_hhf);
      atomicAdd(dbg + 12, e_promo);
      atomicAdd(dbg + 13, e_xf);
      atomicAdd(dbg + 14, e_epi);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == C::kEpiWarps) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// Small-layer form (C_in = KC: one channel chunk, all 9 weight tiles fit in shared memory): the 32 -> 64 layer.
// The single-CTA kernel above streams 24 KB per k-block for 192 cycles of tensor work and sits at the L2 -> SM
// delivery rate.  Here
//   * the WEIGHTS are loaded once per CTA and stay resident (9 taps x (hi, lo) x C_out x KC x 2 B = 72 KB);
//   * the activation is loaded as one 10-row box per horizontal tap kw (rows h0-1 .. h0+8) and the three vertical
//     taps read it through descriptors advanced by whole tile rows, as in the CTA-pair + halo kernel below:
//     60 KB per tile instead of 216 KB;
//   * three hi*hi accumulators (see the halo kernel) and one accumulator per cross term, each of the three products
//     of the split issued by its own thread.
// k-blocks run in (kw, kh) order; promotion groups stay 128 K-elements (4 k-blocks of 32 channels).
// Measured (profiles/r02t_conv_timers.log): ~6.0 k cycles per tile, of which the issuing threads wait < 0.7 k: the 54
// MMAs of a tile take ~100 cycles each -- 6 KB of shared-memory operands per MMA of N = 64 in the 64-byte-swizzled
// layout (32 channels = 64-byte rows) are delivered at ~60 B/clk, about half of what the 128-byte-swizzled operands
// of the 128-channel layers get.  This layer is bound by shared-memory operand delivery, not by issue or by TMA.
// Since `conv3x3_ws_kernel` (below) this kernel only runs the layer when its pooling is not 2x2, or with
// OCRS_B200_CONV_WS=0.
// ------------------------------------------------------------------------------------------
template <int KC_, int COUT>
struct ResCfg {
  static constexpr int kKC = KC_;
  static constexpr int kABox = (kTH + 2) * kTW * kKC * 2;     // one plane of the 10 x 16 pixel box
  static constexpr int kStageBytes = 2 * kABox;                // hi + lo
  static constexpr int kStages = 4;
  static constexpr int kBTile = COUT * kKC * 2;                // one plane of one tap's weights
  static constexpr int kBBytes = 9 * 2 * kBTile;               // resident weights
  static constexpr int kGroupKb = 128 / kKC;
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBBytes + 1024 + kBarBytes;
  static constexpr int kTmemCols = 8 * COUT;      // HH[3] | X1[2] | X2[2] (7 x COUT, rounded up to a power of two)
  static constexpr int kEpiWarps = 4 * (COUT / 32);
  static constexpr int kMmaWarps = 3;              // hi*hi, hi*lo, lo*hi: one issuing thread each
  static constexpr int kThreads = 32 * (kEpiWarps + 1 + kMmaWarps);
  static_assert(kTmemCols <= 512, "tensor memory");
  static_assert(kABox % 1024 == 0 && kBTile % 1024 == 0, "operand tiles must keep the swizzle phase");
  static_assert(kSmemBytes <= 227 * 1024, "resident weights do not fit");
};

template <int KC_, int COUT>
__global__ void __launch_bounds__((ResCfg<KC_, COUT>::kThreads), 1)
conv3x3_res_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                   const CUtensorMap* __restrict__ maps, const RaggedDesc* __restrict__ groups, int n_groups, int n_tiles,
                   int* __restrict__ counter, const float* __restrict__ bias, act_t* __restrict__ out_hi,
                   act_t* __restrict__ out_lo, int Cin, int relu, int ph, int pw, float promo_scale, int* __restrict__ ovf,
                   unsigned long long* __restrict__ dbg) {
  using C = ResCfg<KC_, COUT>;
  constexpr int KC = C::kKC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t base = (smem0 + 1023u) & ~1023u;
  const uint32_t bres = base + C::kStages * C::kStageBytes;      // resident weights: tap t at bres + t * 2 * kBTile (hi, lo)
  const uint32_t bar_base = bres + C::kBBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t aux = bar_base + 8u * (2 * C::kStages);
  auto hh_full = [&](int b) { return aux + 8u * b; };
  auto hh_empty = [&](int b) { return aux + 8u * (3 + b); };
  // cross-term accumulators: i = 0 hi*lo, 1 lo*hi; each double-buffered by tile parity tp (the promotion warps read
  // them at the very end of a tile: with one buffer the cross-term threads idled ~1.0-1.4 k cycles per tile)
  auto x_full = [&](int i, int tp) { return aux + 8u * (6 + 2 * i + tp); };
  auto x_empty = [&](int i, int tp) { return aux + 8u * (10 + 2 * i + tp); };
  const uint32_t b_full = aux + 8u * 14;
  auto sched_full = [&](int s) { return aux + 8u * (15 + s); };
  auto sched_empty = [&](int s) { return aux + 8u * (15 + kSched + s); };
  const uint32_t tmem_slot = aux + 8u * (15 + 2 * kSched);
  const uint32_t ring = tmem_slot + 8u + ((tmem_slot + 8u) & 8u);  // 16-byte aligned
  static_assert(8 * (2 * C::kStages + 15 + 2 * kSched) + 24 + kSched * (int)sizeof(TileEntry) <= C::kBarBytes, "barrier area");
  TileEntry* ring_p = reinterpret_cast<TileEntry*>(smem_raw + (ring - smem0));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), C::kMmaWarps);  // one commit per issuing thread
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(hh_full(b), 1);
      mbar_init(hh_empty(b), C::kEpiWarps);
    }
    for (int i = 0; i < 2; ++i)
      for (int tp = 0; tp < 2; ++tp) {
        mbar_init(x_full(i, tp), 1);
        mbar_init(x_empty(i, tp), C::kEpiWarps);
      }
    mbar_init(b_full, 1);
    for (int s = 0; s < kSched; ++s) {
      mbar_init(sched_full(s), 1);
      mbar_init(sched_empty(s), C::kEpiWarps + C::kMmaWarps);
    }
    fence_barrier_init();
  }
  if (warp == C::kEpiWarps) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  constexpr int nkb = 9;  // Cin == KC
  (void)Cin;

  if (warp == C::kEpiWarps) {
    if (lane == 0) {
      // ---------------- weights once, then tile scheduler + TMA producer ----------------
      mbar_expect_tx(b_full, C::kBBytes);
      for (int t = 0; t < 9; ++t) {
        tma_load_2d(bres + t * 2 * C::kBTile, &tm_w_hi, t * KC, 0, b_full);
        tma_load_2d(bres + t * 2 * C::kBTile + C::kBTile, &tm_w_lo, t * KC, 0, b_full);
      }
      uint32_t stage = 0, par = 0;
      int cur_g = 0, last_map_g = -1;
      RaggedDesc gd = groups[0];
      int t = atomicAdd(counter, 1);
      const bool timing = dbg != nullptr;
      unsigned long long p_slot = 0, p_dec = 0, p_emp = 0;
      const long long p_begin = timing ? clock64() : 0;
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        long long p0 = timing ? clock64() : 0;
        mbar_wait(sched_empty(slot), ((ti / kSched) & 1) ^ 1);
        long long p1 = timing ? clock64() : 0;
        TileEntry e;
        e.g = -1;
        e.n = e.h0 = e.w0 = e.H = e.W = e.OH = e.OW = 0;
        e.out_off = 0;
        e.pad = 0;
        if (t < n_tiles) {
          while (cur_g + 1 < n_groups && __ldg(&groups[cur_g + 1].first) <= t) {
            ++cur_g;
            gd = groups[cur_g];
          }
          int l = t - gd.first;
          e.g = cur_g;
          e.w0 = (l % gd.tiles_w) * kTW;
          l /= gd.tiles_w;
          e.h0 = (l % gd.tiles_h) * kTH;
          e.n = l / gd.tiles_h;
          e.H = gd.H; e.W = gd.W; e.OH = gd.OH; e.OW = gd.OW;
          e.out_off = gd.out_off;
        }
        ring_p[slot] = e;
        mbar_arrive(sched_full(slot));
        if (timing) {
          p_slot += (unsigned long long)(p1 - p0);
          p_dec += (unsigned long long)(clock64() - p1);
        }
        if (e.g < 0) break;
        const int t_next = atomicAdd(counter, 1);
        const CUtensorMap* m_hi = maps + 2 * e.g;
        const CUtensorMap* m_lo = m_hi + 1;
        if (e.g != last_map_g) {
          fence_tensormap_acquire(m_hi);
          fence_tensormap_acquire(m_lo);
          last_map_g = e.g;
        }
        for (int kw = 0; kw < 3; ++kw) {
          long long p2 = timing ? clock64() : 0;
          mbar_wait(empty_bar(stage), par ^ 1);
          if (timing) p_emp += (unsigned long long)(clock64() - p2);
          const uint32_t st = base + stage * C::kStageBytes;
          mbar_expect_tx(full_bar(stage), C::kStageBytes);
          tma_load_4d(st, m_hi, 0, e.w0 + kw - 1, e.h0 - 1, e.n, full_bar(stage));
          tma_load_4d(st + C::kABox, m_lo, 0, e.w0 + kw - 1, e.h0 - 1, e.n, full_bar(stage));
          if (++stage == C::kStages) { stage = 0; par ^= 1; }
        }
        t = t_next;
      }
      if (timing) {
        atomicAdd(dbg + 10, p_slot);
        atomicAdd(dbg + 11, p_dec);
        atomicAdd(dbg + 12, p_emp);
        atomicAdd(dbg + 13, (unsigned long long)(clock64() - p_begin));
      }
    }
  } else if (warp > C::kEpiWarps) {
    if (lane == 0) {
      // ---------------- MMA issuers ----------------
      // A single thread issues one tcgen05.mma per ~53+ cycles whatever its size (profiles/r02a_ubench_mma.log), and
      // this layer's MMAs are only 32 cycles of tensor work each: the three products of the split go out from three
      // threads, each into its own accumulator (role 0: hi*hi with the promotion protocol, 1: hi*lo, 2: lo*hi), so
      // the order of accumulation inside every accumulator is fixed.
      const int role = warp - (C::kEpiWarps + 1);
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t stage = 0, par = 0, hb = 0, hpar = 0;
      mbar_wait(b_full, 0);
      const uint64_t db_base = make_desc<KC>(bres);
      const bool timing = dbg != nullptr;
      unsigned long long c_sched = 0, c_hhe = 0, c_full = 0, c_xe = 0, n_tiles_done = 0;
      const long long t_begin = timing ? clock64() : 0;
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        long long c0 = timing ? clock64() : 0;
        mbar_wait(sched_full(slot), (ti / kSched) & 1);
        if (timing) c_sched += (unsigned long long)(clock64() - c0);
        int g;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(g) : "r"(ring + slot * (uint32_t)sizeof(TileEntry)) : "memory");
        mbar_arrive(sched_empty(slot));
        if (g < 0) break;
        const int tp = (int)(ti & 1);
        // role 1: columns (3 + tp) * COUT, role 2: (5 + tp) * COUT
        const uint32_t d_x = tmem_base + (uint32_t)((1 + 2 * role + tp) * COUT);
        if (role != 0) {
          long long x0 = timing ? clock64() : 0;
          mbar_wait(x_empty(role - 1, tp), ((ti >> 1) & 1) ^ 1);
          tc_fence_after();
          if (timing) c_xe += (unsigned long long)(clock64() - x0);
        }
        ++n_tiles_done;
        int kbi = 0;
        for (int kw = 0; kw < 3; ++kw) {
          long long w0 = timing ? clock64() : 0;
          mbar_wait(full_bar(stage), par);
          tc_fence_after();
          if (timing) c_full += (unsigned long long)(clock64() - w0);
          const uint64_t d0 = make_desc<KC>(base + stage * C::kStageBytes);
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh, ++kbi) {
            const uint64_t da0 = d0 + (uint64_t)((kh * kTW * KC * 2) >> 4);
            const uint64_t db0 = db_base + (uint64_t)(((kh * 3 + kw) * 2 * C::kBTile) >> 4);
            if (role == 0) {
              const uint32_t d_hh = tmem_base + hb * COUT;
              if ((kbi & (C::kGroupKb - 1)) == 0) {
                long long g0 = timing ? clock64() : 0;
                mbar_wait(hh_empty(hb), hpar ^ 1);
                tc_fence_after();
                if (timing) c_hhe += (unsigned long long)(clock64() - g0);
              }
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                umma_bf16(d_hh, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, ((kbi & (C::kGroupKb - 1)) | k) ? 1u : 0u);
              if ((kbi & (C::kGroupKb - 1)) == C::kGroupKb - 1 || kbi == nkb - 1) {
                umma_commit(hh_full(hb));
                if (++hb == 3) { hb = 0; hpar ^= 1; }
              }
            } else {
              // role 1: A_hi x B_lo, role 2: A_lo x B_hi
              const uint64_t da = da0 + (role == 2 ? (uint64_t)(C::kABox >> 4) : 0ull);
              const uint64_t db = db0 + (role == 1 ? (uint64_t)(C::kBTile >> 4) : 0ull);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) umma_bf16(d_x, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kbi | k) ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == C::kStages) { stage = 0; par ^= 1; }
        }
        if (role != 0) umma_commit(x_full(role - 1, tp));
      }
      if (timing) {
        // per-role sums: [16 + 8 * role + ...]: tiles, ring, x_empty, hh_empty, operands, total
        unsigned long long* o = dbg + 16 + 8 * role;
        atomicAdd(o + 0, n_tiles_done);
        atomicAdd(o + 1, c_sched);
        atomicAdd(o + 2, c_xe);
        atomicAdd(o + 3, c_hhe);
        atomicAdd(o + 4, c_full);
        atomicAdd(o + 5, (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // ---------------- promotion + epilogue ----------------
    constexpr int CH = 32;
    const int wq = warp & 3, hsel = warp >> 2;
    const uint32_t lane_base = ((uint32_t)(wq * 32) << 16) + (uint32_t)(hsel * CH);
    constexpr int ngroups = (nkb + C::kGroupKb - 1) / C::kGroupKb;
    uint32_t hb = 0, hpar = 0;
    for (uint32_t ti = 0;; ++ti) {
      const uint32_t slot = ti & (kSched - 1);
      mbar_wait(sched_full(slot), (ti / kSched) & 1);
      const TileEntry e = ring_p[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(sched_empty(slot));
      if (e.g < 0) break;
      float acc[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = 0.f;
      for (int g = 0; g < ngroups; ++g) {
        mbar_wait(hh_full(hb), hpar);
        tc_fence_after();
        {
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_base + hb * COUT, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(__uint_as_float(r[j]), promo_scale, acc[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(hh_empty(hb));
        if (++hb == 3) { hb = 0; hpar ^= 1; }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {  // + hi*lo, then + lo*hi
        const int tp = (int)(ti & 1);
        mbar_wait(x_full(i, tp), (ti >> 1) & 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)((3 + 2 * i + tp) * COUT), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(x_empty(i, tp));
      }
      epilogue_store<COUT>(acc, e, wq, hsel, lane, relu, ph, pw, true, bias, out_hi, out_lo, ovf);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == C::kEpiWarps) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// Weights-stationary, transposed form for the 32 -> 64 layer with a fused 2x2 pool (`conv3x3_ws_kernel`).
// The resident-weights kernel above is bound by shared-memory operand delivery: an MMA of N = 64 reads 4 KB of
// pixels + 2 KB of weights (64-byte-swizzled rows, ~60 B/clk) for 32 cycles of tensor work.  Here the roles are swapped,
// D^T[c_out][pixel] = W x X^T:
//   * A = [W_hi (64 rows); W_lo (64 rows)] of all nine taps lives in TENSOR MEMORY (lane = row, 144 columns), loaded
//     once per CTA -- no shared-memory reads for it at all;
//   * B = the activation box in shared memory (128 pixels x 16 channels = 4 KB per MMA of 64 cycles);
//   * TWO MMAs per 16 K-elements instead of three:  [W_hi; W_lo] x X_hi  ->  D1 = [hi*hi ; lo*hi]  (lanes 0-63 / 64-127)
//                                                    [W_hi;  .  ] x X_lo  ->  D2 = [hi*lo ;  .   ]
//     one issuing thread per accumulator.
// Tensor memory: A 160 + D1 128 + D2 128 columns.  There is no room for the promotion double buffer (that would be 528
// columns), and this layer does not need it: K = 288 is 18 MMAs per accumulator, i.e. a truncation bias of at most
// 9 ulp of the running sum (the promotion exists for the K = 576 / 1152 layers); the expected bias is scaled out like
// in the other kernels.  The accumulators are single-buffered: the MMAs of the next tile wait until the 16 epilogue
// warps have pulled the tile into registers (a bubble of a few hundred cycles per tile), the epilogue arithmetic and
// stores then overlap the next tile's MMAs.  A first version with N = 64 and the promotion protocol was parity-green but
// slower than the resident-weights kernel (per-tile fixed costs; profiles/r02ad_ws_kernel.md).
// Output lanes are channels and registers are pixels: warp (lane quadrant q, pixel chunk pc) holds channel
// 32*(q%2) + lane for tile rows 2pc, 2pc+1 -- one 2x2 pooling row pair, so the pool is a max over registers.  The lo*hi
// sums (quadrants 2, 3) reach the warps that own the channel (quadrants 0, 1) through shared memory, double-buffered
// by tile parity, one named barrier per warp pair.
// ------------------------------------------------------------------------------------------
struct WsCfg {
  static constexpr int kKC = 32, kCout = 64, kN = kTH * kTW;                   // 128 pixels
  static constexpr int kABox = (kTH + 2) * kTW * kKC * 2;                      // one plane of the 10 x 16 pixel box: 10 KB
  static constexpr int kStageBytes = 2 * kABox;
  static constexpr int kStages = 4;
  static constexpr int kXchBytes = 2 * 2 * 4 * 32 * 32 * 4;                    // [parity][channel half][pixel chunk][px][lane] f32
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kXchBytes + 1024 + kBarBytes;
  static constexpr int kColA = 0, kColD1 = 192, kColD2 = 320, kTmemCols = 512;  // A: 144 (5 x 32 stored) | D1 128 | D2 128
  static constexpr int kEpiWarps = 16;
  static constexpr int kThreads = 32 * (kEpiWarps + 3);                         // + TMA / scheduler, D1 issuer, D2 issuer
  static_assert(kABox % 512 == 0, "box planes must keep the 64-byte swizzle phase");
};

__global__ void __launch_bounds__(WsCfg::kThreads, 1)
conv3x3_ws_kernel(const CUtensorMap* __restrict__ maps, const RaggedDesc* __restrict__ groups, int n_groups, int n_tiles,
                  int* __restrict__ counter, const act_t* __restrict__ w_hi, const act_t* __restrict__ w_lo,
                  const float* __restrict__ bias, act_t* __restrict__ out_hi, act_t* __restrict__ out_lo, int relu,
                  float debias_scale, int* __restrict__ ovf) {
  using C = WsCfg;
  constexpr int KC = C::kKC, COUT = C::kCout;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t base = (smem0 + 1023u) & ~1023u;
  const uint32_t xch = base + C::kStages * C::kStageBytes;
  const uint32_t bar_base = xch + C::kXchBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t aux = bar_base + 8u * (2 * C::kStages);
  const uint32_t d1_full = aux, d1_empty = aux + 8u, d2_full = aux + 16u, d2_empty = aux + 24u;
  auto sched_full = [&](int s) { return aux + 8u * (4 + s); };
  auto sched_empty = [&](int s) { return aux + 8u * (4 + kSched + s); };
  const uint32_t tmem_slot = aux + 8u * (4 + 2 * kSched);
  const uint32_t ring = tmem_slot + 16u;
  static_assert(8 * (2 * C::kStages + 4 + 2 * kSched) + 16 + kSched * (int)sizeof(TileEntry) <= C::kBarBytes, "barrier area");
  TileEntry* ring_p = reinterpret_cast<TileEntry*>(smem_raw + (ring - smem0));
  float* xch_p = reinterpret_cast<float*>(smem_raw + (xch - smem0));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 2);  // one commit per issuing thread
    }
    mbar_init(d1_full, 1);
    mbar_init(d1_empty, C::kEpiWarps);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, C::kEpiWarps / 2);  // only the channel-owning warps read D2
    for (int s = 0; s < kSched; ++s) {
      mbar_init(sched_full(s), 1);
      mbar_init(sched_empty(s), C::kEpiWarps + 2);
    }
    fence_barrier_init();
  }
  if (warp == C::kEpiWarps) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    // ---- one-time: the weights become the MMA A operand in tensor memory.  Lane = row (quadrants 0, 1: W_hi rows
    // 0..63; quadrants 2, 3: W_lo rows 0..63), columns = packed fp16 pairs along k = tap * 32 + channel (144 words).
    const act_t* src = (warp < 2 ? w_hi : w_lo) + (size_t)((warp & 1) * 32 + lane) * (9 * KC);
    const uint4* src4 = reinterpret_cast<const uint4*>(src);  // a row is 576 B = 36 x 16 B
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)C::kColA;
#pragma unroll 1
    for (int c = 0; c < 5; ++c) {  // 5 x 32 columns; the last 16 of the 160 are padding
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int q = c * 8 + j;
        const uint4 v = q < 36 ? __ldg(src4 + q) : make_uint4(0, 0, 0, 0);
        r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
      }
      tmem_st32(lane_base + (uint32_t)(c * 32), r);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // c_format F32 (bit 4), a/b format F16, N >> 3 at 17, M >> 4 at 24
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(C::kN >> 3) << 17) | ((128u >> 4) << 24);

  if (warp == C::kEpiWarps) {
    if (lane == 0) {
      // ---------------- tile scheduler + TMA producer ----------------
      uint32_t stage = 0, par = 0;
      int cur_g = 0, last_map_g = -1;
      RaggedDesc gd = groups[0];
      int t = atomicAdd(counter, 1);
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        mbar_wait(sched_empty(slot), ((ti / kSched) & 1) ^ 1);
        TileEntry e;
        e.g = -1;
        e.n = e.h0 = e.w0 = e.H = e.W = e.OH = e.OW = 0;
        e.out_off = 0;
        e.pad = 0;
        if (t < n_tiles) {
          while (cur_g + 1 < n_groups && __ldg(&groups[cur_g + 1].first) <= t) {
            ++cur_g;
            gd = groups[cur_g];
          }
          int l = t - gd.first;
          e.g = cur_g;
          e.w0 = (l % gd.tiles_w) * kTW;
          l /= gd.tiles_w;
          e.h0 = (l % gd.tiles_h) * kTH;
          e.n = l / gd.tiles_h;
          e.H = gd.H; e.W = gd.W; e.OH = gd.OH; e.OW = gd.OW;
          e.out_off = gd.out_off;
        }
        ring_p[slot] = e;
        mbar_arrive(sched_full(slot));
        if (e.g < 0) break;
        const int t_next = atomicAdd(counter, 1);
        const CUtensorMap* m_hi = maps + 2 * e.g;
        const CUtensorMap* m_lo = m_hi + 1;
        if (e.g != last_map_g) {
          fence_tensormap_acquire(m_hi);
          fence_tensormap_acquire(m_lo);
          last_map_g = e.g;
        }
        for (int kw = 0; kw < 3; ++kw) {
          mbar_wait(empty_bar(stage), par ^ 1);
          const uint32_t st = base + stage * C::kStageBytes;
          mbar_expect_tx(full_bar(stage), C::kStageBytes);
          tma_load_4d(st, m_hi, 0, e.w0 + kw - 1, e.h0 - 1, e.n, full_bar(stage));
          tma_load_4d(st + C::kABox, m_lo, 0, e.w0 + kw - 1, e.h0 - 1, e.n, full_bar(stage));
          if (++stage == C::kStages) { stage = 0; par ^= 1; }
        }
        t = t_next;
      }
    }
  } else if (warp > C::kEpiWarps) {
    if (lane == 0) {
      // ---------------- MMA issuers: role 0 -> D1 = [W_hi; W_lo] x X_hi, role 1 -> D2 = [W_hi; .] x X_lo ----------------
      const int role = warp - (C::kEpiWarps + 1);
      const uint32_t d = tmem_base + (uint32_t)(role == 0 ? C::kColD1 : C::kColD2);
      const uint32_t d_empty = role == 0 ? d1_empty : d2_empty, d_full = role == 0 ? d1_full : d2_full;
      uint32_t stage = 0, par = 0;
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        mbar_wait(sched_full(slot), (ti / kSched) & 1);
        int g;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(g) : "r"(ring + slot * (uint32_t)sizeof(TileEntry)) : "memory");
        mbar_arrive(sched_empty(slot));
        if (g < 0) break;
        mbar_wait(d_empty, (ti & 1) ^ 1);  // the epilogue warps have the previous tile in registers
        tc_fence_after();
        for (int kw = 0; kw < 3; ++kw) {
          mbar_wait(full_bar(stage), par);
          tc_fence_after();
          // B operand: the pixel rows kh .. kh+7 of the box (16 pixels x 64 B per row), hi plane for D1, lo plane for D2
          const uint64_t d0 = make_desc<KC>(base + stage * C::kStageBytes + (role == 1 ? C::kABox : 0));
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint64_t db0 = d0 + (uint64_t)((kh * kTW * KC * 2) >> 4);
            const uint32_t ta0 = tmem_base + (uint32_t)(C::kColA + (kh * 3 + kw) * 16);  // 2 k-steps x 8 columns per tap
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_bf16_ts(d, ta0 + (uint32_t)(8 * k), db0 + (uint64_t)(2 * k), idesc, (kw | kh | k) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == C::kStages) { stage = 0; par ^= 1; }
        }
        umma_commit(d_full);
      }
    }
  } else {
    // ---------------- epilogue: warp = (lane quadrant q, pixel chunk pc) ----------------
    const int q = warp & 3, pc = warp >> 2;
    const bool owner = q < 2;           // quadrants 0, 1: hi*hi lanes = the warps that own channel 32*q + lane
    const int ch = (q & 1) * 32 + lane;  // output channel of this thread
    const uint32_t lane_base = ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * pc);
    const float bias_c = bias[ch];
    const float scale = owner ? debias_scale : 1.0f;  // the de-bias scale belongs to the hi*hi sums only
    const int bar_id = 1 + (q & 1) * 4 + pc;          // one named barrier per (channel half, pixel chunk) warp pair
    for (uint32_t ti = 0;; ++ti) {
      const uint32_t slot = ti & (kSched - 1);
      mbar_wait(sched_full(slot), (ti / kSched) & 1);
      const TileEntry e = ring_p[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(sched_empty(slot));
      if (e.g < 0) break;
      const uint32_t tp = ti & 1;
      float acc[32];
      mbar_wait(d1_full, tp);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)C::kColD1, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]) * scale;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d1_empty);
      // exchange slot of this (tile parity, channel half, pixel chunk): [pixel j][lane]
      float* xs = xch_p + (size_t)(((tp * 2 + (q & 1)) * 4 + pc) * 32 * 32);
      if (!owner) {
#pragma unroll
        for (int j = 0; j < 32; ++j) xs[j * 32 + lane] = acc[j];
        named_bar_sync(bar_id, 64);
        continue;
      }
      mbar_wait(d2_full, tp);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)C::kColD2, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);  // + hi*lo
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty);
      named_bar_sync(bar_id, 64);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += xs[j * 32 + lane];        // + lo*hi
      // bias, ReLU, 2x2 max-pool over the thread's two tile rows (registers j and 16 + j), fp16 split, NHWC store
      const int oh = (e.h0 + 2 * pc) >> 1;
      if (oh < e.OH) {
        const size_t row_pix = (size_t)e.out_off + ((size_t)e.n * e.OH + oh) * e.OW;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int ow = (e.w0 >> 1) + jj;
          float v = fmaxf(fmaxf(acc[2 * jj], acc[2 * jj + 1]), fmaxf(acc[16 + 2 * jj], acc[16 + 2 * jj + 1])) + bias_c;
          if (relu) v = fmaxf(v, 0.f);
          uint16_t hi, lo;
          split1(v, hi, lo, ovf);
          if (ow < e.OW) {
            out_hi[(row_pix + ow) * COUT + ch] = __ushort_as_half(hi);
            out_lo[(row_pix + ow) * COUT + ch] = __ushort_as_half(lo);
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == C::kEpiWarps) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// CTA-pair form of the kernel for the COUT = 128, KC = 64 layers: two CTAs of a cluster (the two SMs of a
// TPC) run ONE tcgen05.mma.cta_group::2 of M = 256 per step -- each CTA owns its own 128-pixel tile (A
// operand, its own TMEM accumulators) and HALF of the weight tile (64 of the 128 output channels); the
// tensor cores of both SMs read both halves.  Per CTA and k-block the operand traffic drops from 64 KB to
// 48 KB of TMA writes and from 8 KB to 6 KB of shared-memory reads per MMA, which is what bounds the
// single-CTA kernel (it sits at the L2 -> SM delivery rate with the tensor pipe ~50 % active).
// Protocol (rank 0 = leader):
//   * tiles are drawn in PAIRS from the global counter by the leader's scheduler thread, which writes the
//     peer's ring entry through DSMEM; a missing second tile is a dummy (TMA zero fill, no stores);
//   * both CTAs' TMA loads complete on the LEADER's full barrier (expect_tx = 2 stages' bytes);
//   * only the leader issues MMAs; tcgen05.commit ... multicast::cluster releases the stage in both CTAs
//     and publishes the accumulators to both CTAs' promotion warps;
//   * the promotion warps of both CTAs return accumulator buffers on the leader's barriers (remote arrive).
// The arithmetic per output pixel is that of the single-CTA kernel (same MMA order into the same two
// accumulators), so the results are bit-identical.
// ------------------------------------------------------------------------------------------
template <int COUT>
struct PairCfg {
  static constexpr int kKC = 64;
  static constexpr int kABytes = 128 * kKC * 2;            // one plane of the A tile
  static constexpr int kBHalf = (COUT / 2) * kKC * 2;      // one plane of this CTA's half of the weights
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBHalf;
  static constexpr int kStages = 4;
  static constexpr int kGroupKb = 2;
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + kBarBytes;
  static constexpr int kTmemCols = 4 * COUT;
  static constexpr int kEpiWarps = 4 * (COUT / 32);
  static constexpr int kThreads = 32 * (kEpiWarps + 2);
};

__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// TMA loads whose completion is signalled on `bar_cluster` (a shared::cluster address: the leader's barrier)
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                             uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// bounded cluster-scope wait (young kernel: a protocol bug must trap, not hang the GPU)
__device__ __forceinline__ void mbar_wait_cluster_trap(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 26); ++i) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

template <int COUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((PairCfg<COUT>::kThreads), 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                    const CUtensorMap* __restrict__ maps, const RaggedDesc* __restrict__ groups, int n_groups, int n_tiles,
                    int* __restrict__ counter, const float* __restrict__ bias, act_t* __restrict__ out_hi,
                    act_t* __restrict__ out_lo, int Cin, int relu, int ph, int pw, float promo_scale, int* __restrict__ ovf,
                    unsigned long long* __restrict__ dbg) {
  using C = PairCfg<COUT>;
  constexpr int KC = C::kKC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t base = (smem0 + 1023u) & ~1023u;
  const uint32_t bar_base = base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };       // the leader's are used
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t aux = bar_base + 8u * (2 * C::kStages);
  auto hh_full = [&](int b) { return aux + 8u * b; };
  auto hh_empty = [&](int b) { return aux + 8u * (2 + b); };      // the leader's are used
  auto x_full = [&](int b) { return aux + 8u * (4 + b); };
  auto x_empty = [&](int b) { return aux + 8u * (6 + b); };       // the leader's are used
  auto sched_full = [&](int s) { return aux + 8u * (8 + s); };
  auto sched_empty = [&](int s) { return aux + 8u * (8 + kSched + s); };  // the leader's are used
  const uint32_t tmem_slot = aux + 8u * (8 + 2 * kSched);
  const uint32_t ring = tmem_slot + 16u;
  static_assert(8 * (2 * 4 + 8 + 2 * kSched) + 16 + kSched * (int)sizeof(TileEntry) <= C::kBarBytes, "barrier area");
  TileEntry* ring_p = reinterpret_cast<TileEntry*>(smem_raw + (ring - smem0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);   // leader: its producer arms; both CTAs' loads complete_tx on it
      mbar_init(empty_bar(s), 1);  // one multicast commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(hh_full(b), 1);
      mbar_init(hh_empty(b), 2 * C::kEpiWarps);  // leader: the promotion warps of both CTAs
      mbar_init(x_full(b), 1);
      mbar_init(x_empty(b), 2 * C::kEpiWarps);
    }
    for (int s = 0; s < kSched; ++s) {
      mbar_init(sched_full(s), 1);
      mbar_init(sched_empty(s), 2 * C::kEpiWarps + 2);  // leader: MMA thread + peer's TMA thread + all promotion warps
    }
    fence_barrier_init();
  }
  if (warp == C::kEpiWarps) tmem_alloc2(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  cluster_sync_all();  // both CTAs' barriers and tensor memory exist before anything crosses the pair

  const int chunks = Cin / KC;
  const int nkb = 9 * chunks;
  const int n_pairs = (n_tiles + 1) / 2;

  if (warp == C::kEpiWarps) {
    if (lane == 0) {
      // ---------------- tile scheduler (leader) + TMA producer (both CTAs) ----------------
      uint32_t stage = 0, par = 0;
      int cur_g = 0, last_map_g = -1;
      RaggedDesc gd = groups[0];
      int pr = leader ? atomicAdd(counter, 1) : 0;
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        TileEntry e;
        if (leader) {
          mbar_wait_cluster_trap(sched_empty(slot), ((ti / kSched) & 1) ^ 1);
          TileEntry ent[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            TileEntry& x = ent[k];
            x.g = -1;
            x.n = x.h0 = x.w0 = x.H = x.W = x.OH = x.OW = 0;
            x.out_off = 0;
            x.pad = 0;  // pad = 1: real tile, 0: dummy (the pair's second tile does not exist)
            if (pr < n_pairs) {
              const int t = min(2 * pr + k, n_tiles - 1);
              while (cur_g + 1 < n_groups && __ldg(&groups[cur_g + 1].first) <= t) {
                ++cur_g;
                gd = groups[cur_g];
              }
              int l = t - gd.first;
              x.g = cur_g;
              x.w0 = (l % gd.tiles_w) * kTW;
              l /= gd.tiles_w;
              x.h0 = (l % gd.tiles_h) * kTH;
              x.n = l / gd.tiles_h;
              x.H = gd.H; x.W = gd.W; x.OH = gd.OH; x.OW = gd.OW;
              x.out_off = gd.out_off;
              x.pad = (2 * pr + k < n_tiles) ? 1 : 0;
              if (!x.pad) x.n = gd.N;  // outside the tensor: TMA fills zeros
            }
          }
          ring_p[slot] = ent[0];
          // the peer's entry through DSMEM (three 16-byte stores), then release-arrive on both rings
          const uint32_t peer_entry = mapa(ring + slot * (uint32_t)sizeof(TileEntry), 1);
          const uint4* src = reinterpret_cast<const uint4*>(&ent[1]);
#pragma unroll
          for (int q = 0; q < 3; ++q) st_cluster_v4(peer_entry + 16u * q, src[q]);
          mbar_arrive_remote_release(mapa(sched_full(slot), 1));
          mbar_arrive(sched_full(slot));
          e = ent[0];
        } else {
          mbar_wait_cluster_trap(sched_full(slot), (ti / kSched) & 1);
          e = ring_p[slot];
          mbar_arrive_remote_release(mapa(sched_empty(slot), 0));
        }
        if (e.g < 0) break;
        int pr_next = 0;
        if (leader) pr_next = atomicAdd(counter, 1);  // in flight while this tile's loads are issued
        const CUtensorMap* m_hi = maps + 2 * e.g;
        const CUtensorMap* m_lo = m_hi + 1;
        if (e.g != last_map_g) {
          fence_tensormap_acquire(m_hi);
          fence_tensormap_acquire(m_lo);
          last_map_g = e.g;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_trap(empty_bar(stage), par ^ 1);
          const uint32_t st = base + stage * C::kStageBytes;
          const uint32_t lead_full = mapa(full_bar(stage), 0);
          if (leader) mbar_expect_tx(full_bar(stage), 2 * C::kStageBytes);
          const int tap = kb / chunks, c0 = (kb - tap * chunks) * KC;
          const int kh = tap / 3, kw = tap - kh * 3;
          tma2_load_4d(st, m_hi, c0, e.w0 + kw - 1, e.h0 + kh - 1, e.n, lead_full);
          tma2_load_4d(st + C::kABytes, m_lo, c0, e.w0 + kw - 1, e.h0 + kh - 1, e.n, lead_full);
          tma2_load_2d(st + 2 * C::kABytes, &tm_w_hi, tap * Cin + c0, (int)crank * (COUT / 2), lead_full);
          tma2_load_2d(st + 2 * C::kABytes + C::kBHalf, &tm_w_lo, tap * Cin + c0, (int)crank * (COUT / 2), lead_full);
          if (++stage == C::kStages) { stage = 0; par ^= 1; }
        }
        pr = pr_next;
      }
    }
  } else if (warp == C::kEpiWarps + 1) {
    if (lane == 0 && leader) {
      // ---------------- MMA issuer (leader only) ----------------
      // c_format F32 (bit 4), a/b format F16, N >> 3 at 17, M >> 4 at 24 with M = 256 over the pair
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((256u >> 4) << 24);
      uint32_t stage = 0, par = 0, gc = 0;
      unsigned long long c_sched = 0, c_xe = 0, c_hhe = 0, c_full = 0, c_iss = 0, n_tiles_done = 0;
      const bool timing = dbg != nullptr;
      const long long t_begin = timing ? clock64() : 0;
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        long long c0 = timing ? clock64() : 0;
        mbar_wait_trap(sched_full(slot), (ti / kSched) & 1);
        int g;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(g) : "r"(ring + slot * (uint32_t)sizeof(TileEntry)) : "memory");
        mbar_arrive(sched_empty(slot));
        if (g < 0) break;
        const uint32_t tp = ti & 1;
        const uint32_t d_x = tmem_base + 2 * COUT + tp * COUT;
        long long c1 = timing ? clock64() : 0;
        mbar_wait_cluster_trap(x_empty(tp), ((ti >> 1) & 1) ^ 1);
        long long c2 = timing ? clock64() : 0;
        c_sched += (unsigned long long)(c1 - c0);
        c_xe += (unsigned long long)(c2 - c1);
        for (int kb = 0; kb < nkb; ++gc) {
          const uint32_t b = gc & 1;
          const uint32_t d_hh = tmem_base + b * COUT;
          const int nk = min(C::kGroupKb, nkb - kb);
          long long g0 = timing ? clock64() : 0;
          mbar_wait_cluster_trap(hh_empty(b), ((gc >> 1) & 1) ^ 1);
          tc_fence_after();
          long long g1 = timing ? clock64() : 0;
          c_hhe += (unsigned long long)(g1 - g0);
          for (int j = 0; j < nk; ++j) {
            long long w0 = timing ? clock64() : 0;
            mbar_wait_cluster_trap(full_bar(stage), par);
            tc_fence_after();
            if (timing) c_full += (unsigned long long)(clock64() - w0);
            const uint64_t d0 = make_desc<KC>(base + stage * C::kStageBytes);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              const uint64_t da_hi = d0 + (uint64_t)(2 * k), da_lo = da_hi + (uint64_t)(C::kABytes >> 4);
              const uint64_t db_hi = da_hi + (uint64_t)((2 * C::kABytes) >> 4), db_lo = db_hi + (uint64_t)(C::kBHalf >> 4);
              umma2_f16(d_hh, da_hi, db_hi, idesc, (j | k) ? 1u : 0u);
              umma2_f16(d_x, da_hi, db_lo, idesc, (kb | j | k) ? 1u : 0u);
              umma2_f16(d_x, da_lo, db_hi, idesc, 1u);
            }
            umma2_commit_mc(empty_bar(stage), 3);
            if (++stage == C::kStages) { stage = 0; par ^= 1; }
          }
          umma2_commit_mc(hh_full(b), 3);
          if (timing) c_iss += (unsigned long long)(clock64() - g1);
          kb += nk;
        }
        umma2_commit_mc(x_full(tp), 3);
        ++n_tiles_done;
      }
      if (timing) {
        atomicAdd(dbg + 0, 2 * n_tiles_done);
        atomicAdd(dbg + 1, n_tiles_done * (unsigned long long)nkb);
        atomicAdd(dbg + 2, c_sched);
        atomicAdd(dbg + 3, c_xe);
        atomicAdd(dbg + 4, c_hhe);
        atomicAdd(dbg + 5, c_full);
        atomicAdd(dbg + 6, c_iss - c_full);
        atomicAdd(dbg + 7, 0ull);
        atomicAdd(dbg + 8, (unsigned long long)(clock64() - t_begin));
        atomicAdd(dbg + 9, 1ull);
      }
    }
  } else {
    // ---------------- promotion + epilogue (both CTAs, own tensor memory) ----------------
    constexpr int CH = 32;
    const int wq = warp & 3, hsel = warp >> 2;
    const uint32_t lane_base = ((uint32_t)(wq * 32) << 16) + (uint32_t)(hsel * CH);
    const int ngroups = (nkb + C::kGroupKb - 1) / C::kGroupKb;
    uint32_t gc = 0;
    for (uint32_t ti = 0;; ++ti) {
      const uint32_t slot = ti & (kSched - 1);
      mbar_wait_cluster_trap(sched_full(slot), (ti / kSched) & 1);
      const TileEntry e = ring_p[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_release(mapa(sched_empty(slot), 0));
      if (e.g < 0) break;
      const bool real_tile = e.pad != 0;
      float acc[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = 0.f;
      for (int g = 0; g < ngroups; ++g, ++gc) {
        const uint32_t b = gc & 1;
        mbar_wait_trap(hh_full(b), (gc >> 1) & 1);
        tc_fence_after();
        {
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_base + b * COUT, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(__uint_as_float(r[j]), promo_scale, acc[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(mapa(hh_empty(b), 0));  // no data handed over: the TMEM reads are done
      }
      const uint32_t tp = ti & 1;
      mbar_wait_trap(x_full(tp), (ti >> 1) & 1);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + 2 * COUT + tp * COUT, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_relaxed(mapa(x_empty(tp), 0));

      epilogue_store<COUT>(acc, e, wq, hsel, lane, relu, ph, pw, real_tile, bias, out_hi, out_lo, ovf);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's operands / signal its barriers
  if (warp == C::kEpiWarps) tmem_dealloc2(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// CTA pair + halo reuse.  Same pair protocol as conv3x3_pair_kernel; the operand pipeline is re-cut so that the
// 3x3 window re-reads almost nothing from L2: per (64-channel chunk, horizontal tap kw) ONE box of 10 rows x 16
// pixels of the activation is loaded (rows h0-1 .. h0+8) and the three vertical taps read it through descriptors
// advanced by whole 16-pixel rows (2 KB = two swizzle atoms, so the 128-byte swizzle phase is unchanged), plus
// this CTA's half of the three taps' weight tiles.  Per k-block and CTA: 30 KB of TMA traffic instead of 64 KB
// (single CTA) / 48 KB (pair) -- below the L2 -> SM delivery rate that bounds the other two forms.
// k-blocks run in (chunk, kw, kh) order, promotion groups are still 128 K-elements (2 or 4 k-blocks).
// (KC = 32 works too -- 64-byte rows, SWIZZLE_64B, a tile row is 1 KB = two atoms -- but is not instantiated.)
// ------------------------------------------------------------------------------------------
template <int KC_, int COUT>
struct HaloCfg {
  static constexpr int kKC = KC_;
  static constexpr int kABox = (kTH + 2) * kTW * kKC * 2;   // one plane of the 10 x 16 pixel box: 20 KB
  static constexpr int kBHalf = (COUT / 2) * kKC * 2;        // one plane of this CTA's half of one tap's weights: 8 KB
  static constexpr int kStageBytes = 2 * kABox + 3 * 2 * kBHalf;  // 88 KB for 64 x 128, 32 KB for 32 x 64
  static constexpr int kStagesRaw = (180 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 4 ? 4 : kStagesRaw;
  static constexpr int kGroupKb = 128 / kKC;     // k-blocks per promotion group (128 K-elements)
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + kBarBytes;
  static constexpr int kTmemCols = 4 * COUT;
  static constexpr int kEpiWarps = 4 * (COUT / 32);
  static constexpr int kMmaWarps = 2;  // one issuing thread for hi*hi, one for the cross terms
  static constexpr int kThreads = 32 * (kEpiWarps + 1 + kMmaWarps);
  static_assert(kABox % 1024 == 0 && kBHalf % 1024 == 0, "operand tiles must keep the swizzle phase");
};

template <int KC_, int COUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((HaloCfg<KC_, COUT>::kThreads), 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                    const CUtensorMap* __restrict__ maps, const RaggedDesc* __restrict__ groups, int n_groups, int n_tiles,
                    int* __restrict__ counter, const float* __restrict__ bias, act_t* __restrict__ out_hi,
                    act_t* __restrict__ out_lo, int Cin, int relu, int ph, int pw, float promo_scale, int* __restrict__ ovf,
                    unsigned long long* __restrict__ dbg) {
  using C = HaloCfg<KC_, COUT>;
  constexpr int KC = C::kKC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t base = (smem0 + 1023u) & ~1023u;
  const uint32_t bar_base = base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };       // the leader's are used
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t aux = bar_base + 8u * (2 * C::kStages);
  // tensor memory: three hi*hi accumulators (a promotion group each) + ONE cross-term accumulator.  The third
  // hi*hi buffer is what lets the MMA thread run 6 k-blocks ahead while the promotion warps are in the epilogue of
  // the previous tile; the cross-term accumulator needs no double buffer because it is read out at the very start
  // of the epilogue, and the next tile's first hi*hi MMAs are issued before the MMA thread waits for that.
  auto hh_full = [&](int b) { return aux + 8u * b; };
  auto hh_empty = [&](int b) { return aux + 8u * (3 + b); };      // the leader's are used
  const uint32_t x_full = aux + 8u * 6;
  const uint32_t x_empty = aux + 8u * 7;                          // the leader's is used
  auto sched_full = [&](int s) { return aux + 8u * (8 + s); };
  auto sched_empty = [&](int s) { return aux + 8u * (8 + kSched + s); };  // the leader's are used
  const uint32_t tmem_slot = aux + 8u * (8 + 2 * kSched);
  const uint32_t ring = tmem_slot + 16u;
  static_assert(8 * (2 * 4 + 8 + 2 * kSched) + 16 + kSched * (int)sizeof(TileEntry) <= C::kBarBytes, "barrier area");
  TileEntry* ring_p = reinterpret_cast<TileEntry*>(smem_raw + (ring - smem0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);   // leader: its producer arms; both CTAs' loads complete_tx on it
      mbar_init(empty_bar(s), C::kMmaWarps);  // one multicast commit per issuing thread
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(hh_full(b), 1);
      mbar_init(hh_empty(b), 2 * C::kEpiWarps);  // leader: the promotion warps of both CTAs
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 2 * C::kEpiWarps);
    for (int s = 0; s < kSched; ++s) {
      mbar_init(sched_full(s), 1);
      mbar_init(sched_empty(s), 2 * C::kEpiWarps + 1 + C::kMmaWarps);  // leader: MMA threads + peer's TMA thread + all promotion warps
    }
    fence_barrier_init();
  }
  if (warp == C::kEpiWarps) tmem_alloc2(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  cluster_sync_all();  // both CTAs' barriers and tensor memory exist before anything crosses the pair

  const int chunks = Cin / KC;
  const int nkb = 9 * chunks;
  const int n_pairs = (n_tiles + 1) / 2;

  if (warp == C::kEpiWarps) {
    if (lane == 0) {
      // ---------------- tile scheduler (leader) + TMA producer (both CTAs) ----------------
      uint32_t stage = 0, par = 0;
      int cur_g = 0, last_map_g = -1;
      RaggedDesc gd = groups[0];
      int pr = leader ? atomicAdd(counter, 1) : 0;
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        TileEntry e;
        if (leader) {
          mbar_wait_cluster_trap(sched_empty(slot), ((ti / kSched) & 1) ^ 1);
          TileEntry ent[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            TileEntry& x = ent[k];
            x.g = -1;
            x.n = x.h0 = x.w0 = x.H = x.W = x.OH = x.OW = 0;
            x.out_off = 0;
            x.pad = 0;  // pad = 1: real tile, 0: dummy (the pair's second tile does not exist)
            if (pr < n_pairs) {
              const int t = min(2 * pr + k, n_tiles - 1);
              while (cur_g + 1 < n_groups && __ldg(&groups[cur_g + 1].first) <= t) {
                ++cur_g;
                gd = groups[cur_g];
              }
              int l = t - gd.first;
              x.g = cur_g;
              x.w0 = (l % gd.tiles_w) * kTW;
              l /= gd.tiles_w;
              x.h0 = (l % gd.tiles_h) * kTH;
              x.n = l / gd.tiles_h;
              x.H = gd.H; x.W = gd.W; x.OH = gd.OH; x.OW = gd.OW;
              x.out_off = gd.out_off;
              x.pad = (2 * pr + k < n_tiles) ? 1 : 0;
              if (!x.pad) x.n = gd.N;  // outside the tensor: TMA fills zeros
            }
          }
          ring_p[slot] = ent[0];
          // the peer's entry through DSMEM (three 16-byte stores), then release-arrive on both rings
          const uint32_t peer_entry = mapa(ring + slot * (uint32_t)sizeof(TileEntry), 1);
          const uint4* src = reinterpret_cast<const uint4*>(&ent[1]);
#pragma unroll
          for (int q = 0; q < 3; ++q) st_cluster_v4(peer_entry + 16u * q, src[q]);
          mbar_arrive_remote_release(mapa(sched_full(slot), 1));
          mbar_arrive(sched_full(slot));
          e = ent[0];
        } else {
          mbar_wait_cluster_trap(sched_full(slot), (ti / kSched) & 1);
          e = ring_p[slot];
          mbar_arrive_remote_release(mapa(sched_empty(slot), 0));
        }
        if (e.g < 0) break;
        int pr_next = 0;
        if (leader) pr_next = atomicAdd(counter, 1);  // in flight while this tile's loads are issued
        const CUtensorMap* m_hi = maps + 2 * e.g;
        const CUtensorMap* m_lo = m_hi + 1;
        if (e.g != last_map_g) {
          fence_tensormap_acquire(m_hi);
          fence_tensormap_acquire(m_lo);
          last_map_g = e.g;
        }
        // one stage per (channel chunk, horizontal tap): the 10-row box of the activation serves the three
        // vertical taps (descriptor advanced by whole rows), plus this CTA's half of their three weight tiles
        for (int c = 0; c < chunks; ++c) {
          for (int kw = 0; kw < 3; ++kw) {
            mbar_wait_trap(empty_bar(stage), par ^ 1);
            const uint32_t st = base + stage * C::kStageBytes;
            const uint32_t lead_full = mapa(full_bar(stage), 0);
            if (leader) mbar_expect_tx(full_bar(stage), 2 * C::kStageBytes);
            tma2_load_4d(st, m_hi, c * KC, e.w0 + kw - 1, e.h0 - 1, e.n, lead_full);
            tma2_load_4d(st + C::kABox, m_lo, c * KC, e.w0 + kw - 1, e.h0 - 1, e.n, lead_full);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const uint32_t bt = st + 2 * C::kABox + kh * 2 * C::kBHalf;
              tma2_load_2d(bt, &tm_w_hi, (kh * 3 + kw) * Cin + c * KC, (int)crank * (COUT / 2), lead_full);
              tma2_load_2d(bt + C::kBHalf, &tm_w_lo, (kh * 3 + kw) * Cin + c * KC, (int)crank * (COUT / 2), lead_full);
            }
            if (++stage == C::kStages) { stage = 0; par ^= 1; }
          }
        }
        pr = pr_next;
      }
    }
  } else if (warp > C::kEpiWarps) {
    if (lane == 0 && leader) {
      // ---------------- MMA issuers (leader only) ----------------
      // TWO issuing threads: role 0 issues the hi*hi MMAs and runs the promotion protocol (hh_empty / hh_full), role 1
      // issues the cross terms (hi*lo, lo*hi, in that order per 16 K-elements) into the cross-term accumulator.  One
      // thread issued all twelve MMAs of a k-block at ~53-64 cycles each AND paid every barrier wait (~100-150 cycles
      // even when satisfied) in series: 1337 cycles per k-block for 768 cycles of tensor work
      // (profiles/r02p_conv_timers.log).  Each accumulator is written by one thread only, so the order of
      // accumulation is fixed.  Both threads wait for the operands and both release the stage (two commits).
      // The waits are CTA-scope (default semantics) although the peer arrives on these barriers too: the peer hands
      // over no generic-proxy data through them (operands arrive through the async proxy, accumulator buffers are
      // ordered by the tcgen05 fences).
      // (Splitting the operand ring into an activation ring and a per-k-block weight ring, to issue loads two
      // boxes ahead, was slower: three more barrier waits per box cost more than the exposed latency they removed
      // -- 1630 vs 1337 cycles per k-block, profiles/r02q_conv_split_rings_timers.log.  So was a second "full"
      // barrier per stage for the kh = 1, 2 weight tiles, meant to start a stage's first MMAs before its last bytes
      // have landed: 1192-1217 vs 1071-1107 cycles, profiles/r02s_conv_second_full_barrier_timers.log.  And two
      // hi*hi + two cross-term accumulators instead of three + one: 1118-1171, profiles/r02v_conv_2hh2x_timers.log.)
      // c_format F32 (bit 4), a/b format F16, N >> 3 at 17, M >> 4 at 24 with M = 256 over the pair
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((256u >> 4) << 24);
      const int role = warp - (C::kEpiWarps + 1);
      uint32_t stage = 0, par = 0, hb = 0, hpar = 0;  // hb: hi*hi buffer of the current group, hpar: its use parity
      const uint32_t d_x = tmem_base + 3 * COUT;
      unsigned long long c_sched = 0, c_xe = 0, c_hhe = 0, c_full = 0, c_iss = 0, n_tiles_done = 0;
      const bool timing = dbg != nullptr;
      const long long t_begin = timing ? clock64() : 0;
      for (uint32_t ti = 0;; ++ti) {
        const uint32_t slot = ti & (kSched - 1);
        long long c0 = timing ? clock64() : 0;
        mbar_wait_trap(sched_full(slot), (ti / kSched) & 1);
        int g;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(g) : "r"(ring + slot * (uint32_t)sizeof(TileEntry)) : "memory");
        mbar_arrive(sched_empty(slot));
        if (g < 0) break;
        if (timing) c_sched += (unsigned long long)(clock64() - c0);
        if (role == 1) {
          // the promotion warps read the previous tile's cross terms right after that tile's last commit
          long long x0 = timing ? clock64() : 0;
          mbar_wait_trap(x_empty, (ti & 1) ^ 1);
          tc_fence_after();
          if (timing) c_xe += (unsigned long long)(clock64() - x0);
        }
        int kbi = 0;  // k-block of the tile, in (chunk, kw, kh) order; promotion groups = kGroupKb k-blocks
        for (int sidx = 0; sidx < 3 * chunks; ++sidx) {
          long long w0 = timing ? clock64() : 0;
          mbar_wait_trap(full_bar(stage), par);
          tc_fence_after();
          if (timing) c_full += (unsigned long long)(clock64() - w0);
          const uint64_t d0 = make_desc<KC>(base + stage * C::kStageBytes);
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh, ++kbi) {
            // A: rows kh .. kh+7 of the 10-row box (16 pixels x 128 B per row); B: the tap's weight half
            const uint64_t da0 = d0 + (uint64_t)((kh * kTW * KC * 2) >> 4);
            const uint64_t db0 = d0 + (uint64_t)((2 * C::kABox + kh * 2 * C::kBHalf) >> 4);
            if (role == 0) {
              const uint32_t d_hh = tmem_base + hb * COUT;
              if ((kbi & (C::kGroupKb - 1)) == 0) {
                long long g0 = timing ? clock64() : 0;
                mbar_wait_trap(hh_empty(hb), hpar ^ 1);
                tc_fence_after();
                if (timing) c_hhe += (unsigned long long)(clock64() - g0);
              }
              long long g1 = timing ? clock64() : 0;
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                umma2_f16(d_hh, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, ((kbi & (C::kGroupKb - 1)) | k) ? 1u : 0u);
              if ((kbi & (C::kGroupKb - 1)) == C::kGroupKb - 1 || kbi == nkb - 1) {
                umma2_commit_mc(hh_full(hb), 3);
                if (++hb == 3) { hb = 0; hpar ^= 1; }
              }
              if (timing) c_iss += (unsigned long long)(clock64() - g1);
            } else {
              long long g1 = timing ? clock64() : 0;
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) {
                const uint64_t da_hi = da0 + (uint64_t)(2 * k), da_lo = da_hi + (uint64_t)(C::kABox >> 4);
                const uint64_t db_hi = db0 + (uint64_t)(2 * k), db_lo = db_hi + (uint64_t)(C::kBHalf >> 4);
                umma2_f16(d_x, da_hi, db_lo, idesc, (kbi | k) ? 1u : 0u);
                umma2_f16(d_x, da_lo, db_hi, idesc, 1u);
              }
              if (timing) c_iss += (unsigned long long)(clock64() - g1);
            }
          }
          umma2_commit_mc(empty_bar(stage), 3);
          if (++stage == C::kStages) { stage = 0; par ^= 1; }
        }
        if (role == 1) umma2_commit_mc(x_full, 3);
        ++n_tiles_done;
      }
      if (timing) {
        // [0..9]: the hi*hi thread; [16..23]: the cross-term thread (same fields)
        unsigned long long* o = dbg + (role == 0 ? 0 : 16);
        atomicAdd(o + 0, 2 * n_tiles_done);
        atomicAdd(o + 1, n_tiles_done * (unsigned long long)nkb);
        atomicAdd(o + 2, c_sched);
        atomicAdd(o + 3, c_xe);
        atomicAdd(o + 4, c_hhe);
        atomicAdd(o + 5, c_full);
        atomicAdd(o + 6, c_iss);
        if (role == 0) {
          atomicAdd(dbg + 8, (unsigned long long)(clock64() - t_begin));
          atomicAdd(dbg + 9, 1ull);
        }
      }
    }
  } else {
    // ---------------- promotion + epilogue (both CTAs, own tensor memory) ----------------
    constexpr int CH = 32;
    const int wq = warp & 3, hsel = warp >> 2;
    const uint32_t lane_base = ((uint32_t)(wq * 32) << 16) + (uint32_t)(hsel * CH);
    const int ngroups = (nkb + C::kGroupKb - 1) / C::kGroupKb;
    uint32_t hb = 0, hpar = 0;
    const bool etime = dbg != nullptr && warp == 0 && lane == 0;  // promotion warp 0 of every CTA
    unsigned long long e_ring = 0, e_hhf = 0, e_promo = 0, e_xf = 0, e_epi = 0, e_tiles = 0;
    for (uint32_t ti = 0;; ++ti) {
      const uint32_t slot = ti & (kSched - 1);
      long long q0 = etime ? clock64() : 0;
      mbar_wait_cluster_trap(sched_full(slot), (ti / kSched) & 1);
      if (etime) e_ring += (unsigned long long)(clock64() - q0);
      const TileEntry e = ring_p[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_release(mapa(sched_empty(slot), 0));
      if (e.g < 0) break;
      const bool real_tile = e.pad != 0;
      float acc[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = 0.f;
      for (int g = 0; g < ngroups; ++g) {
        long long q1 = etime ? clock64() : 0;
        mbar_wait_trap(hh_full(hb), hpar);
        tc_fence_after();
        long long q2 = etime ? clock64() : 0;
        {
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_base + hb * COUT, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(__uint_as_float(r[j]), promo_scale, acc[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(mapa(hh_empty(hb), 0));  // no data handed over: the TMEM reads are done
        if (++hb == 3) { hb = 0; hpar ^= 1; }
        if (etime) {
          e_hhf += (unsigned long long)(q2 - q1);
          e_promo += (unsigned long long)(clock64() - q2);
        }
      }
      long long q3 = etime ? clock64() : 0;
      mbar_wait_trap(x_full, ti & 1);
      tc_fence_after();
      long long q4 = etime ? clock64() : 0;
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + 3 * COUT, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_relaxed(mapa(x_empty, 0));

      epilogue_store<COUT>(acc, e, wq, hsel, lane, relu, ph, pw, real_tile, bias, out_hi, out_lo, ovf);
      if (etime) {
        e_xf += (unsigned long long)(q4 - q3);
        e_epi += (unsigned long long)(clock64() - q4);
        ++e_tiles;
      }
    }
    if (etime) {
      atomicAdd(dbg + 10, e_ring);
      atomicAdd(dbg + 11, e_hhf);
      atomicAdd(dbg + 12, e_promo);
      atomicAdd(dbg + 13, e_xf);
      atomicAdd(dbg + 14, e_epi);
      atomicAdd(dbg + 15, e_tiles);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's operands / signal its barriers
  if (warp == C::kEpiWarps) tmem_dealloc2(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// layout converters / pooling (bandwidth-bound helpers)
// ------------------------------------------------------------------------------------------

__global__ void nchw_to_nhwc_split_kernel(const float* __restrict__ x, act_t* __restrict__ hi,
                                          act_t* __restrict__ lo, int C, int64_t HW, int64_t total_pix, int* __restrict__ ovf) {
  // one thread per (pixel, 8-channel group): coalesced reads along pixels, 16-byte writes
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int groups = C / 8;
  int64_t pix = idx % total_pix;
  int g = (int)(idx / total_pix);
  if (g >= groups) return;
  int64_t n = pix / HW, rem = pix - n * HW;
  const float* src = x + (n * C + (int64_t)g * 8) * HW + rem;
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    split2(src[(int64_t)j * HW], src[(int64_t)(j + 1) * HW], ph[j / 2], pl[j / 2], ovf);
  }
  *reinterpret_cast<uint4*>(hi + pix * C + g * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  *reinterpret_cast<uint4*>(lo + pix * C + g * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

__global__ void nhwc_split_to_nchw_kernel(const act_t* __restrict__ hi, const act_t* __restrict__ lo,
                                          float* __restrict__ y, int C, int64_t HW, int64_t total_pix) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int groups = C / 8;
  int64_t pix = idx % total_pix;
  int g = (int)(idx / total_pix);
  if (g >= groups) return;
  int64_t n = pix / HW, rem = pix - n * HW;
  uint4 vh = *reinterpret_cast<const uint4*>(hi + pix * C + g * 8);
  uint4 vl = *reinterpret_cast<const uint4*>(lo + pix * C + g * 8);
  const uint32_t hh[4] = {vh.x, vh.y, vh.z, vh.w}, ll[4] = {vl.x, vl.y, vl.z, vl.w};
  float* dst = y + (n * C + (int64_t)g * 8) * HW + rem;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = join_lo(hh[j], ll[j]);
    float b = join_hi(hh[j], ll[j]);
    dst[(int64_t)(2 * j) * HW] = a;
    dst[(int64_t)(2 * j + 1) * HW] = b;
  }
}

__device__ __forceinline__ int find_group(const RaggedDesc* __restrict__ g, int n, int key) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&g[mid].first) <= key) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// ragged over width groups: block -> group by `first`; one thread per (n, w, 8-channel group): mean over
// h, written at row out_off + w*N + n of the packed [rows, C] feature matrix
__global__ void nhwc_split_avg_to_seq_kernel(const act_t* __restrict__ hi, const act_t* __restrict__ lo,
                                             float* __restrict__ y, int C, const RaggedDesc* __restrict__ groups,
                                             int n_groups) {
  const int gi = find_group(groups, n_groups, blockIdx.x);
  const RaggedDesc d = groups[gi];
  const int N = d.N, H = d.H, W = d.W;
  const int cg = C / 8;
  const int64_t idx = (int64_t)(blockIdx.x - d.first) * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * W * cg) return;
  const int g = (int)(idx % cg);
  const int w = (int)((idx / cg) % W);
  const int n = (int)(idx / ((int64_t)cg * W));
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int h = 0; h < H; ++h) {
    const int64_t pix = d.in_off + ((int64_t)n * H + h) * W + w;
    const uint4 vh = *reinterpret_cast<const uint4*>(hi + pix * C + g * 8);
    const uint4 vl = *reinterpret_cast<const uint4*>(lo + pix * C + g * 8);
    const uint32_t hh[4] = {vh.x, vh.y, vh.z, vh.w}, ll[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] += join_lo(hh[j], ll[j]);
      acc[2 * j + 1] += join_hi(hh[j], ll[j]);
    }
  }
  const float inv = (float)H;
  float4* dst = reinterpret_cast<float4*>(y + (d.out_off + (int64_t)w * N + n) * C + g * 8);
  dst[0] = make_float4(acc[0] / inv, acc[1] / inv, acc[2] / inv, acc[3] / inv);
  dst[1] = make_float4(acc[4] / inv, acc[5] / inv, acc[6] / inv, acc[7] / inv);
}

__global__ void maxpool_nhwc_split_kernel(const act_t* __restrict__ x_hi, const act_t* __restrict__ x_lo,
                                          act_t* __restrict__ y_hi, act_t* __restrict__ y_lo, int N,
                                          int H, int W, int C, int ph, int pw) {
  const int OH = H / ph, OW = W / pw, groups = C / 8;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)N * OH * OW * groups;
  if (idx >= total) return;
  int g = (int)(idx % groups);
  int64_t opix = idx / groups;
  int ow = (int)(opix % OW);
  int oh = (int)((opix / OW) % OH);
  int n = (int)(opix / ((int64_t)OW * OH));
  float best[8];
  uint32_t bh[4] = {0, 0, 0, 0}, bl[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
  for (int r = 0; r < ph; ++r)
    for (int s = 0; s < pw; ++s) {
      int64_t ipix = ((int64_t)n * H + oh * ph + r) * W + ow * pw + s;
      uint4 vh = *reinterpret_cast<const uint4*>(x_hi + ipix * C + g * 8);
      uint4 vl = *reinterpret_cast<const uint4*>(x_lo + ipix * C + g * 8);
      const uint32_t hh[4] = {vh.x, vh.y, vh.z, vh.w}, ll[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = join_lo(hh[j], ll[j]);
        float b = join_hi(hh[j], ll[j]);
        if (a > best[2 * j]) {
          best[2 * j] = a;
          bh[j] = (bh[j] & 0xFFFF0000u) | (hh[j] & 0xFFFFu);
          bl[j] = (bl[j] & 0xFFFF0000u) | (ll[j] & 0xFFFFu);
        }
        if (b > best[2 * j + 1]) {
          best[2 * j + 1] = b;
          bh[j] = (bh[j] & 0xFFFFu) | (hh[j] & 0xFFFF0000u);
          bl[j] = (bl[j] & 0xFFFFu) | (ll[j] & 0xFFFF0000u);
        }
      }
    }
  *reinterpret_cast<uint4*>(y_hi + opix * C + g * 8) = make_uint4(bh[0], bh[1], bh[2], bh[3]);
  *reinterpret_cast<uint4*>(y_lo + opix * C + g * 8) = make_uint4(bl[0], bl[1], bl[2], bl[3]);
}

// one thread = one pooled output pixel, all Cout channels (weights broadcast from shared memory);
// ragged over width groups (block -> group by `first`)
template <int COUT>
__global__ void __launch_bounds__(128)
stem_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
            act_t* __restrict__ out_hi, act_t* __restrict__ out_lo, const RaggedDesc* __restrict__ groups, int n_groups,
            int* __restrict__ ovf) {
  __shared__ float sw[COUT * 9 + COUT];
  for (int i = threadIdx.x; i < COUT * 10; i += blockDim.x) sw[i] = i < COUT * 9 ? wgt[i] : bias[i - COUT * 9];
  __syncthreads();
  const int gi = find_group(groups, n_groups, blockIdx.x);
  const RaggedDesc d = groups[gi];
  const int H = d.H, W = d.W;
  const int OH = H / 2, OW = W / 2;
  const int64_t idx = (int64_t)(blockIdx.x - d.first) * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)d.N * OH * OW) return;
  const int ow = (int)(idx % OW), oh = (int)((idx / OW) % OH), n = (int)(idx / ((int64_t)OW * OH));
  const float* xi = x + d.in_off + (int64_t)n * H * W;
  float p[4][4];  // input patch rows 2oh-1 .. 2oh+2, cols 2ow-1 .. 2ow+2 (zero padded)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ih = 2 * oh - 1 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int iw = 2 * ow - 1 + c;
      p[r][c] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(xi + (int64_t)ih * W + iw) : 0.f;
    }
  }
  act_t* oh_ptr = out_hi + (d.out_off + idx) * COUT;
  act_t* ol_ptr = out_lo + (d.out_off + idx) * COUT;
#pragma unroll 1
  for (int c8 = 0; c8 < COUT; c8 += 8) {
    uint32_t ph[4], pl[4];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* k = sw + (c8 + j) * 9;
      float best = -INFINITY;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float a = 0.f;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) a = fmaf(p[dy + r][dx + s], k[r * 3 + s], a);
          best = fmaxf(best, a);
        }
      v[j] = fmaxf(best + sw[COUT * 9 + c8 + j], 0.f);
    }
    uint32_t bad = 0u;  // packed conversions (two values per F2FP) instead of scalar F2F at a quarter of the rate
#pragma unroll
    for (int j = 0; j < 4; ++j) bad |= split2_packed(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
    if (bad) *ovf = 1;
    *reinterpret_cast<uint4*>(oh_ptr + c8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(ol_ptr + c8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  OCRS_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev != cached_dev) {
    OCRS_CUDA_CHECK(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
    cached_dev = dev;
  }
  return cached;
}

// Scale applied to every promoted hi*hi partial sum.  1 + kappa * 0.5 * 8 * 2^-23 would undo the
// expected truncation of the 8 MMAs behind a partial (see the kernel comment); kappa = 0.5 by
// default (measured: halves the remaining mean error of every layer); OCRS_B200_TC_DEBIAS=0 turns it off.
float promo_scale() {
  static const float s = [] {
    const char* e = std::getenv("OCRS_B200_TC_DEBIAS");
    const double kappa = e ? std::atof(e) : 0.5;
    return (float)(1.0 + kappa * 0.5 * 8.0 / 8388608.0);
  }();
  return s;
}

// MMA issue order inside a promotion group: 0 (default) = interleaved per 16 K-elements, 1 = all hi*hi
// first (OCRS_B200_CONV_HH_FIRST=1); same arithmetic either way
bool conv_hh_first() {
  static const bool on = [] { const char* e = std::getenv("OCRS_B200_CONV_HH_FIRST"); return e != nullptr && e[0] == '1'; }();
  return on;
}

bool conv_debug() {
  static const bool on = std::getenv("OCRS_B200_CONV_DEBUG") != nullptr;
  return on;
}

// 128-channel layers: 0 = single-CTA kernel, 1 = CTA pair, 2 (default) = CTA pair + halo reuse (its activation
// maps have a 10-row box).  OCRS_B200_CONV_MODE selects; the other layers always run the single-CTA kernel.
int conv_mode() {
  static const int m = [] {
    const char* e = std::getenv("OCRS_B200_CONV_MODE");
    if (e != nullptr) return std::atoi(e);
    const char* p = std::getenv("OCRS_B200_CONV_PAIR");
    return (p != nullptr && p[0] == '1') ? 1 : 2;
  }();
  return m;
}


template <int KC, int COUT, bool HALO>
struct PairSel {
  using Cfg = PairCfg<COUT>;
  static auto kernel() { return conv3x3_pair_kernel<COUT>; }
};
template <int KC, int COUT>
struct PairSel<KC, COUT, true> {
  using Cfg = HaloCfg<KC, COUT>;
  static auto kernel() { return conv3x3_halo_kernel<KC, COUT>; }
};

template <int KC, int COUT, bool HALO>
void launch_conv_pair(const CUtensorMap* d_maps, const RaggedDesc* d_groups, int n_groups, int n_tiles, int* d_counter,
                      const ConvWeightsTC& w, act_t* y_hi, act_t* y_lo, int relu, int ph, int pw, int* ovf, cudaStream_t st) {
  using C = typename PairSel<KC, COUT, HALO>::Cfg;
  static_assert(HALO || KC == 64, "the plain CTA-pair kernel is built for 64-channel k-blocks");
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const int Cin = w.Cin;
  uint64_t wd[2] = {(uint64_t)9 * Cin, (uint64_t)COUT};
  uint64_t ws[1] = {(uint64_t)9 * Cin * 2};
  uint32_t wb[2] = {(uint32_t)C::kKC, (uint32_t)(COUT / 2)};  // each CTA loads its half of the output channels
  CUtensorMap tm_w_hi = make_map(w.w_hi.ptr, 2, wd, ws, wb, swz);
  CUtensorMap tm_w_lo = make_map(w.w_lo.ptr, 2, wd, ws, wb, swz);
  auto kern = PairSel<KC, COUT, HALO>::kernel();
  OCRS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  const int n_pairs = (n_tiles + 1) / 2;
  const int grid = 2 * std::max(1, std::min(n_pairs, sm_count() / 2));
  unsigned long long* d_dbg = nullptr;
  if (conv_debug()) {
    OCRS_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_dbg), 24 * sizeof(unsigned long long), st));
    OCRS_CUDA_CHECK(cudaMemsetAsync(d_dbg, 0, 24 * sizeof(unsigned long long), st));
  }
  kern<<<grid, C::kThreads, C::kSmemBytes, st>>>(tm_w_hi, tm_w_lo, d_maps, d_groups, n_groups, n_tiles, d_counter,
                                                  w.bias.as<float>(), y_hi, y_lo, Cin, relu, ph, pw, promo_scale(), ovf, d_dbg);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
  if (d_dbg) {
    unsigned long long h[24];
    OCRS_CUDA_CHECK(cudaMemcpyAsync(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    OCRS_CUDA_CHECK(cudaStreamSynchronize(st));
    OCRS_CUDA_CHECK(cudaFreeAsync(d_dbg, st));
    const double kb = (double)std::max<unsigned long long>(h[1], 1), cl = (double)std::max<unsigned long long>(h[9], 1);
    fprintf(stderr,
            "[conv dbg PAIR%s] Cin %d Cout %d pool %dx%d: %llu tiles, %d groups, %.0f clusters | per k-block of M=256 (cycles): ring %.0f, "
            "x_empty %.0f, hh_empty %.0f, operands %.0f, issue %.0f | loop total %.0f (per cluster %.0f cycles)",
            HALO ? "+HALO" : "", Cin, COUT, ph, pw, h[0], n_groups, cl, h[2] / kb, h[3] / kb, h[4] / kb, h[5] / kb, h[6] / kb, h[8] / kb,
            h[8] / cl);
    if (h[17]) {
      const double kx = (double)h[17];
      fprintf(stderr, " | cross-term thread: ring %.0f, x_empty %.0f, operands %.0f, issue %.0f", h[18] / kx, h[19] / kx, h[21] / kx, h[22] / kx);
    }
    if (h[15]) {
      const double tl = (double)h[15];
      fprintf(stderr, " | promotion warp 0 of every CTA, per TILE: ring %.0f, wait hh_full %.0f, promote %.0f, wait x_full %.0f, X + epilogue %.0f",
              h[10] / tl, h[11] / tl, h[12] / tl, h[13] / tl, h[14] / tl);
    }
    fprintf(stderr, "\n");
  }
}

void launch_conv_ws(const CUtensorMap* d_maps, const RaggedDesc* d_groups, int n_groups, int n_tiles, int* d_counter,
                    const ConvWeightsTC& w, act_t* y_hi, act_t* y_lo, int relu, int* ovf, cudaStream_t st) {
  using C = WsCfg;
  OCRS_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  const int grid = std::max(1, std::min(n_tiles, sm_count()));
  // expected truncation of the 18 MMAs behind a hi*hi sum, like promo_scale() for the 8 behind a promotion group
  const float debias = 1.0f + (promo_scale() - 1.0f) * (18.0f / 8.0f);
  conv3x3_ws_kernel<<<grid, C::kThreads, C::kSmemBytes, st>>>(d_maps, d_groups, n_groups, n_tiles, d_counter, w.w_hi.as<act_t>(),
                                                              w.w_lo.as<act_t>(), w.bias.as<float>(), y_hi, y_lo, relu, debias, ovf);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

template <int KC, int COUT>
void launch_conv_res(const CUtensorMap* d_maps, const RaggedDesc* d_groups, int n_groups, int n_tiles, int* d_counter,
                     const ConvWeightsTC& w, act_t* y_hi, act_t* y_lo, int relu, int ph, int pw, int* ovf, cudaStream_t st) {
  using C = ResCfg<KC, COUT>;
  const int Cin = w.Cin;
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  uint64_t wd[2] = {(uint64_t)9 * Cin, (uint64_t)COUT};
  uint64_t ws[1] = {(uint64_t)9 * Cin * 2};
  uint32_t wb[2] = {(uint32_t)KC, (uint32_t)COUT};
  CUtensorMap tm_w_hi = make_map(w.w_hi.ptr, 2, wd, ws, wb, swz);
  CUtensorMap tm_w_lo = make_map(w.w_lo.ptr, 2, wd, ws, wb, swz);
  OCRS_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_res_kernel<KC, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  const int grid = std::max(1, std::min(n_tiles, sm_count()));
  unsigned long long* d_dbg = nullptr;
  if (conv_debug()) {
    OCRS_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_dbg), 40 * sizeof(unsigned long long), st));
    OCRS_CUDA_CHECK(cudaMemsetAsync(d_dbg, 0, 40 * sizeof(unsigned long long), st));
  }
  conv3x3_res_kernel<KC, COUT><<<grid, C::kThreads, C::kSmemBytes, st>>>(tm_w_hi, tm_w_lo, d_maps, d_groups, n_groups, n_tiles, d_counter,
                                                                         w.bias.as<float>(), y_hi, y_lo, Cin, relu, ph, pw, promo_scale(),
                                                                         ovf, d_dbg);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
  if (d_dbg) {
    unsigned long long h[40];
    OCRS_CUDA_CHECK(cudaMemcpyAsync(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    OCRS_CUDA_CHECK(cudaStreamSynchronize(st));
    OCRS_CUDA_CHECK(cudaFreeAsync(d_dbg, st));
    const double tl = (double)std::max<unsigned long long>(h[16], 1);
    fprintf(stderr,
            "[conv dbg RES] Cin %d Cout %d pool %dx%d: %llu tiles, %d CTAs | per TILE (cycles): producer: ring slot %.0f, decode %.0f, "
            "stage waits %.0f, total %.0f",
            Cin, COUT, ph, pw, h[16], grid, h[10] / tl, h[11] / tl, h[12] / tl, h[13] / tl);
    const char* names[3] = {"hi*hi", "hi*lo", "lo*hi"};
    for (int r = 0; r < 3; ++r) {
      const unsigned long long* o = h + 16 + 8 * r;
      fprintf(stderr, " | %s thread: ring %.0f, x_empty %.0f, hh_empty %.0f, operands %.0f, total %.0f", names[r], o[1] / tl, o[2] / tl,
              o[3] / tl, o[4] / tl, o[5] / tl);
    }
    fprintf(stderr, "\n");
  }
}

template <int KC, int COUT>
void launch_conv(const CUtensorMap* d_maps, const RaggedDesc* d_groups, int n_groups, int n_tiles, int* d_counter,
                 const ConvWeightsTC& w, act_t* y_hi, act_t* y_lo, int relu, int ph, int pw, int* ovf, cudaStream_t st) {
  using C = Cfg<KC, COUT>;
  const int Cin = w.Cin;
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  uint64_t wd[2] = {(uint64_t)9 * Cin, (uint64_t)COUT};
  uint64_t ws[1] = {(uint64_t)9 * Cin * 2};
  uint32_t wb[2] = {(uint32_t)KC, (uint32_t)COUT};
  CUtensorMap tm_w_hi = make_map(w.w_hi.ptr, 2, wd, ws, wb, swz);
  CUtensorMap tm_w_lo = make_map(w.w_lo.ptr, 2, wd, ws, wb, swz);
  OCRS_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc_kernel<KC, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::kSmemBytes));  // per-device attribute, cheap to repeat
  // persistent CTAs, launched as clusters of 2 (no cluster-level protocol: the CTAs are independent; the
  // pair sits on the two SMs of one TPC, measured 4-7 % faster per layer than without the attribute)
  const int grid = std::max(2, std::min((n_tiles + 1) / 2 * 2, sm_count() / 2 * 2));
  unsigned long long* d_dbg = nullptr;
  if (conv_debug()) {
    OCRS_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_dbg), 16 * sizeof(unsigned long long), st));
    OCRS_CUDA_CHECK(cudaMemsetAsync(d_dbg, 0, 16 * sizeof(unsigned long long), st));
  }
  conv3x3_tc_kernel<KC, COUT><<<grid, C::kThreads, C::kSmemBytes, st>>>(
      tm_w_hi, tm_w_lo, d_maps, d_groups, n_groups, n_tiles, d_counter, w.bias.as<float>(), y_hi, y_lo, Cin, relu, ph, pw,
      promo_scale(), ovf, d_dbg, conv_hh_first() ? 1 : 0);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
  if (d_dbg) {
    unsigned long long h[16];
    OCRS_CUDA_CHECK(cudaMemcpyAsync(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    OCRS_CUDA_CHECK(cudaStreamSynchronize(st));
    OCRS_CUDA_CHECK(cudaFreeAsync(d_dbg, st));
    const double kb = (double)std::max<unsigned long long>(h[1], 1), ctas = (double)std::max<unsigned long long>(h[9], 1);
    const double tl = (double)std::max<unsigned long long>(h[0], 1);
    fprintf(stderr,
            "[conv dbg] Cin %d Cout %d pool %dx%d: %llu tiles, %d groups, %.0f CTAs | per k-block (cycles): ring %.0f, x_empty %.0f, "
            "hh_empty %.0f, operands %.0f, issue HH %.0f, issue X %.0f | loop total %.0f (per CTA %.0f cycles) | promotion warp 0 of "
            "every CTA, per TILE: ring %.0f, wait hh_full %.0f, promote %.0f, wait x_full %.0f, X + epilogue %.0f\n",
            Cin, COUT, ph, pw, h[0], n_groups, ctas, h[2] / kb, h[3] / kb, h[4] / kb, h[5] / kb, h[6] / kb, h[7] / kb, h[8] / kb,
            h[8] / ctas, h[10] / tl, h[11] / tl, h[12] / tl, h[13] / tl, h[14] / tl);
  }
}

}  // namespace

bool available() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0;
    cudaDeviceProp prop{};
    bool ok = cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess &&
              prop.major == 10 && get_encode() != nullptr;
    if (!ok) cudaGetLastError();
    cached = ok ? 1 : 0;
  }
  return cached == 1;
}

bool conv_supported(int Cin, int Cout, int R, int S, int stride_h, int stride_w, int pad_t, int pad_l, int pad_b,
                    int pad_r, int dil_h, int dil_w, int groups) {
  return R == 3 && S == 3 && stride_h == 1 && stride_w == 1 && pad_t == 1 && pad_l == 1 && pad_b == 1 && pad_r == 1 &&
         dil_h == 1 && dil_w == 1 && groups == 1 && (Cin == 32 || Cin % 64 == 0) && (Cout == 64 || Cout == 128);
}

bool stem_supported(int Cin, int Cout, int R, int S, int stride_h, int stride_w, int pad_t, int pad_l, int pad_b,
                    int pad_r, int dil_h, int dil_w, int groups) {
  return Cin == 1 && R == 3 && S == 3 && stride_h == 1 && stride_w == 1 && pad_t == 1 && pad_l == 1 && pad_b == 1 &&
         pad_r == 1 && dil_h == 1 && dil_w == 1 && groups == 1 && (Cout == 32 || Cout == 64);
}

std::unique_ptr<StemWeights> prepare_stem(const float* w, const float* b, int Cout) {
  auto s = std::make_unique<StemWeights>();
  s->Cout = Cout;
  std::vector<float> bias((size_t)Cout, 0.f);
  if (b) bias.assign(b, b + Cout);
  s->w.reserve((size_t)Cout * 9 * 4);
  s->bias.reserve((size_t)Cout * 4);
  OCRS_CUDA_CHECK(cudaMemcpy(s->w.ptr, w, (size_t)Cout * 9 * 4, cudaMemcpyHostToDevice));
  OCRS_CUDA_CHECK(cudaMemcpy(s->bias.ptr, bias.data(), (size_t)Cout * 4, cudaMemcpyHostToDevice));
  return s;
}

void stem_ragged(const float* x, const StemWeights& w, act_t* y_hi, act_t* y_lo, const RaggedDesc* d_groups, int n_groups,
                 int n_blocks, int* ovf, cudaStream_t st) {
  if (n_blocks == 0) return;
  if (w.Cout == 32)
    stem_kernel<32><<<(unsigned)n_blocks, 128, 0, st>>>(x, w.w.as<float>(), w.bias.as<float>(), y_hi, y_lo, d_groups, n_groups, ovf);
  else
    stem_kernel<64><<<(unsigned)n_blocks, 128, 0, st>>>(x, w.w.as<float>(), w.bias.as<float>(), y_hi, y_lo, d_groups, n_groups, ovf);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

namespace {
// stream-ordered upload of a small host blob (pageable source: the runtime stages it before returning)
void* upload_async(const void* host, size_t bytes, cudaStream_t st) {
  void* d = nullptr;
  OCRS_CUDA_CHECK(cudaMallocAsync(&d, bytes, st));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, st));
  return d;
}
}  // namespace

void stem_conv_relu_pool2(const float* x, const StemWeights& w, act_t* y_hi, act_t* y_lo, int N, int H,
                          int W, int* ovf, cudaStream_t st) {
  int64_t total = (int64_t)N * (H / 2) * (W / 2);
  if (!total) return;
  RaggedDesc d{};
  d.N = N; d.H = H; d.W = W; d.OH = H / 2; d.OW = W / 2;
  auto* dd = static_cast<RaggedDesc*>(upload_async(&d, sizeof(d), st));
  stem_ragged(x, w, y_hi, y_lo, dd, 1, (int)ceil_div(total, 128), ovf, st);
  OCRS_CUDA_CHECK(cudaFreeAsync(dd, st));
}

std::unique_ptr<ConvWeightsTC> prepare_weights(const float* w, const float* b, int Cin, int Cout) {
  auto out = std::make_unique<ConvWeightsTC>();
  out->Cin = Cin;
  out->Cout = Cout;
  const size_t K = (size_t)9 * Cin;
  std::vector<act_t> hi(K * Cout), lo(K * Cout);
  for (int co = 0; co < Cout; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          float v = w[(((size_t)co * Cin + ci) * 3 + kh) * 3 + kw];
          OCRS_CHECK(std::fabs(v) <= 65504.f, kModelLoad, "conv weight exceeds the fp16 range of the tensor-core path");
          act_t h = __float2half_rn(v);
          act_t l = __float2half_rn(v - __half2float(h));
          size_t k = (size_t)(kh * 3 + kw) * Cin + ci;
          hi[(size_t)co * K + k] = h;
          lo[(size_t)co * K + k] = l;
        }
  std::vector<float> bias((size_t)Cout, 0.f);
  if (b) bias.assign(b, b + Cout);
  out->w_hi.reserve(hi.size() * 2);
  out->w_lo.reserve(lo.size() * 2);
  out->bias.reserve(bias.size() * 4);
  OCRS_CUDA_CHECK(cudaMemcpy(out->w_hi.ptr, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
  OCRS_CUDA_CHECK(cudaMemcpy(out->w_lo.ptr, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
  OCRS_CUDA_CHECK(cudaMemcpy(out->bias.ptr, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
  return out;
}

int conv_tiles(int N, int H, int W) { return N * (int)ceil_div(H, kTH) * (int)ceil_div(W, kTW); }

void conv_fill_tiles(RaggedDesc* d) {
  d->tiles_w = (int)ceil_div(d->W, kTW);
  d->tiles_h = (int)ceil_div(d->H, kTH);
}

// the layers the CTA-pair + halo kernel runs: the 128-channel ones.  (Instantiated for the 32 -> 64 layer it was no
// faster per SM, and with three batches in flight the benchmark stopped making progress, cause not found:
// profiles/r02n_halo32.md.  That layer runs the resident-weights kernel instead.)
bool conv_uses_halo(int Cin, int Cout) { return conv_mode() == 2 && Cin % 64 == 0 && Cout == 128; }
// the 32 -> 64 layer with a fused 2x2 pool: the weights-stationary transposed kernel (same 8 x 16 tiles and 10-row boxes
// as the resident-weights kernel; OCRS_B200_CONV_WS=0 selects that one)
bool conv_ws_enabled() {
  static const bool on = [] { const char* e = std::getenv("OCRS_B200_CONV_WS"); return e == nullptr || e[0] != '0'; }();
  return on;
}
// the 32 -> 64 layer: resident weights + 10-row activation boxes (OCRS_B200_CONV_RES=0: the streaming kernel)
bool conv_uses_res(int Cin, int Cout) {
  static const bool on = [] { const char* e = std::getenv("OCRS_B200_CONV_RES"); return e == nullptr || e[0] != '0'; }();
  return on && Cin == 32 && Cout == 64;
}

void make_act_maps(const act_t* hi, const act_t* lo, int N, int H, int W, int Cin, CUtensorMap out[2], int Cout) {
  const int KC = (Cin % 64 == 0) ? 64 : 32;
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  uint64_t xd[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t xs[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
  // the halo kernel loads 10-row boxes (rows h0-1 .. h0+8), the others 8-row boxes per tap
  const bool halo = conv_uses_halo(Cin, Cout) || conv_uses_res(Cin, Cout);
  uint32_t xb[4] = {(uint32_t)KC, (uint32_t)kTW, (uint32_t)(halo ? kTH + 2 : kTH), 1};
  out[0] = make_map(hi, 4, xd, xs, xb, swz);
  out[1] = make_map(lo, 4, xd, xs, xb, swz);
}

void conv3x3_ragged(const CUtensorMap* d_maps, const RaggedDesc* d_groups, int n_groups, int n_tiles, int* d_counter,
                    const ConvWeightsTC& w, act_t* y_hi, act_t* y_lo, int relu, int ph, int pw, int* ovf, cudaStream_t st) {
  if (n_tiles == 0 || n_groups == 0) return;
  OCRS_CHECK((ph == 1 || ph == 2) && (pw == 1 || pw == 2), kInternal, "conv3x3 (tensor core): fused pool must be 1 or 2");
  const bool k64 = (w.Cin % 64 == 0);
  if (conv_uses_res(w.Cin, w.Cout) && ph == 2 && pw == 2 && conv_ws_enabled()) {
    launch_conv_ws(d_maps, d_groups, n_groups, n_tiles, d_counter, w, y_hi, y_lo, relu, ovf, st);
    return;
  }
  if (conv_uses_res(w.Cin, w.Cout)) {
    launch_conv_res<32, 64>(d_maps, d_groups, n_groups, n_tiles, d_counter, w, y_hi, y_lo, relu, ph, pw, ovf, st);
    return;
  }
  if (conv_uses_halo(w.Cin, w.Cout)) {
    launch_conv_pair<64, 128, true>(d_maps, d_groups, n_groups, n_tiles, d_counter, w, y_hi, y_lo, relu, ph, pw, ovf, st);
    return;
  }
  if (k64 && w.Cout == 128 && conv_mode() == 1) {
    launch_conv_pair<64, 128, false>(d_maps, d_groups, n_groups, n_tiles, d_counter, w, y_hi, y_lo, relu, ph, pw, ovf, st);
    return;
  }
  if (k64 && w.Cout == 128) launch_conv<64, 128>(d_maps, d_groups, n_groups, n_tiles, d_counter, w, y_hi, y_lo, relu, ph, pw, ovf, st);
  else if (k64 && w.Cout == 64) launch_conv<64, 64>(d_maps, d_groups, n_groups, n_tiles, d_counter, w, y_hi, y_lo, relu, ph, pw, ovf, st);
  else if (!k64 && w.Cout == 128) launch_conv<32, 128>(d_maps, d_groups, n_groups, n_tiles, d_counter, w, y_hi, y_lo, relu, ph, pw, ovf, st);
  else if (!k64 && w.Cout == 64) launch_conv<32, 64>(d_maps, d_groups, n_groups, n_tiles, d_counter, w, y_hi, y_lo, relu, ph, pw, ovf, st);
  else throw Error(kInternal, "conv3x3 (tensor core): unsupported channel configuration");
}

void conv3x3(const act_t* x_hi, const act_t* x_lo, const ConvWeightsTC& w, act_t* y_hi,
             act_t* y_lo, int N, int H, int W, int relu, int ph, int pw, int* ovf, cudaStream_t st) {
  if (N == 0 || H == 0 || W == 0) return;
  // one-group ragged launch: [maps hi, lo | desc | counter] in one stream-ordered blob
  struct Blob { CUtensorMap maps[2]; RaggedDesc d; int counter; int pad[3]; } blob;
  make_act_maps(x_hi, x_lo, N, H, W, w.Cin, blob.maps, w.Cout);
  blob.d = RaggedDesc{};
  blob.d.N = N; blob.d.H = H; blob.d.W = W; blob.d.OH = H / ph; blob.d.OW = W / pw;
  conv_fill_tiles(&blob.d);
  blob.counter = 0;
  auto* dev = static_cast<Blob*>(upload_async(&blob, sizeof(blob), st));
  conv3x3_ragged(dev->maps, &dev->d, 1, conv_tiles(N, H, W), &dev->counter, w, y_hi, y_lo, relu, ph, pw, ovf, st);
  OCRS_CUDA_CHECK(cudaFreeAsync(dev, st));
}

void nchw_to_nhwc_split(const float* x, act_t* hi, act_t* lo, int N, int C, int H, int W, int* ovf,
                        cudaStream_t st) {
  OCRS_CHECK(C % 8 == 0, kInternal, "nchw_to_nhwc_split: C must be a multiple of 8");
  int64_t HW = (int64_t)H * W, total_pix = (int64_t)N * HW, total = total_pix * (C / 8);
  if (!total) return;
  nchw_to_nhwc_split_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(x, hi, lo, C, HW, total_pix, ovf);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void nhwc_split_to_nchw(const act_t* hi, const act_t* lo, float* y, int N, int C, int H, int W,
                        cudaStream_t st) {
  OCRS_CHECK(C % 8 == 0, kInternal, "nhwc_split_to_nchw: C must be a multiple of 8");
  int64_t HW = (int64_t)H * W, total_pix = (int64_t)N * HW, total = total_pix * (C / 8);
  if (!total) return;
  nhwc_split_to_nchw_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(hi, lo, y, C, HW, total_pix);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void avg_to_seq_ragged(const act_t* hi, const act_t* lo, float* y, int C, const RaggedDesc* d_groups, int n_groups,
                       int n_blocks, cudaStream_t st) {
  OCRS_CHECK(C % 8 == 0, kInternal, "nhwc_split_avg_to_seq: C must be a multiple of 8");
  if (n_blocks == 0) return;
  nhwc_split_avg_to_seq_kernel<<<(unsigned)n_blocks, 256, 0, st>>>(hi, lo, y, C, d_groups, n_groups);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

void nhwc_split_avg_to_seq(const act_t* hi, const act_t* lo, float* y, int N, int C, int H, int W, cudaStream_t st) {
  int64_t total = (int64_t)N * W * (C / 8);
  if (!total) return;
  RaggedDesc d{};
  d.N = N; d.H = H; d.W = W; d.OH = 1; d.OW = W;
  auto* dd = static_cast<RaggedDesc*>(upload_async(&d, sizeof(d), st));
  avg_to_seq_ragged(hi, lo, y, C, dd, 1, (int)ceil_div(total, 256), st);
  OCRS_CUDA_CHECK(cudaFreeAsync(dd, st));
}

void maxpool_nhwc_split(const act_t* x_hi, const act_t* x_lo, act_t* y_hi,
                        act_t* y_lo, int N, int H, int W, int C, int ph, int pw, cudaStream_t st) {
  OCRS_CHECK(C % 8 == 0, kInternal, "maxpool_nhwc_split: C must be a multiple of 8");
  int64_t total = (int64_t)N * (H / ph) * (W / pw) * (C / 8);
  if (!total) return;
  maxpool_nhwc_split_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(x_hi, x_lo, y_hi, y_lo, N, H, W, C, ph, pw);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

}  // namespace tc
}  // namespace ocrs
