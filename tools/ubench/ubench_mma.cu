// Micro-benchmark (development aid): latency of one GRU-step-sized batch of tcgen05 MMAs,
//   48 x (M = 128, N, K = 16) bf16 -> one accumulator, issue + commit + mbarrier wait,
// as a function of N, of where A lives (shared memory descriptor vs tensor memory), of how the
// descriptors are produced (rebuilt per MMA vs base + constant add) and of the number of issuing
// warps (each warp issues 48 / W MMAs into its own accumulator and commits to its own barrier).
// One CTA; operands are zeros.  Every wait is bounded.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../ocrs_b200/csrc/tc_ptx.cuh"

using namespace ocrs::tc::ptx;

__device__ __forceinline__ bool wait_bounded(uint32_t bar, uint32_t parity, int* err) {
  for (int i = 0; i < 2000000; ++i) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  *err = 1;
  return false;
}

template <int N, int A_TMEM, int ADDS, int W>
__global__ void __launch_bounds__(256, 1) mma_kernel(int iters, long long* out, int* err) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // B: [N][256] bf16 hi plane + lo plane, 4 K-subtiles of 64 (SW128): N*128 B each
  const uint32_t kSub = N * 128, kPlane = 4 * kSub;
  const uint32_t b_base = base;
  const uint32_t a_base = base + 2 * kPlane;           // A in smem: [128][256] hi + lo = 2 * 64 KB
  const uint32_t bar = a_base + (A_TMEM ? 0 : 2 * 65536);
  const uint32_t tmem_slot = bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t o = threadIdx.x * 16; o < (bar - base); o += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0) : "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar + 8 * i, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (warp < W && lane == 0) {
    const uint32_t d = tmem_base + warp * N;            // accumulators at columns [0, W*N) (<= 256)
    const uint32_t my_bar = bar + 8 * warp;
    constexpr int per = 16 / W;                          // k16 steps per warp (3 MMAs each)
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t db0 = make_desc<64>(b_base);
      const uint64_t da0 = make_desc<64>(a_base);
#pragma unroll
      for (int q = 0; q < per; ++q) {
        const int kq = warp * per + q;                   // 0..15: kk = kq / 4, k = kq % 4
        const int kk = kq >> 2, k = kq & 3;
        uint64_t db_hi, db_lo, da_hi = 0, da_lo = 0;
        if (ADDS) {
          db_hi = db0 + (uint64_t)((kk * kSub + k * 32) >> 4);
          db_lo = db_hi + (uint64_t)(kPlane >> 4);
          if (!A_TMEM) { da_hi = da0 + (uint64_t)((kk * 16384 + k * 32) >> 4); da_lo = da_hi + (uint64_t)(65536 >> 4); }
        } else {
          db_hi = make_desc<64>(b_base + kk * kSub + k * 32);
          db_lo = make_desc<64>(b_base + kPlane + kk * kSub + k * 32);
          if (!A_TMEM) { da_hi = make_desc<64>(a_base + kk * 16384 + k * 32); da_lo = make_desc<64>(a_base + 65536 + kk * 16384 + k * 32); }
        }
        if (A_TMEM) {
          const uint32_t ta_hi = tmem_base + 256u + (uint32_t)(kq * 8), ta_lo = ta_hi + 128u;
          umma_bf16_ts(d, ta_hi, db_hi, idesc, q ? 1u : 0u);
          umma_bf16_ts(d, ta_hi, db_lo, idesc, 1u);
          umma_bf16_ts(d, ta_lo, db_hi, idesc, 1u);
        } else {
          umma_bf16(d, da_hi, db_hi, idesc, q ? 1u : 0u);
          umma_bf16(d, da_hi, db_lo, idesc, 1u);
          umma_bf16(d, da_lo, db_hi, idesc, 1u);
        }
      }
      umma_commit(my_bar);
      if (!wait_bounded(my_bar, it & 1, err)) break;
      tc_fence_after();
    }
    t1 = clock64();
    if (warp == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N, int A_TMEM, int ADDS, int W>
void run(int iters) {
  long long* d_out; int* d_err;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
  size_t smem = 1024 + 2 * 4 * N * 128 + (A_TMEM ? 0 : 2 * 65536) + 256;
  cudaFuncSetAttribute(mma_kernel<N, A_TMEM, ADDS, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mma_kernel<N, A_TMEM, ADDS, W><<<1, 256, smem>>>(10, d_out, d_err);
  mma_kernel<N, A_TMEM, ADDS, W><<<1, 256, smem>>>(iters, d_out, d_err);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0; int err = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
  printf("N %3d  A in %s  desc %s  issuing warps %d: %7.1f cycles per 48-MMA step (%5.1f per MMA of warp 0's %d)  [%s%s]\n", N,
         A_TMEM ? "TMEM" : "smem", ADDS ? "adds   " : "rebuilt", W, (double)cyc / iters, (double)cyc / iters / (48 / W), 48 / W,
         cudaGetErrorString(e), err ? ", WATCHDOG" : "");
  cudaFree(d_out); cudaFree(d_err);
}

int main() {
  run<32, 1, 0, 1>(500);
  run<32, 1, 1, 1>(500);
  run<32, 0, 1, 1>(500);
  run<64, 1, 1, 1>(500);
  run<128, 1, 1, 1>(500);
  run<32, 1, 1, 2>(500);
  run<32, 1, 1, 4>(500);
  run<64, 1, 1, 2>(500);
  run<64, 1, 1, 4>(500);
  return 0;
}
