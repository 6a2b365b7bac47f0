// Micro-benchmark (development aid, not part of the library): cost of the per-step all-gather of the
// GRU cluster kernel.  Each of the 8 CTAs of a cluster sends a slice of `kb` KB to every CTA of the
// cluster per iteration; three transports:
//   0  st.shared::cluster.v4 from 256 threads + fence.acq_rel.cluster + relaxed remote mbarrier arrives
//      (what gru_cluster_kernel does today)
//   1  cp.async.bulk.shared::cluster.shared::cta with mbarrier::complete_tx on the DESTINATION's barrier
//      (one bulk copy per destination, issued by 8 threads)
//   2  like 1, issued by one thread
//   3  like 1 with double-buffered staging and no cluster barrier per iteration (the real protocol)
// Prints cycles per iteration (clock64 of CTA 0, thread 0).  Every wait is bounded (watchdog).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity, int* err) {
  for (int i = 0; i < 4000000; ++i) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  *err = 1;
  return false;
}

constexpr int kCl = 8;

template <int MODE>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(512, 1)
xchg_kernel(int iters, int slice_bytes, long long* out, int* err) {
  extern __shared__ __align__(128) uint8_t smem[];
  // layout: [2 buffers x kCl slices x slice_bytes] | staging [slice_bytes] | barriers
  const uint32_t base = smem_u32(smem);
  const uint32_t buf_bytes = kCl * slice_bytes;
  const uint32_t stage = base + 2 * buf_bytes;
  const uint32_t bar0 = stage + 2 * slice_bytes;  // 2 barriers (staging is double-buffered for mode 3)
  const uint32_t rank = ctarank();
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      uint32_t cnt = MODE == 0 ? kCl : 1;
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * b), "r"(cnt));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int nb = it & 1;
    // "gate math": every thread writes its part of the staging slice
    const uint32_t stg = stage + ((MODE == 3) ? (uint32_t)(it & 1) * slice_bytes : 0u);
    for (int o = threadIdx.x * 16; o < slice_bytes; o += blockDim.x * 16)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stg + o), "r"(it) : "memory");
    if (MODE == 0) {
      __syncthreads();
      // 16-byte chunks: chunk c of the slice goes to all 8 destinations
      const int chunks = slice_bytes / 16;
      for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + c * 16));
        const uint32_t dst = base + nb * buf_bytes + rank * slice_bytes + c * 16;
#pragma unroll
        for (int pc = 0; pc < kCl; ++pc)
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(mapa(dst, pc)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
#pragma unroll
        for (int pc = 0; pc < kCl; ++pc)
          asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa(bar0 + 8 * nb, pc)) : "memory");
      }
    } else {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * nb), "r"(buf_bytes) : "memory");
      const uint32_t dst = base + nb * buf_bytes + rank * slice_bytes;
      if (MODE == 1 || MODE == 3) {
        if (threadIdx.x < kCl)
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mapa(dst, threadIdx.x)), "r"(stg), "r"(slice_bytes), "r"(mapa(bar0 + 8 * nb, threadIdx.x)) : "memory");
      } else if (threadIdx.x == 0) {
#pragma unroll
        for (int pc = 0; pc < kCl; ++pc)
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mapa(dst, pc)), "r"(stg), "r"(slice_bytes), "r"(mapa(bar0 + 8 * nb, pc)) : "memory");
      }
    }
    // every thread waits for the 8 slices of this iteration (the MMA thread would)
    if (!mbar_wait_bounded(bar0 + 8 * nb, (it >> 1) & 1, err)) break;
    __syncthreads();  // staging slice may be rewritten (bulk copies of this CTA have been consumed by all
                      // peers only when THEIR barriers flipped; conservative: sync the cluster every iteration in mode 1/2)
    if (MODE == 1 || MODE == 2) cluster_sync();
  }
  long long t1 = clock64();
  cluster_sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

template <int MODE>
void run(int iters, int slice_bytes) {
  long long* d_out; int* d_err;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
  size_t smem = 2 * kCl * slice_bytes + 2 * slice_bytes + 64;
  cudaFuncSetAttribute(xchg_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  xchg_kernel<MODE><<<kCl, 512, smem>>>(10, slice_bytes, d_out, d_err);  // warm-up
  xchg_kernel<MODE><<<kCl, 512, smem>>>(iters, slice_bytes, d_out, d_err);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0; int err = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
  printf("mode %d slice %5d B (x8 peers = %3d KB out per CTA): %8.1f cycles/iter  [%s%s]\n", MODE, slice_bytes, 8 * slice_bytes / 1024,
         (double)cyc / iters, cudaGetErrorString(e), err ? ", WATCHDOG" : "");
  cudaFree(d_out); cudaFree(d_err);
}

int main() {
  for (int sb : {2048, 4096, 8192}) {
    run<0>(2000, sb);
    run<1>(2000, sb);
    run<2>(2000, sb);
    run<3>(2000, sb);
  }
  return 0;
}
