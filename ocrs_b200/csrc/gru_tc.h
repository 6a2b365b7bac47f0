// Bidirectional GRU of the recognition CRNN on the tensor cores (sm_100a):
//   (1) input projections  xw = X W^T + Wb   : split-bf16 TMA/tcgen05 GEMM over all timesteps
//   (2) recurrence         h_t = f(xw_t, h_{t-1} R^T): one persistent kernel per layer; each CTA owns
//       <= 32 lines of one direction for all T steps, keeps h in registers + shared memory (as the
//       MMA B operand), streams R from L2 with TMA each step and accumulates the 768 gate
//       pre-activations in TMEM (6 tiles of 128 rows x 32 lines).
// ONNX GRU semantics: gate order z, r, h; linear_before_reset = 1 (what PyTorch exports).
// Replaces rten's GRU operator (reached through `Model::run`, ocrs/src/model.rs:33-40).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <functional>
#include <memory>

#include "common.h"

namespace ocrs {
namespace tc {

struct GruWeightsTC {
  int D = 0, H = 0, I = 0;
  DeviceBuffer w_hi, w_lo;  // bf16 [D*3H][I]
  DeviceBuffer wb;          // f32  [D*3H]
  DeviceBuffer r_hi, r_lo;  // bf16 [D*3H][H]
  DeviceBuffer rb;          // f32  [D*3H]
};

bool gru_supported(int D, int H, int I);
std::unique_ptr<GruWeightsTC> prepare_gru(const float* W, const float* R, const float* B, int D, int H, int I);

// Scratch allocator: returns device memory valid until the stream work enqueued by gru_forward
// has been consumed (the executor passes a stream-ordered allocator).
using ScratchAlloc = std::function<void*(size_t bytes)>;

// X: [T,N,I] f32; h0: [D,N,H] f32 or null (zeros); Y: [T,D,N,H] f32; Yh: [D,N,H] f32 or null.
// reverse[d] != 0 -> direction d runs from t = T-1 down to 0.
void gru_forward(const float* X, const GruWeightsTC& w, const float* h0, float* Y, float* Yh, int T, int N,
                 const int* reverse, const ScratchAlloc& alloc, cudaStream_t st);

// Linear layer on the same split-bf16 tensor-core GEMM (the CRNN's final 512 -> classes projection over the
// packed rows of all lines): Y[rows, Npad] = X[rows, K] * W^T + b, Npad = N rounded up to 128 (extra columns
// are zero weights, to be ignored by the consumer).
struct LinearWeightsTC {
  int N = 0, Npad = 0, K = 0;
  DeviceBuffer w_hi, w_lo;  // bf16 [Npad][K]
  DeviceBuffer bias;        // f32 [Npad]
};
bool linear_supported(int N, int K);
// Wt: [N][K] row-major (i.e. the ONNX MatMul weight [K][N] transposed); bias may be null.
std::unique_ptr<LinearWeightsTC> prepare_linear(const float* Wt, const float* bias, int N, int K);
void linear_forward(const float* X, int64_t rows, const LinearWeightsTC& w, float* Y, const ScratchAlloc& alloc,
                    cudaStream_t st);

// One sequence of a ragged batch: own length and own element strides into xw / Y.
struct SeqLine {
  int32_t T;                    // timesteps
  int32_t valid;                // 0 = padding slot
  int64_t xw_base, xw_tstride;  // xw element offset at t = 0 and per-timestep stride (+ d*3H + gate*H + unit)
  int64_t y_base, y_tstride;    // Y  element offset at t = 0 and per-timestep stride (+ d*y_dstride + unit)
};

// Ragged GRU layer over `rows` packed input rows X [rows, I]: writes Y with the strides given per line.
// `lines` must be ordered by T descending (tiles of 32 lines run max-T steps).
// `after_projection` (optional) is called between the launch of the input-projection GEMM and the recurrent kernel
// (the executor's profiler times the two separately).
void gru_forward_lines(const float* X, int64_t rows, const GruWeightsTC& w, const SeqLine* lines_host, int n_lines,
                       float* Y, int64_t y_dstride, const int* reverse, const ScratchAlloc& alloc, cudaStream_t st,
                       const std::function<void()>* after_projection = nullptr);

}  // namespace tc
}  // namespace ocrs
