// fp32 operator kernels for the two networks (NCHW, row-major), launched on a caller stream.
// These implement the ONNX operators the reference registers for its models
// (ocrs/src/wasm_api.rs:35-56) and stand where rten's CPU operator kernels stood
// (`rten::Model::run_one`, reached from ocrs/src/model.rs:33-40).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace ocrs {
namespace nn {

struct ConvParams {
  int N, C, H, W;      // input
  int K, R, S;         // output channels, kernel h/w
  int stride_h, stride_w, pad_t, pad_l, dil_h, dil_w, groups;
  int OH, OW;
  int relu;            // fused ReLU epilogue
};

// y = conv(x, w) + b.  w: [K, C/groups, R, S].  b may be null.
void conv2d(const float* x, const float* w, const float* b, float* y, const ConvParams& p, cudaStream_t st);

struct ConvTParams {
  int N, C, H, W;   // input
  int K, R, S;      // output channels (per group * groups), kernel
  int stride_h, stride_w, pad_t, pad_l, groups;
  int OH, OW;
  int relu;
};
// w: [C, K/groups, R, S]
void conv_transpose2d(const float* x, const float* w, const float* b, float* y, const ConvTParams& p, cudaStream_t st);

// Fused depthwise 3x3 (pad 1, stride s) + pointwise 1x1 (+ ReLU): y = relu?(pw(dw(x))).
// x [N,C,H,W]; dw_w [C][9], dw_b [C]; pw_w [K][C], pw_b [K]; y [N,K,OH,OW].
bool dwpw_supported(int C, int K);
// Same with a VIRTUAL input: the channel concatenation of up to two tensors, each optionally end-padded
// (bottom / right) with a constant up to Hv x Wv -- the U-Net decoder's Pad + Concat are never written.
struct SepInput {
  const float* x;  // [N, C, H, W]
  int C, H, W;
  float pad;       // value of the end padding (H..Hv, W..Wv)
};
void dwpw_conv2(const SepInput* srcs, int n_src, const float* dw_w, const float* dw_b, const float* pw_w, const float* pw_b,
                float* y, int N, int Hv, int Wv, int K, int stride, int relu, cudaStream_t st);
void dwpw_conv(const float* x, const float* dw_w, const float* dw_b, const float* pw_w, const float* pw_b, float* y,
               int N, int C, int H, int W, int K, int stride, int relu, cudaStream_t st);

// ConvTranspose 2x2 stride 2 (no padding), any C -> K, optional ReLU.  w: [C][K][2][2].
void conv_transpose_2x2s2(const float* x, const float* w, const float* b, float* y, int N, int C, int H, int W, int K,
                          int relu, cudaStream_t st);

// ConvTranspose 2x2 s2 (C -> Km) + ReLU, then Conv 1x1 (Km -> 1) + Sigmoid, fused; C, Km <= 16.
// y: [N,1,2H,2W].
void conv_transpose_2x2s2_head(const float* x, const float* w, const float* b, const float* w2, const float* b2,
                               float* y, int N, int C, int H, int W, int Km, cudaStream_t st);

struct PoolParams {
  int NC, H, W, R, S, stride_h, stride_w, pad_t, pad_l, OH, OW;
  int count_include_pad;
};
void max_pool2d(const float* x, float* y, const PoolParams& p, cudaStream_t st);
void avg_pool2d(const float* x, float* y, const PoolParams& p, cudaStream_t st);

void relu(const float* x, float* y, int64_t n, cudaStream_t st);
void sigmoid(const float* x, float* y, int64_t n, cudaStream_t st);
void tanh_op(const float* x, float* y, int64_t n, cudaStream_t st);
// y[i] = a[i] + b[i % nb]  (trailing-dims broadcast: b's shape is a suffix of a's)
void add_bcast_suffix(const float* a, const float* b, float* y, int64_t n, int64_t nb, cudaStream_t st);

// C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ relu)
void sgemm_nt(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int relu,
              cudaStream_t st);

// Generic permute of a <=6-D tensor: y = x.permute(perm).contiguous()
void permute(const float* x, float* y, const int64_t* shape, const int* perm, int ndim, cudaStream_t st);

// Copies a [outer, len_src, inner] block into y at offset `dst_off` along the concat axis.
void concat_copy(const float* x, float* y, int64_t outer, int64_t len_src, int64_t len_dst, int64_t dst_off,
                 int64_t inner, cudaStream_t st);

// Constant pad / crop on a <=4-D tensor (begin/end may be negative = crop).
void pad4d(const float* x, float* y, const int64_t in_shape[4], const int64_t begin[4], const int64_t out_shape[4],
           float value, cudaStream_t st);

void fill(float* y, float v, int64_t n, cudaStream_t st);

// log-softmax over the last axis
void log_softmax_lastdim(const float* x, float* y, int64_t rows, int cols, cudaStream_t st);
// same with a row stride `ldx` (elements) on the input; y stays dense [rows, cols]
void log_softmax_rows(const float* x, int64_t ldx, float* y, int64_t rows, int cols, cudaStream_t st);

// One GRU time step for both directions (ONNX gate order z, r, h).
//   xw   : [D][T][N][3H] input projections incl. Wb
//   R    : [D][3H][H], Rb: [D][3H]
//   h_in/h_out : [D][N][H]
//   Y    : [T][D][N][H]
// direction d processes time index t_d = (d == reverse) ? T-1-step : step.
void gru_step(const float* xw, const float* R, const float* Rb, const float* h_in, float* h_out, float* Y, int D,
              int T, int N, int H, int step, const int* dir_reverse_host, int linear_before_reset,
              cudaStream_t st);

}  // namespace nn
}  // namespace ocrs
