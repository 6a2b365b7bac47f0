// Dense 3x3 convolutions of the recognition CRNN on the 5th-generation tensor cores:
// implicit GEMM, TMA-fed, tcgen05.mma with the accumulator in TMEM (sm_100a only).
//
// Precision: the reference computes in fp32 and BASELINE.json asks for log-probs within 1e-3 of
// it, so operands are carried as split fp16 (x = hi + lo, 22 significant bits) in NHWC and every
// k-block issues three MMAs (hi*hi + hi*lo + lo*hi) into one fp32 accumulator.  fp16 has a
// narrow range: every kernel that produces a split value sets `*ovf` when |x| > 65504 (or NaN),
// and the executor then repeats the run on the fp32 CUDA-core kernels.
//
// Replaces rten's Conv operator kernels (reached through `Model::run`, ocrs/src/model.rs:33-40)
// for the layers where `group == 1`, kernel 3x3, stride 1, pad 1 and C_in is a multiple of 32.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>

#include "common.h"

namespace ocrs {
namespace tc {

// Element type of the split activation / weight planes: x = hi + lo, both fp16 (22 significant
// bits; bf16 pairs carry only 16 and missed the 1e-3 log-prob bound on real text lines).
using act_t = __half;

// True when the tcgen05 path can be used on this device / build (sm_100 family, driver entry
// point for cuTensorMapEncodeTiled resolvable).
bool available();

struct ConvWeightsTC {
  int Cin = 0, Cout = 0;
  DeviceBuffer w_hi, w_lo;  // fp16 [Cout][9*Cin], k = (kh*3 + kw)*Cin + ci
  DeviceBuffer bias;        // f32 [Cout]
};

// Eligibility of one Conv node (shapes only).
bool conv_supported(int Cin, int Cout, int R, int S, int stride_h, int stride_w, int pad_t, int pad_l, int pad_b,
                    int pad_r, int dil_h, int dil_w, int groups);

// Re-lays out ONNX weights [Cout][Cin][3][3] (+ bias, may be null) for the kernel.
std::unique_ptr<ConvWeightsTC> prepare_weights(const float* w, const float* b, int Cin, int Cout);

// y = maxpool_{ph x pw}(relu?(conv3x3(x) + bias)); x, y: NHWC split fp16.  x: [N,H,W,Cin],
// y: [N,H/ph,W/pw,Cout]; ph, pw in {1, 2} (kernel == stride, no padding, floor).
void conv3x3(const act_t* x_hi, const act_t* x_lo, const ConvWeightsTC& w, act_t* y_hi,
             act_t* y_lo, int N, int H, int W, int relu, int ph, int pw, int* ovf, cudaStream_t st);

// First layer of the CRNN: conv3x3(1 -> Cout, stride 1, pad 1) + bias + ReLU + MaxPool(2,2), fused,
// CUDA cores (K = 9 is not tensor-core work).  x: [N,1,H,W] f32; out: NHWC split fp16 [N,H/2,W/2,Cout].
// w: ONNX layout [Cout][1][3][3].  Cout must be a multiple of 8 and <= 64.
struct StemWeights {
  int Cout = 0;
  DeviceBuffer w, bias;  // f32 [Cout*9], [Cout]
};
bool stem_supported(int Cin, int Cout, int R, int S, int stride_h, int stride_w, int pad_t, int pad_l, int pad_b,
                    int pad_r, int dil_h, int dil_w, int groups);
std::unique_ptr<StemWeights> prepare_stem(const float* w, const float* b, int Cout);
void stem_conv_relu_pool2(const float* x, const StemWeights& w, act_t* y_hi, act_t* y_lo, int N, int H,
                          int W, int* ovf, cudaStream_t st);

// ---- ragged execution over width groups ---------------------------------------------------------
// The recognition batch is a set of width groups (recognition.rs:431-459), each its own NHWC tensor.
// One launch per layer covers all of them: every group has one RaggedDesc per layer (device array,
// ordered by `first`) and, for the tensor-core layers, a pair of TMA tensor maps (hi, lo planes).
struct RaggedDesc {
  int32_t N, H, W;           // input dims of this layer for the group
  int32_t OH, OW;            // output dims (after the fused pool)
  int32_t tiles_w, tiles_h;  // conv: 8x16-pixel tiles per image
  int32_t first;             // first tile (conv) / first thread block (stem, tail) of the group in the launch
  int64_t in_off;            // stem: first f32 element of the group's [N,1,H,W] input; tail: first input pixel
  int64_t out_off;           // first output pixel (NHWC) of the group; tail: first row of the packed [rows, C] output
};
static_assert(sizeof(RaggedDesc) == 48, "RaggedDesc layout");

int conv_tiles(int N, int H, int W);      // tiles of one group in a conv launch
void conv_fill_tiles(RaggedDesc* d);      // sets tiles_w / tiles_h from H / W
// TMA tensor maps {hi, lo} of a group's NHWC input [N,H,W,Cin] (host side; copy them to the device array).
// `Cout` of the consuming layer selects the box shape (the CTA-pair + halo kernel of the 128-channel layers,
// OCRS_B200_CONV_MODE=2, loads 10-row boxes).
void make_act_maps(const act_t* hi, const act_t* lo, int N, int H, int W, int Cin, CUtensorMap out[2], int Cout = 0);
// d_maps: device [2 * n_groups]; d_counter: device int, zero before the launch (tile scheduler).
void conv3x3_ragged(const CUtensorMap* d_maps, const RaggedDesc* d_groups, int n_groups, int n_tiles, int* d_counter,
                    const ConvWeightsTC& w, act_t* y_hi, act_t* y_lo, int relu, int ph, int pw, int* ovf, cudaStream_t st);
// blocks of 128 threads, one thread per pooled output pixel: n_blocks = sum over groups of ceil(N*OH*OW / 128)
void stem_ragged(const float* x, const StemWeights& w, act_t* y_hi, act_t* y_lo, const RaggedDesc* d_groups, int n_groups,
                 int n_blocks, int* ovf, cudaStream_t st);
// blocks of 256 threads, one thread per (n, w, 8 channels): n_blocks = sum over groups of ceil(N*W*C/8 / 256)
void avg_to_seq_ragged(const act_t* hi, const act_t* lo, float* y, int C, const RaggedDesc* d_groups, int n_groups,
                       int n_blocks, cudaStream_t st);

// Layout / precision converters and the pooling used between tensor-core layers.
void nchw_to_nhwc_split(const float* x, act_t* hi, act_t* lo, int N, int C, int H, int W, int* ovf,
                        cudaStream_t st);
void nhwc_split_to_nchw(const act_t* hi, const act_t* lo, float* y, int N, int C, int H, int W,
                        cudaStream_t st);
// Tail of the CRNN's conv stack fused into one pass: AveragePool over the full height (H -> 1),
// squeeze, and the [N,C,W] -> [W,N,C] transpose.  x: NHWC split fp16 [N,H,W,C]; y: f32 [W,N,C].
void nhwc_split_avg_to_seq(const act_t* hi, const act_t* lo, float* y, int N, int C, int H, int W, cudaStream_t st);
// max-pool with kernel == stride == (ph, pw), no padding; NHWC split in and out.
void maxpool_nhwc_split(const act_t* x_hi, const act_t* x_lo, act_t* y_hi,
                        act_t* y_lo, int N, int H, int W, int C, int ph, int pw, cudaStream_t st);

}  // namespace tc
}  // namespace ocrs
