// Graph executor: runs an ONNX graph on one GPU with the kernels in nn_kernels.cu.
// This is the B200 implementation of the reference's `Model` seam (ocrs/src/model.rs:6-17):
//   input_shape()  <->  Model::input_shape   (model.rs:9, rten impl :20-31)
//   run()          <->  Model::run           (model.rs:12-16, rten impl :33-40)
// The executor is graph-driven: operators, shapes and fusions are derived from the loaded file.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "common.h"
#include "onnx_reader.h"

namespace ocrs {

// Stream-ordered device allocation (cudaMallocAsync); freed on the same stream.
struct Storage {
  void* ptr = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;
  bool owned = true;
  Storage(size_t n, cudaStream_t st);
  Storage(void* p, size_t n) : ptr(p), bytes(n), owned(false) {}
  ~Storage();
  Storage(const Storage&) = delete;
  Storage& operator=(const Storage&) = delete;
};

struct DTensor {
  std::shared_ptr<Storage> storage;
  float* data = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

struct ModelCost {
  double flops = 0;        // 2 * MACs of Conv / ConvTranspose / MatMul / GRU for the last run
  double min_bytes = 0;    // input + output + weights
};

// ONNX operators the executor implements (the 20 of ocrs/src/wasm_api.rs:35-56 plus Squeeze, Identity,
// Constant, Tanh).  Host-side; used by Model::load and by ocrs_b200_model_inspect.
const std::set<std::string>& supported_ops();

class Model {
 public:
  // Parses an ONNX file image and uploads weights to `device`.
  static std::unique_ptr<Model> load(const uint8_t* bytes, size_t len, int device);
  ~Model();

  // Declared shape of the first graph input; -1 = symbolic (rten::Dimension::Symbolic).
  const std::vector<int64_t>& input_shape() const { return input_shape_; }
  int device() const { return device_; }

  // Runs the graph.  `input` lives on the device; the result is stream-ordered on `st`.
  // Re-entrant: any number of threads may call run() on one Model with distinct streams.
  DTensor run(const DTensor& input, cudaStream_t st, ModelCost* cost = nullptr, Profiler* prof = nullptr,
              const std::string& prof_prefix = "", const std::string* stop_at = nullptr) const;

  // ---- packed sequence head (graph-level fusion, optional) ----------------------------------
  // When the graph ends in  X_seq[T,N,C] -> {bidirectional GRU (h0 = 0) -> Transpose(0,2,1,3) ->
  // Reshape(0,0,-1)} x L -> MatMul(const) + Add(const) -> LogSoftmax(last axis), the suffix is
  // row-wise except for the recurrence, so several runs with different (T, N) can share one ragged
  // execution: run_prefix() stops at X_seq, run_seq_head() finishes all groups at once.
  struct PackedGroup {
    int T = 0, N = 0;
    int64_t row_off = 0;  // first row of this group's [T, N, C] block in the packed buffer
  };
  bool has_seq_head() const;
  int seq_head_channels() const;  // C of X_seq
  int seq_head_classes() const;
  DTensor run_prefix(const DTensor& input, cudaStream_t st, ModelCost* cost = nullptr, Profiler* prof = nullptr,
                     const std::string& prof_prefix = "") const;
  // Packed conv prefix: when the graph from its input to X_seq is one tensor-core conv chain (stem conv +
  // 3x3 convs + pooling tail), every layer runs ONCE over all width groups ("ragged" launches, dynamic
  // tile scheduling) and the feature rows are written straight into the packed [rows, C] layout.
  struct PrefixGroup {
    const float* x_base = nullptr;  // common base pointer of all groups (the recognition batch buffer)
    int64_t x_off = 0;              // element offset of this group's [N,1,H,W] f32 block
    int N = 0, W = 0;
  };
  bool has_packed_prefix(int in_h) const;
  // Returns the packed features [rows, C]; `out_groups` receives (T, N, row_off) per group, in order.
  DTensor run_prefix_packed(const std::vector<PrefixGroup>& groups, int in_h, cudaStream_t st,
                            std::vector<PackedGroup>* out_groups, ModelCost* cost = nullptr, Profiler* prof = nullptr,
                            const std::string& prof_prefix = "") const;
  // X: packed [rows, C] f32.  Returns log-probs [rows, classes]; group g occupies rows
  // [row_off, row_off + T*N) in [T, N, classes] order.
  DTensor run_seq_head(const float* X, int64_t rows, const std::vector<PackedGroup>& groups, cudaStream_t st,
                       ModelCost* cost = nullptr, Profiler* prof = nullptr, const std::string& prof_prefix = "") const;

  // The tensor-core conv path carries activations as split fp16; a kernel that meets |x| > 65504
  // raises a device flag.  Protocol (safe with concurrent runs on one Model):
  //     int tok = m.tc_token();  run(...);  synchronise the stream(s);  if (m.take_tc_overflow(tok)) repeat the run;
  // take_tc_overflow returns true when the run started on the tensor-core chains and either the flag is
  // raised (the chains are then switched off for this model: fp32 CUDA-core kernels from then on) or
  // another run has switched them off in the meantime (its check may have consumed this run's flag).
  int tc_token() const;
  bool take_tc_overflow(int token) const;

  size_t weight_bytes() const { return weight_bytes_; }
  const onnx::Graph& graph() const { return graph_; }

 public:
  struct Impl;

 private:
  Model();
  std::unique_ptr<Impl> impl_;
  onnx::Graph graph_;
  std::vector<int64_t> input_shape_;
  int device_ = 0;
  size_t weight_bytes_ = 0;
  // device copies of float initializers (name -> buffer); some re-laid-out per consumer
  std::map<std::string, std::shared_ptr<Storage>> dev_weights_;
  std::map<std::string, std::shared_ptr<Storage>> dev_weights_t_;  // transposed MatMul rhs
  std::map<std::string, int> use_count_;
  std::vector<int> fuse_relu_;     // per node: fuse a following Relu
  std::vector<int> skip_;          // per node: folded into its producer
  std::vector<std::string> fused_bias_;  // per MatMul node: name of bias initializer folded from Add
  std::map<std::string, std::string> alias_;  // output name -> producer output it aliases
};

// Configures the device's default mempool to keep freed blocks cached.
void configure_device_pool(int device);

}  // namespace ocrs
