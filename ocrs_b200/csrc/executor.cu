// ONNX graph executor (host orchestration; kernels live in nn_kernels.cu).
#include "executor.h"
#include <atomic>
#include <mutex>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <set>

#include "conv_tc.h"
#include "gru_tc.h"
#include "nn_kernels.h"

namespace ocrs {

std::atomic<int64_t> g_kernel_launches{0};

Profiler::~Profiler() {
  for (auto& p : pending_) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : free_events_) cudaEventDestroy(e);
}
cudaEvent_t Profiler::get_event() {
  if (!free_events_.empty()) { cudaEvent_t e = free_events_.back(); free_events_.pop_back(); return e; }
  cudaEvent_t e;
  OCRS_CUDA_CHECK(cudaEventCreate(&e));
  return e;
}
int Profiler::begin(const std::string& name, cudaStream_t st) {
  if (!enabled) return -1;
  Pending p{name, get_event(), get_event(), 0, 0, g_kernel_launches.load(), 0};
  OCRS_CUDA_CHECK(cudaEventRecord(p.a, st));
  pending_.push_back(std::move(p));
  return (int)pending_.size() - 1;
}
void Profiler::end(int token, cudaStream_t st, double flops, double bytes) {
  if (token < 0) return;
  Pending& p = pending_[(size_t)token];
  OCRS_CUDA_CHECK(cudaEventRecord(p.b, st));
  p.flops = flops;
  p.bytes = bytes;
  p.launches = g_kernel_launches.load() - p.launches0;
}
void Profiler::collect() {
  for (auto& p : pending_) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      OpProfile& o = acc_[p.name];
      o.ms += ms;
      o.flops += p.flops;
      o.bytes += p.bytes;
      o.calls += 1;
      o.launches += p.launches;
    }
    free_events_.push_back(p.a);
    free_events_.push_back(p.b);
  }
  pending_.clear();
}

using onnx::Attr;
using onnx::Node;
using onnx::TensorData;

Storage::Storage(size_t n, cudaStream_t st) : bytes(n), stream(st) {
  if (n == 0) n = 4;
  OCRS_CUDA_CHECK(cudaMallocAsync(&ptr, n, st));
}
Storage::~Storage() {
  if (ptr && owned) {
    if (cudaFreeAsync(ptr, stream) != cudaSuccess) {
      cudaGetLastError();  // do not leave a sticky error behind for unrelated launches
      cudaFree(ptr);
      cudaGetLastError();
    }
  }
}

const std::set<std::string>& supported_ops() {
  static const std::set<std::string> kSupported = {
      "Add", "AveragePool", "Cast", "Concat", "ConstantOfShape", "Conv", "ConvTranspose", "GRU", "Gather",
      "LogSoftmax", "MatMul", "MaxPool", "Pad", "Relu", "Reshape", "Shape", "Sigmoid", "Slice", "Transpose",
      "Unsqueeze", "Squeeze", "Identity", "Constant", "Tanh"};
  return kSupported;
}

void configure_device_pool(int device) {
  cudaMemPool_t pool;
  OCRS_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  OCRS_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  // Pre-size the pool once per device: activations are allocated stream-ordered on several
  // streams, and a pool that has to grow from the OS in the middle of a batch stalls every
  // stream of the device.  One allocate + free leaves the memory cached in the pool.
  static std::mutex mu;
  static std::set<int> warmed;
  std::lock_guard<std::mutex> lk(mu);
  if (warmed.insert(device).second) {
    size_t mb = 12288;
    if (const char* e = std::getenv("OCRS_B200_POOL_PREWARM_MB")) mb = (size_t)std::strtoull(e, nullptr, 10);
    size_t free_b = 0, total_b = 0;
    OCRS_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    size_t bytes = std::min(mb << 20, free_b / 4);
    if (bytes > 0) {
      void* p = nullptr;
      if (cudaMallocAsync(&p, bytes, (cudaStream_t)0) == cudaSuccess) {
        cudaFreeAsync(p, (cudaStream_t)0);
        cudaStreamSynchronize((cudaStream_t)0);
      } else {
        cudaGetLastError();  // not fatal: the pool simply grows on demand
      }
    }
  }
}

namespace {

struct Value {
  bool is_int = false;
  std::vector<int64_t> shape;   // for ints: shape of the int tensor
  std::vector<int64_t> ivals;   // host ints
  DTensor t;                    // device floats (shape mirrors `shape`)
  const TensorData* host_f32 = nullptr;  // host copy for float initializers (pad value etc.)
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

DTensor alloc_tensor(const std::vector<int64_t>& shape, cudaStream_t st) {
  DTensor t;
  t.shape = shape;
  t.storage = std::make_shared<Storage>((size_t)t.numel() * sizeof(float), st);
  t.data = reinterpret_cast<float*>(t.storage->ptr);
  return t;
}

Value dev_value(DTensor t) {
  Value v;
  v.shape = t.shape;
  v.t = std::move(t);
  return v;
}
Value int_value(std::vector<int64_t> vals, std::vector<int64_t> shape) {
  Value v;
  v.is_int = true;
  v.ivals = std::move(vals);
  v.shape = std::move(shape);
  return v;
}

int64_t norm_axis(int64_t a, int64_t nd) {
  if (a < 0) a += nd;
  OCRS_CHECK(a >= 0 && a < nd, kRunFailed, "axis out of range");
  return a;
}

std::shared_ptr<Storage> upload(const void* host, size_t bytes) {
  auto s = std::make_shared<Storage>(bytes, (cudaStream_t) nullptr);
  OCRS_CUDA_CHECK(cudaMemcpy(s->ptr, host, bytes, cudaMemcpyHostToDevice));
  return s;
}

}  // namespace

struct TcUnit {
  int conv_node = -1, pool_node = -1;
  int ph = 1, pw = 1, relu = 0;
  std::unique_ptr<tc::ConvWeightsTC> w;
};
struct TcChain {
  std::unique_ptr<tc::StemWeights> stem;  // optional fused first layer: Conv(1->C)+ReLU+MaxPool(2,2)
  std::vector<TcUnit> units;
  std::string out_name;  // published name of the chain's result
  // optional fused tail: AveragePool((kh,1)) -> Reshape([0,C,-1]) -> Transpose(2,0,1), applied when
  // the chain's output height equals kh at run time
  struct Tail { int avg_node = -1, reshape_node = -1, transpose_node = -1, kh = 0; std::string out_name; } tail;
};

struct Model::Impl {
  // per GRU node: Wb [D][3H], Rb [D][3H]
  std::map<int, std::pair<std::shared_ptr<Storage>, std::shared_ptr<Storage>>> gru_bias;
  std::vector<std::string> out_rename;  // per node: published name of output 0 ("" = own name)
  // tensor-core conv chains (Conv3x3 [+Relu] [+MaxPool]) keyed by their first node
  std::map<int, TcChain> tc_chains;
  std::map<int, std::unique_ptr<tc::GruWeightsTC>> tc_gru;  // per GRU node
  std::unique_ptr<tc::LinearWeightsTC> head_fc;             // the packed head's Linear on the tensor cores
  // packed sequence head
  struct SeqHead {
    bool present = false;
    std::string x_name;          // X_seq value
    std::vector<int> gru_nodes;  // in execution order
    int fc_node = -1;            // MatMul node (bias fused)
    int C = 0, H = 0, classes = 0;
  } head;
  std::vector<int> tc_member;  // node is executed as part of a chain started earlier
  // detection-net fusions (fp32 CUDA cores)
  std::vector<int> dwpw_partner;     // per depthwise Conv node: index of the fused pointwise Conv (-1 = none)
  // virtual input of a fused depthwise + pointwise node: the Concat (and the Pads feeding it) in front of
  // it are absorbed -- the kernel reads the sources directly
  struct SepPlan {
    std::vector<std::string> src;   // source value names, in channel order
    std::vector<int> pad_b, pad_r;  // end padding of each source (rows, columns)
    std::vector<float> pad_v;
  };
  std::map<int, SepPlan> sep_plan;
  struct HeadFuse { int conv1x1 = -1, sigmoid = -1; };
  std::map<int, HeadFuse> head_fuse; // per ConvTranspose node: fused 1x1 conv + sigmoid
  bool tc_enabled = false;
  std::atomic<bool> tc_conv_on{true};  // cleared for good after an fp16 range overflow
  std::mutex ovf_mu;                    // serialises the flag read / switch-off in take_tc_overflow
  DeviceBuffer tc_ovf;                  // int32 flag written by the split-fp16 kernels
};

Model::Model() = default;
Model::~Model() = default;

std::unique_ptr<Model> Model::load(const uint8_t* bytes, size_t len, int device) {
  std::unique_ptr<Model> m(new Model());
  m->graph_ = onnx::parse_model(bytes, len);
  m->device_ = device;
  OCRS_CUDA_CHECK(cudaSetDevice(device));
  configure_device_pool(device);
  const auto& g = m->graph_;
  OCRS_CHECK(!g.inputs.empty(), kModelLoad, "model has no inputs");  // model.rs:24
  OCRS_CHECK(!g.outputs.empty(), kModelLoad, "model has no outputs");
  OCRS_CHECK(!g.inputs[0].dims.empty(), kModelLoad, "model does not specify expected input shape");  // model.rs:28
  m->input_shape_ = g.inputs[0].dims;

  for (const auto& n : g.nodes)
    OCRS_CHECK(supported_ops().count(n.op), kModelLoad, "unsupported ONNX operator: " + n.op);

  auto impl = std::make_unique<Impl>();
  const int nn_ = (int)g.nodes.size();
  m->fuse_relu_.assign(nn_, 0);
  m->skip_.assign(nn_, 0);
  m->fused_bias_.assign(nn_, "");
  impl->out_rename.assign(nn_, "");

  // consumer counts over the raw graph
  std::map<std::string, int> consumers;
  std::map<std::string, int> consumer_node;  // last consumer node index (valid when count == 1)
  for (int i = 0; i < nn_; ++i)
    for (const auto& in : g.nodes[i].inputs)
      if (!in.empty()) { consumers[in]++; consumer_node[in] = i; }
  std::set<std::string> graph_outs;
  for (const auto& o : g.outputs) graph_outs.insert(o.name);

  // fusions: Conv/ConvTranspose -> Relu ; MatMul -> Add(const bias)
  for (int i = 0; i < nn_; ++i) {
    const Node& n = g.nodes[i];
    if (n.outputs.empty()) continue;
    const std::string& out = n.outputs[0];
    if (consumers[out] != 1 || graph_outs.count(out)) continue;
    int j = consumer_node[out];
    const Node& c = g.nodes[j];
    if ((n.op == "Conv" || n.op == "ConvTranspose") && c.op == "Relu") {
      m->fuse_relu_[i] = 1;
      m->skip_[j] = 1;
      impl->out_rename[i] = c.outputs[0];
    } else if (n.op == "MatMul" && c.op == "Add" && c.inputs.size() == 2) {
      const std::string& other = c.inputs[0] == out ? c.inputs[1] : c.inputs[0];
      auto it = g.initializers.find(other);
      auto wit = n.inputs.size() == 2 ? g.initializers.find(n.inputs[1]) : g.initializers.end();
      if (it != g.initializers.end() && it->second.dtype == onnx::kFloat && it->second.dims.size() == 1 &&
          wit != g.initializers.end() && wit->second.dims.size() == 2 &&
          wit->second.dims[1] == it->second.dims[0]) {
        m->fused_bias_[i] = other;
        m->skip_[j] = 1;
        impl->out_rename[i] = c.outputs[0];
      }
    }
  }

  // ---- depthwise 3x3 + pointwise 1x1 fusion; ConvTranspose(2x2,s2)+ReLU+Conv1x1+Sigmoid fusion ----
  impl->dwpw_partner.assign(nn_, -1);
  if (std::getenv("OCRS_B200_DISABLE_DET_FUSION") == nullptr) {
    auto sole = [&](const std::string& name) -> int {
      if (graph_outs.count(name)) return -1;
      int found = -1, cnt = 0;
      for (int j = 0; j < nn_; ++j) {
        if (m->skip_[j]) continue;
        for (const auto& in : g.nodes[j].inputs)
          if (in == name) { ++cnt; found = j; }
      }
      return cnt == 1 ? found : -1;
    };
    auto winit = [&](const Node& n, size_t idx) -> const TensorData* {
      if (n.inputs.size() <= idx || n.inputs[idx].empty()) return nullptr;
      auto it = g.initializers.find(n.inputs[idx]);
      return it == g.initializers.end() ? nullptr : &it->second;
    };
    for (int i = 0; i < nn_; ++i) {
      const Node& n = g.nodes[i];
      if (m->skip_[i]) continue;
      if (n.op == "Conv" && !m->fuse_relu_[i]) {
        const TensorData* w = winit(n, 1);
        if (!w || w->dims.size() != 4) continue;
        int grp = (int)n.attr_i("group", 1);
        auto pads = n.attr_ints("pads", {0, 0, 0, 0});
        auto st = n.attr_ints("strides", {1, 1});
        auto dl = n.attr_ints("dilations", {1, 1});
        bool is_dw = grp == w->dims[0] && w->dims[1] == 1 && w->dims[2] == 3 && w->dims[3] == 3 && pads == std::vector<int64_t>{1, 1, 1, 1} &&
                     st[0] == st[1] && (st[0] == 1 || st[0] == 2) && dl == std::vector<int64_t>{1, 1} &&
                     (n.inputs.size() < 3 || n.inputs[2].empty() || winit(n, 2));
        if (!is_dw) continue;
        int j = sole(n.outputs[0]);
        if (j < 0 || g.nodes[j].op != "Conv" || g.nodes[j].inputs[0] != n.outputs[0]) continue;
        const Node& pn = g.nodes[j];
        const TensorData* pw = winit(pn, 1);
        if (!pw || pw->dims.size() != 4 || pw->dims[2] != 1 || pw->dims[3] != 1 || pn.attr_i("group", 1) != 1) continue;
        if (pn.attr_ints("pads", {0, 0, 0, 0}) != std::vector<int64_t>{0, 0, 0, 0} ||
            pn.attr_ints("strides", {1, 1}) != std::vector<int64_t>{1, 1}) continue;
        if (pn.inputs.size() > 2 && !pn.inputs[2].empty() && !winit(pn, 2)) continue;
        if (pw->dims[1] != w->dims[0] || !nn::dwpw_supported((int)w->dims[0], (int)pw->dims[0])) continue;
        impl->dwpw_partner[i] = j;
        m->skip_[j] = 1;
        impl->out_rename[i] = impl->out_rename[j].empty() ? pn.outputs[0] : impl->out_rename[j];
        // ---- virtual input: Concat(axis 1) of tensors that may each be end-padded by a constant Pad ----
        if (st[0] == 1) {
          auto producer_of = [&](const std::string& name) -> int {
            for (int q = 0; q < i; ++q)
              if (!m->skip_[q] && !g.nodes[q].outputs.empty() && g.nodes[q].outputs[0] == name && impl->out_rename[q].empty()) return q;
            return -1;
          };
          const int cn = producer_of(n.inputs[0]);
          if (cn >= 0 && g.nodes[cn].op == "Concat" && g.nodes[cn].attr_i("axis", 0) == 1 && g.nodes[cn].inputs.size() >= 1 &&
              g.nodes[cn].inputs.size() <= 2 && sole(n.inputs[0]) == i) {
            Impl::SepPlan plan;
            bool ok = true;
            std::vector<int> pad_nodes;
            for (const auto& in_name : g.nodes[cn].inputs) {
              int pb = 0, pr = 0;
              float pv = 0.f;
              std::string src = in_name;
              const int pd = producer_of(in_name);
              if (pd >= 0 && g.nodes[pd].op == "Pad" && sole(in_name) == cn && g.nodes[pd].attr_s("mode", "constant") == "constant" &&
                  g.nodes[pd].inputs.size() >= 2 && g.nodes[pd].inputs.size() <= 3) {
                const TensorData* pt = winit(g.nodes[pd], 1);
                const TensorData* vt = g.nodes[pd].inputs.size() > 2 && !g.nodes[pd].inputs[2].empty() ? winit(g.nodes[pd], 2) : nullptr;
                bool pad_ok = pt != nullptr && (g.nodes[pd].inputs.size() < 3 || g.nodes[pd].inputs[2].empty() || vt != nullptr);
                std::vector<int64_t> pv64;
                if (pad_ok) {
                  pv64 = pt->as_int64();
                  pad_ok = pv64.size() == 8 && pv64[0] == 0 && pv64[1] == 0 && pv64[2] == 0 && pv64[3] == 0 && pv64[4] == 0 &&
                           pv64[5] == 0 && pv64[6] > -4096 && pv64[7] > -4096 && pv64[6] < 4096 && pv64[7] < 4096;  // negative = crop
                }
                if (pad_ok && vt != nullptr) pad_ok = vt->dtype == onnx::kFloat && vt->numel() == 1;
                if (pad_ok) {
                  pb = (int)pv64[6];
                  pr = (int)pv64[7];
                  pv = vt ? vt->f32()[0] : 0.f;
                  src = g.nodes[pd].inputs[0];
                  pad_nodes.push_back(pd);
                }
              }
              if (g.initializers.count(src)) ok = false;  // constant sources stay on the generic path
              plan.src.push_back(src);
              plan.pad_b.push_back(pb);
              plan.pad_r.push_back(pr);
              plan.pad_v.push_back(pv);
            }
            if (ok) {
              m->skip_[cn] = 1;
              for (int pd : pad_nodes) m->skip_[pd] = 1;
              impl->sep_plan[i] = std::move(plan);
            }
          }
        }
      } else if (n.op == "ConvTranspose" && m->fuse_relu_[i]) {
        const TensorData* w = winit(n, 1);
        if (!w || w->dims.size() != 4 || w->dims[2] != 2 || w->dims[3] != 2 || n.attr_i("group", 1) != 1) continue;
        if (n.attr_ints("strides", {1, 1}) != std::vector<int64_t>{2, 2} ||
            n.attr_ints("pads", {0, 0, 0, 0}) != std::vector<int64_t>{0, 0, 0, 0}) continue;
        if (w->dims[0] > 16 || w->dims[1] > 16) continue;
        int j = sole(impl->out_rename[i]);
        if (j < 0 || g.nodes[j].op != "Conv" || m->fuse_relu_[j]) continue;
        const Node& cn = g.nodes[j];
        const TensorData* w2 = winit(cn, 1);
        if (!w2 || w2->dims.size() != 4 || w2->dims[0] != 1 || w2->dims[1] != w->dims[1] || w2->dims[2] != 1 || w2->dims[3] != 1) continue;
        if (cn.attr_i("group", 1) != 1 || cn.attr_ints("pads", {0, 0, 0, 0}) != std::vector<int64_t>{0, 0, 0, 0} ||
            cn.attr_ints("strides", {1, 1}) != std::vector<int64_t>{1, 1}) continue;
        if (cn.inputs.size() > 2 && !cn.inputs[2].empty() && !winit(cn, 2)) continue;
        int k = sole(cn.outputs[0]);
        if (k < 0 || g.nodes[k].op != "Sigmoid") {
          // the Sigmoid may produce the graph output: `sole` rejects graph outputs only for the *input* name
          continue;
        }
        Impl::HeadFuse hf;
        hf.conv1x1 = j;
        hf.sigmoid = k;
        impl->head_fuse[i] = hf;
        m->skip_[j] = 1;
        m->skip_[k] = 1;
        impl->out_rename[i] = g.nodes[k].outputs[0];
      }
    }
  }

  // remaining-use counts for the fused schedule
  for (int i = 0; i < nn_; ++i) {
    if (m->skip_[i]) continue;
    auto sp = impl->sep_plan.find(i);
    if (sp != impl->sep_plan.end()) {
      for (const auto& in : sp->second.src) m->use_count_[in]++;
      continue;  // the remaining inputs are initializers
    }
    for (const auto& in : g.nodes[i].inputs)
      if (!in.empty()) m->use_count_[in]++;
  }
  for (const auto& o : g.outputs) m->use_count_[o.name] += 1 << 20;

  // ---- tensor-core chains: Conv(3x3,s1,p1,g1) [+Relu fused] [+MaxPool k==s] ... ----
  impl->tc_member.assign(nn_, 0);
  impl->tc_enabled = tc::available() && std::getenv("OCRS_B200_DISABLE_TC") == nullptr;
  if (impl->tc_enabled) {
    impl->tc_ovf.reserve(4);
    OCRS_CUDA_CHECK(cudaMemset(impl->tc_ovf.ptr, 0, 4));
    auto published = [&](int i) { return impl->out_rename[i].empty() ? g.nodes[i].outputs[0] : impl->out_rename[i]; };
    auto conv_ok = [&](const Node& n) {
      if (n.op != "Conv" || n.inputs.size() < 2) return false;
      auto wit = g.initializers.find(n.inputs[1]);
      if (wit == g.initializers.end() || wit->second.dims.size() != 4) return false;
      if (n.inputs.size() > 2 && !n.inputs[2].empty() && !g.initializers.count(n.inputs[2])) return false;
      auto pads = n.attr_ints("pads", {0, 0, 0, 0});
      auto st = n.attr_ints("strides", {1, 1});
      auto dl = n.attr_ints("dilations", {1, 1});
      int grp = (int)n.attr_i("group", 1);
      const auto& d = wit->second.dims;
      return n.attr_s("auto_pad", "NOTSET") == "NOTSET" &&
             tc::conv_supported((int)d[1] * grp, (int)d[0], (int)d[2], (int)d[3], (int)st[0], (int)st[1], (int)pads[0],
                                (int)pads[1], (int)pads[2], (int)pads[3], (int)dl[0], (int)dl[1], grp);
    };
    auto sole_consumer = [&](const std::string& name) -> int {
      if (graph_outs.count(name)) return -1;
      int found = -1, cnt = 0;
      for (int j = 0; j < nn_; ++j) {
        if (m->skip_[j]) continue;
        for (const auto& in : g.nodes[j].inputs)
          if (in == name) { ++cnt; found = j; }
      }
      return cnt == 1 ? found : -1;
    };
    auto stem_ok = [&](const Node& n) {
      if (n.op != "Conv" || n.inputs.size() < 2) return false;
      auto wit = g.initializers.find(n.inputs[1]);
      if (wit == g.initializers.end() || wit->second.dims.size() != 4) return false;
      if (n.inputs.size() > 2 && !n.inputs[2].empty() && !g.initializers.count(n.inputs[2])) return false;
      auto pads = n.attr_ints("pads", {0, 0, 0, 0});
      auto st = n.attr_ints("strides", {1, 1});
      auto dl = n.attr_ints("dilations", {1, 1});
      int grp = (int)n.attr_i("group", 1);
      const auto& d = wit->second.dims;
      return n.attr_s("auto_pad", "NOTSET") == "NOTSET" &&
             tc::stem_supported((int)d[1] * grp, (int)d[0], (int)d[2], (int)d[3], (int)st[0], (int)st[1], (int)pads[0],
                                (int)pads[1], (int)pads[2], (int)pads[3], (int)dl[0], (int)dl[1], grp);
    };
    for (int i = 0; i < nn_; ++i) {
      if (m->skip_[i] || impl->tc_member[i]) continue;
      TcChain chain;
      int cur = i;
      std::vector<int> stem_nodes;
      if (stem_ok(g.nodes[i]) && m->fuse_relu_[i]) {
        // Conv(1->C) + ReLU + MaxPool(2,2) feeding a tensor-core conv
        int pn = sole_consumer(published(i));
        if (pn >= 0 && g.nodes[pn].op == "MaxPool" && g.nodes[pn].attr_ints("kernel_shape", {}) == std::vector<int64_t>{2, 2} &&
            g.nodes[pn].attr_ints("strides", {1, 1}) == std::vector<int64_t>{2, 2} &&
            g.nodes[pn].attr_ints("pads", {0, 0, 0, 0}) == std::vector<int64_t>{0, 0, 0, 0} && g.nodes[pn].attr_i("ceil_mode", 0) == 0) {
          int nx = sole_consumer(g.nodes[pn].outputs[0]);
          if (nx >= 0 && !m->skip_[nx] && conv_ok(g.nodes[nx]) && g.nodes[nx].inputs[0] == g.nodes[pn].outputs[0] &&
              g.initializers.at(g.nodes[nx].inputs[1]).dims[1] == g.initializers.at(g.nodes[i].inputs[1]).dims[0]) {
            const auto& wt = g.initializers.at(g.nodes[i].inputs[1]);
            const float* bptr = (g.nodes[i].inputs.size() > 2 && !g.nodes[i].inputs[2].empty()) ? g.initializers.at(g.nodes[i].inputs[2]).f32() : nullptr;
            chain.stem = tc::prepare_stem(wt.f32(), bptr, (int)wt.dims[0]);
            stem_nodes = {pn, nx};
            cur = nx;
          }
        }
      }
      if (!chain.stem && !conv_ok(g.nodes[i])) continue;
      while (true) {
        const Node& cn = g.nodes[cur];
        TcUnit u;
        u.conv_node = cur;
        u.relu = m->fuse_relu_[cur];
        const auto& wt = g.initializers.at(cn.inputs[1]);
        const float* bptr = (cn.inputs.size() > 2 && !cn.inputs[2].empty()) ? g.initializers.at(cn.inputs[2]).f32() : nullptr;
        u.w = tc::prepare_weights(wt.f32(), bptr, (int)wt.dims[1], (int)wt.dims[0]);
        std::string out = published(cur);
        int nxt = sole_consumer(out);
        if (nxt >= 0 && g.nodes[nxt].op == "MaxPool" && g.nodes[nxt].inputs[0] == out) {
          const Node& pn = g.nodes[nxt];
          auto ks = pn.attr_ints("kernel_shape", {});
          auto ps = pn.attr_ints("pads", {0, 0, 0, 0});
          auto ss = pn.attr_ints("strides", {1, 1});
          bool ok = ks.size() == 2 && ss.size() == 2 && ks[0] == ss[0] && ks[1] == ss[1] && pn.attr_i("ceil_mode", 0) == 0 &&
                    ps[0] == 0 && ps[1] == 0 && ps[2] == 0 && ps[3] == 0 && pn.outputs.size() == 1;
          if (ok) {
            u.pool_node = nxt;
            u.ph = (int)ks[0];
            u.pw = (int)ks[1];
            out = pn.outputs[0];
            nxt = sole_consumer(out);
          }
        }
        chain.units.push_back(std::move(u));
        chain.out_name = out;
        if (nxt >= 0 && !m->skip_[nxt] && conv_ok(g.nodes[nxt]) && g.nodes[nxt].inputs[0] == out &&
            g.initializers.at(g.nodes[nxt].inputs[1]).dims[1] == g.initializers.at(g.nodes[chain.units.back().conv_node].inputs[1]).dims[0]) {
          cur = nxt;
          continue;
        }
        break;
      }
      {  // fused tail (see TcChain::Tail)
        const int an = sole_consumer(chain.out_name);
        if (an >= 0 && !m->skip_[an] && g.nodes[an].op == "AveragePool" && g.nodes[an].inputs[0] == chain.out_name &&
            g.nodes[an].attr_s("auto_pad", "NOTSET") == "NOTSET" && g.nodes[an].attr_i("ceil_mode", 0) == 0) {
          const Node& a = g.nodes[an];
          auto ks = a.attr_ints("kernel_shape", {});
          auto ps = a.attr_ints("pads", {0, 0, 0, 0});
          auto ss = a.attr_ints("strides", {1, 1});
          const bool a_ok = ks.size() == 2 && ks[1] == 1 && ks[0] >= 1 && ss.size() == 2 && ss[1] == 1 &&
                            ps == std::vector<int64_t>{0, 0, 0, 0};
          const int rn = a_ok ? sole_consumer(a.outputs[0]) : -1;
          if (rn >= 0 && !m->skip_[rn] && g.nodes[rn].op == "Reshape" && g.nodes[rn].inputs.size() == 2 &&
              g.nodes[rn].inputs[0] == a.outputs[0] && g.initializers.count(g.nodes[rn].inputs[1])) {
            const auto shp = g.initializers.at(g.nodes[rn].inputs[1]).as_int64();
            const int64_t cout = g.initializers.at(g.nodes[chain.units.back().conv_node].inputs[1]).dims[0];
            const bool r_ok = shp.size() == 3 && shp[0] == 0 && (shp[1] == cout || shp[1] == 0) && shp[2] == -1 &&
                              g.nodes[rn].attr_i("allowzero", 0) == 0;
            const int tn = r_ok ? sole_consumer(g.nodes[rn].outputs[0]) : -1;
            if (tn >= 0 && !m->skip_[tn] && g.nodes[tn].op == "Transpose" && g.nodes[tn].inputs[0] == g.nodes[rn].outputs[0] &&
                g.nodes[tn].attr_ints("perm", {}) == std::vector<int64_t>{2, 0, 1}) {
              chain.tail.avg_node = an;
              chain.tail.reshape_node = rn;
              chain.tail.transpose_node = tn;
              chain.tail.kh = (int)ks[0];
              chain.tail.out_name = published(tn);
            }
          }
        }
      }
      for (int sn : stem_nodes) impl->tc_member[sn] = 1;
      for (size_t k = 0; k < chain.units.size(); ++k) {
        if (k > 0) impl->tc_member[chain.units[k].conv_node] = 1;
        if (chain.units[k].pool_node >= 0) impl->tc_member[chain.units[k].pool_node] = 1;
      }
      impl->tc_chains[i] = std::move(chain);
    }
    // chain-internal values are never published: drop their use counts
  }

  // upload float initializers
  for (const auto& kv : g.initializers) {
    if (kv.second.dtype != onnx::kFloat) continue;
    m->dev_weights_[kv.first] = upload(kv.second.raw.data(), std::max<size_t>(kv.second.raw.size(), 4));
    m->weight_bytes_ += kv.second.raw.size();
  }
  // per-node weight re-layouts
  for (int i = 0; i < nn_; ++i) {
    const Node& n = g.nodes[i];
    if (n.op == "MatMul" && n.inputs.size() == 2) {
      auto it = g.initializers.find(n.inputs[1]);
      if (it != g.initializers.end() && it->second.dims.size() == 2 && !m->dev_weights_t_.count(n.inputs[1])) {
        int64_t K = it->second.dims[0], N = it->second.dims[1];
        std::vector<float> t((size_t)(K * N));
        const float* w = it->second.f32();
        for (int64_t k = 0; k < K; ++k)
          for (int64_t c = 0; c < N; ++c) t[(size_t)(c * K + k)] = w[k * N + c];
        m->dev_weights_t_[n.inputs[1]] = upload(t.data(), t.size() * 4);
      }
    } else if (n.op == "GRU") {
      OCRS_CHECK(n.inputs.size() >= 3, kModelLoad, "GRU: missing inputs");
      auto wit = g.initializers.find(n.inputs[1]);
      auto rit = g.initializers.find(n.inputs[2]);
      OCRS_CHECK(wit != g.initializers.end() && rit != g.initializers.end(), kModelLoad,
                 "GRU: W and R must be initializers");
      int64_t D = wit->second.dims[0], H3 = wit->second.dims[1];
      std::vector<float> wb((size_t)(D * H3), 0.f), rb((size_t)(D * H3), 0.f);
      if (n.inputs.size() > 3 && !n.inputs[3].empty()) {
        auto bit = g.initializers.find(n.inputs[3]);
        OCRS_CHECK(bit != g.initializers.end(), kModelLoad, "GRU: B must be an initializer");
        const float* b = bit->second.f32();
        for (int64_t d = 0; d < D; ++d) {
          std::memcpy(&wb[(size_t)(d * H3)], b + d * 2 * H3, (size_t)H3 * 4);
          std::memcpy(&rb[(size_t)(d * H3)], b + d * 2 * H3 + H3, (size_t)H3 * 4);
        }
      }
      impl->gru_bias[i] = {upload(wb.data(), wb.size() * 4), upload(rb.data(), rb.size() * 4)};
      {
        int64_t Hh = H3 / 3, Ii = wit->second.dims[2];
        if (impl->tc_enabled && std::getenv("OCRS_B200_DISABLE_TC_GRU") == nullptr &&
            tc::gru_supported((int)D, (int)Hh, (int)Ii)) {
          const float* bptr = (n.inputs.size() > 3 && !n.inputs[3].empty()) ? g.initializers.at(n.inputs[3]).f32() : nullptr;
          impl->tc_gru[i] = tc::prepare_gru(wit->second.f32(), rit->second.f32(), bptr, (int)D, (int)Hh, (int)Ii);
        }
      }
      OCRS_CHECK(n.attr_i("linear_before_reset", 0) != 0, kModelLoad,
                 "GRU with linear_before_reset=0 is not supported (PyTorch exports use 1)");
    }
  }
  // ---- packed sequence head detection (walk back from the graph output) ----
  if (impl->tc_enabled && std::getenv("OCRS_B200_DISABLE_SEQ_HEAD") == nullptr) {
    auto producer = [&](const std::string& name) -> int {
      for (int j = 0; j < nn_; ++j)
        for (const auto& o : g.nodes[j].outputs)
          if (o == name) return j;
      return -1;
    };
    auto& hd = impl->head;
    bool ok = true;
    int cur = producer(g.outputs[0].name);
    ok = ok && cur >= 0 && g.nodes[cur].op == "LogSoftmax";
    if (ok) {
      int64_t ax = g.nodes[cur].attr_i("axis", -1);
      ok = (ax == 2 || ax == -1);
    }
    int addn = ok ? producer(g.nodes[cur].inputs[0]) : -1;
    ok = ok && addn >= 0 && g.nodes[addn].op == "Add" && m->skip_[addn];  // folded into its MatMul
    int mm = -1;
    if (ok) {
      for (int j = 0; j < nn_; ++j)
        if (g.nodes[j].op == "MatMul" && impl->out_rename[j] == g.nodes[addn].outputs[0]) mm = j;
      ok = mm >= 0 && !m->fused_bias_[mm].empty();
    }
    std::string v = ok ? g.nodes[mm].inputs[0] : std::string();
    std::vector<int> grus;
    while (ok) {
      int rs = producer(v);
      if (rs < 0 || g.nodes[rs].op != "Reshape") break;
      auto sh = g.initializers.find(g.nodes[rs].inputs[1]);
      if (sh == g.initializers.end()) { ok = false; break; }
      auto shp = sh->second.as_int64();
      if (!(shp.size() == 3 && shp[0] == 0 && shp[1] == 0 && shp[2] == -1)) { ok = false; break; }
      int tp = producer(g.nodes[rs].inputs[0]);
      if (tp < 0 || g.nodes[tp].op != "Transpose" || g.nodes[tp].attr_ints("perm", {}) != std::vector<int64_t>{0, 2, 1, 3}) { ok = false; break; }
      int gn = producer(g.nodes[tp].inputs[0]);
      if (gn < 0 || g.nodes[gn].op != "GRU" || g.nodes[gn].outputs[0] != g.nodes[tp].inputs[0] || !impl->tc_gru.count(gn) ||
          g.nodes[gn].attr_s("direction", "forward") != "bidirectional") { ok = false; break; }
      // initial_h must be absent or an all-zero ConstantOfShape
      if (g.nodes[gn].inputs.size() > 5 && !g.nodes[gn].inputs[5].empty()) {
        int hp = producer(g.nodes[gn].inputs[5]);
        if (hp < 0 || g.nodes[hp].op != "ConstantOfShape") { ok = false; break; }
        const Attr* a = g.nodes[hp].find("value");
        if (a && a->kind == Attr::kTensor && a->t.dtype == onnx::kFloat && a->t.numel() > 0 && a->t.f32()[0] != 0.f) { ok = false; break; }
      }
      if (g.nodes[gn].inputs.size() > 4 && !g.nodes[gn].inputs[4].empty()) { ok = false; break; }
      // Y_h must be unused
      if (g.nodes[gn].outputs.size() > 1 && !g.nodes[gn].outputs[1].empty() && consumers.count(g.nodes[gn].outputs[1])) { ok = false; break; }
      grus.push_back(gn);
      v = g.nodes[gn].inputs[0];
    }
    ok = ok && !grus.empty();
    if (ok) {
      std::reverse(grus.begin(), grus.end());
      hd.present = true;
      hd.x_name = v;
      hd.gru_nodes = grus;
      hd.fc_node = mm;
      hd.H = impl->tc_gru.at(grus[0])->H;
      hd.C = impl->tc_gru.at(grus[0])->I;
      hd.classes = (int)g.initializers.at(g.nodes[mm].inputs[1]).dims[1];
      for (size_t k = 1; k < grus.size(); ++k)
        if (impl->tc_gru.at(grus[k])->I != 2 * hd.H) hd.present = false;
      if ((int)g.initializers.at(g.nodes[mm].inputs[1]).dims[0] != 2 * hd.H) hd.present = false;
      if (hd.present && tc::linear_supported(hd.classes, 2 * hd.H) && std::getenv("OCRS_B200_DISABLE_TC_LINEAR") == nullptr) {
        const auto& wt = g.initializers.at(g.nodes[mm].inputs[1]);  // [K][N]
        const int64_t K = wt.dims[0], N = wt.dims[1];
        std::vector<float> t((size_t)(K * N));
        const float* wsrc = wt.f32();
        for (int64_t k = 0; k < K; ++k)
          for (int64_t c = 0; c < N; ++c) t[(size_t)(c * K + k)] = wsrc[k * N + c];
        impl->head_fc = tc::prepare_linear(t.data(), g.initializers.at(m->fused_bias_[mm]).f32(), (int)N, (int)K);
      }
    }
  }

  m->impl_ = std::move(impl);
  return m;
}

DTensor Model::run(const DTensor& input, cudaStream_t st, ModelCost* cost, Profiler* prof,
                   const std::string& prof_prefix, const std::string* stop_at) const {
  const Impl* impl = impl_.get();
  const auto& g = graph_;
  OCRS_CHECK(input.shape.size() == input_shape_.size(), kRunFailed,
             "input rank " + std::to_string(input.shape.size()) + " does not match model input rank " +
                 std::to_string(input_shape_.size()));
  for (size_t d = 0; d < input_shape_.size(); ++d)
    OCRS_CHECK(input_shape_[d] < 0 || input_shape_[d] == input.shape[d], kRunFailed,
               "input dim " + std::to_string(d) + " = " + std::to_string(input.shape[d]) + " but model expects " +
                   std::to_string(input_shape_[d]));

  std::map<std::string, Value> env;
  std::map<std::string, int> remaining = use_count_;
  env[g.inputs[0].name] = dev_value(input);
  double flops = 0;

  auto get = [&](const std::string& name) -> Value {
    auto it = env.find(name);
    if (it != env.end()) return it->second;
    auto ii = g.initializers.find(name);
    OCRS_CHECK(ii != g.initializers.end(), kRunFailed, "value not found: " + name);
    const TensorData& td = ii->second;
    if (td.dtype == onnx::kFloat) {
      Value v;
      v.shape = td.dims;
      v.t.shape = td.dims;
      v.t.storage = dev_weights_.at(name);
      v.t.data = reinterpret_cast<float*>(v.t.storage->ptr);
      v.host_f32 = &td;
      return v;
    }
    return int_value(td.as_int64(), td.dims);
  };

  std::vector<char> dyn_skip(g.nodes.size(), 0);  // nodes absorbed by a run-time fusion of this run
  // one decision for the whole graph walk: another thread's overflow may switch the chains off while this
  // run is in flight (the run then finishes on the tensor cores and is caught by take_tc_overflow(token))
  const bool tc_on = impl->tc_conv_on.load(std::memory_order_acquire);
  for (int ni = 0; ni < (int)g.nodes.size(); ++ni) {
    if (skip_[ni] || dyn_skip[ni] || (impl->tc_member[ni] && tc_on)) continue;
    const Node& n = g.nodes[ni];
    auto chain_it = tc_on ? impl->tc_chains.find(ni) : impl->tc_chains.end();
    if (chain_it != impl->tc_chains.end()) {
      int* ovf = impl->tc_ovf.as<int>();
      // ---- Conv3x3(+ReLU)(+MaxPool) chain on the tensor cores (NHWC split-fp16) ----
      const TcChain& ch = chain_it->second;
      Value X = get(n.inputs[0]);
      OCRS_CHECK(!X.is_int && X.shape.size() == 4, kRunFailed, "Conv: expected 4-D input");
      int N_ = (int)X.shape[0], C_ = (int)X.shape[1], H_ = (int)X.shape[2], W_ = (int)X.shape[3];
      int ptok = prof ? prof->begin(prof_prefix + "ConvChain(total)", st) : -1;
      const double flops_before = flops;
      auto alloc_half = [&](int64_t elems) { return std::make_shared<Storage>((size_t)elems * 2, st); };
      std::shared_ptr<Storage> cur_hi, cur_lo;
      if (ch.stem) {
        OCRS_CHECK(C_ == 1, kRunFailed, "Conv: channel mismatch");
        int Co = ch.stem->Cout;
        flops += 2.0 * N_ * H_ * W_ * (double)Co * 9.0;
        int64_t oe = (int64_t)N_ * (H_ / 2) * (W_ / 2) * Co;
        cur_hi = alloc_half(oe);
        cur_lo = alloc_half(oe);
        tc::stem_conv_relu_pool2(X.t.data, *ch.stem, (tc::act_t*)cur_hi->ptr, (tc::act_t*)cur_lo->ptr, N_, H_, W_, ovf, st);
        H_ /= 2; W_ /= 2; C_ = Co;
      } else {
        int64_t elems = (int64_t)N_ * H_ * W_ * C_;
        cur_hi = alloc_half(elems);
        cur_lo = alloc_half(elems);
        tc::nchw_to_nhwc_split(X.t.data, (tc::act_t*)cur_hi->ptr, (tc::act_t*)cur_lo->ptr, N_, C_, H_, W_, ovf, st);
      }
      OCRS_CHECK(C_ == ch.units[0].w->Cin, kRunFailed, "Conv: channel mismatch");
      for (const TcUnit& u : ch.units) {
        int Co = u.w->Cout;
        const bool has_pool = u.pool_node >= 0;
        const bool fuse_pool = has_pool && (u.ph == 1 || u.ph == 2) && (u.pw == 1 || u.pw == 2);
        const int fph = fuse_pool ? u.ph : 1, fpw = fuse_pool ? u.pw : 1;
        int64_t oe = (int64_t)N_ * (H_ / fph) * (W_ / fpw) * Co;
        auto o_hi = alloc_half(oe), o_lo = alloc_half(oe);
        static const bool prof_layers = std::getenv("OCRS_B200_PROF_LAYERS") != nullptr;  // per-layer op names (diagnostics)
        int ktok = prof ? prof->begin(prof_prefix + "conv3x3_tc_kernel" +
                                          (prof_layers ? "/" + std::to_string(C_) + "->" + std::to_string(Co) + "@" + std::to_string(H_) : std::string()),
                                      st)
                        : -1;
        tc::conv3x3((const tc::act_t*)cur_hi->ptr, (const tc::act_t*)cur_lo->ptr, *u.w, (tc::act_t*)o_hi->ptr,
                    (tc::act_t*)o_lo->ptr, N_, H_, W_, u.relu, fph, fpw, ovf, st);
        const double cf = 2.0 * N_ * H_ * W_ * (double)Co * C_ * 9.0;
        if (prof) prof->end(ktok, st, cf, 4.0 * N_ * (H_ * W_ * (double)C_ + (H_ / fph) * (W_ / fpw) * (double)Co));
        flops += cf;
        cur_hi = o_hi; cur_lo = o_lo; C_ = Co; H_ /= fph; W_ /= fpw;
        if (has_pool && !fuse_pool) {
          int OH = H_ / u.ph, OW = W_ / u.pw;
          int64_t pe = (int64_t)N_ * OH * OW * C_;
          auto p_hi = alloc_half(pe), p_lo = alloc_half(pe);
          tc::maxpool_nhwc_split((const tc::act_t*)cur_hi->ptr, (const tc::act_t*)cur_lo->ptr,
                                 (tc::act_t*)p_hi->ptr, (tc::act_t*)p_lo->ptr, N_, H_, W_, C_, u.ph, u.pw, st);
          cur_hi = p_hi; cur_lo = p_lo; H_ = OH; W_ = OW;
        }
      }
      if (ch.tail.avg_node >= 0 && ch.tail.kh == H_) {
        // conv stack -> feature sequence in one pass: mean over the height, [W, N, C]
        DTensor S = alloc_tensor({W_, N_, C_}, st);
        int ttok = prof ? prof->begin(prof_prefix + "AvgPool+Transpose(fused)", st) : -1;
        tc::nhwc_split_avg_to_seq((const tc::act_t*)cur_hi->ptr, (const tc::act_t*)cur_lo->ptr, S.data, N_, C_, H_, W_, st);
        if (prof) prof->end(ttok, st, 0, 4.0 * N_ * W_ * (double)C_ * (H_ + 1));
        if (prof) prof->end(ptok, st, flops - flops_before);
        dyn_skip[ch.tail.avg_node] = dyn_skip[ch.tail.reshape_node] = dyn_skip[ch.tail.transpose_node] = 1;
        env[ch.tail.out_name] = dev_value(S);
        if (stop_at && *stop_at == ch.tail.out_name) {
          if (cost) { cost->flops = flops; cost->min_bytes = 0; }
          return S;
        }
        auto it0 = remaining.find(n.inputs[0]);
        if (it0 != remaining.end() && --it0->second <= 0) env.erase(n.inputs[0]);
        continue;
      }
      DTensor Y = alloc_tensor({N_, C_, H_, W_}, st);
      tc::nhwc_split_to_nchw((const tc::act_t*)cur_hi->ptr, (const tc::act_t*)cur_lo->ptr, Y.data, N_, C_, H_, W_, st);
      if (prof) prof->end(ptok, st, flops - flops_before);
      env[ch.out_name] = dev_value(Y);
      if (stop_at && *stop_at == ch.out_name) {
        if (cost) { cost->flops = flops; cost->min_bytes = 0; }
        return Y;
      }
      auto it0 = remaining.find(n.inputs[0]);
      if (it0 != remaining.end() && --it0->second <= 0) env.erase(n.inputs[0]);
      continue;
    }
    std::vector<Value> in;
    std::vector<bool> present;
    const auto sep_it = impl->sep_plan.find(ni);
    for (size_t ii = 0; ii < n.inputs.size(); ++ii) {
      const std::string& name = n.inputs[ii];
      const bool virt = ii == 0 && sep_it != impl->sep_plan.end();  // absorbed Concat: never materialised
      present.push_back(!name.empty());
      in.push_back((name.empty() || virt) ? Value() : get(name));
    }
    auto has = [&](size_t i) { return i < in.size() && present[i]; };
    std::vector<Value> out;
    const std::string& op = n.op;
    const double flops_before = flops;
    int ptok = prof ? prof->begin(prof_prefix + op, st) : -1;

    if (op == "Conv" && impl->dwpw_partner[ni] >= 0) {
      // depthwise 3x3 + pointwise 1x1 (+ReLU) in one kernel; the input may be a virtual Concat of padded tensors
      const Node& pn = g.nodes[impl->dwpw_partner[ni]];
      Value PW = get(pn.inputs[1]);
      const float* dwb = has(2) ? in[2].t.data : nullptr;
      const float* pwb = (pn.inputs.size() > 2 && !pn.inputs[2].empty()) ? get(pn.inputs[2]).t.data : nullptr;
      nn::SepInput srcs[2];
      std::vector<Value> keep;  // sources stay alive until the launch is enqueued
      int n_src = 0, N_ = 0, C_ = 0, H_ = 0, W_ = 0;
      if (sep_it != impl->sep_plan.end()) {
        const auto& plan = sep_it->second;
        for (size_t k = 0; k < plan.src.size(); ++k) {
          keep.push_back(get(plan.src[k]));
          const Value& S = keep.back();
          OCRS_CHECK(!S.is_int && S.shape.size() == 4, kRunFailed, "Concat: expected 4-D float inputs");
          const int hv = (int)S.shape[2] + plan.pad_b[k], wv = (int)S.shape[3] + plan.pad_r[k];
          OCRS_CHECK(hv > 0 && wv > 0, kRunFailed, "Pad: negative output dim");
          if (k == 0) { N_ = (int)S.shape[0]; H_ = hv; W_ = wv; }
          OCRS_CHECK((int)S.shape[0] == N_ && hv == H_ && wv == W_, kRunFailed, "Concat: shape mismatch");
          srcs[n_src++] = nn::SepInput{S.t.data, (int)S.shape[1], (int)S.shape[2], (int)S.shape[3], plan.pad_v[k]};
          C_ += (int)S.shape[1];
        }
      } else {
        const Value& X = in[0];
        OCRS_CHECK(!X.is_int && X.shape.size() == 4, kRunFailed, "Conv: expected 4-D input");
        N_ = (int)X.shape[0]; C_ = (int)X.shape[1]; H_ = (int)X.shape[2]; W_ = (int)X.shape[3];
        srcs[n_src++] = nn::SepInput{X.t.data, C_, H_, W_, 0.f};
      }
      OCRS_CHECK(in[1].shape[0] == C_, kRunFailed, "Conv: channel mismatch");
      int K_ = (int)PW.shape[0];
      int stride = (int)n.attr_ints("strides", {1, 1})[0];
      int OH = (H_ + 2 - 3) / stride + 1, OW = (W_ + 2 - 3) / stride + 1;
      DTensor Y = alloc_tensor({N_, K_, OH, OW}, st);
      nn::dwpw_conv2(srcs, n_src, in[1].t.data, dwb, PW.t.data, pwb, Y.data, N_, H_, W_, K_, stride,
                     fuse_relu_[impl->dwpw_partner[ni]], st);
      flops += 2.0 * N_ * OH * OW * (double)C_ * (9.0 + K_);
      out.push_back(dev_value(Y));
    } else if (op == "Conv") {
      const Value& X = in[0];
      const Value& W = in[1];
      OCRS_CHECK(!X.is_int && X.shape.size() == 4 && W.shape.size() == 4, kRunFailed, "Conv: expected 4-D x/w");
      OCRS_CHECK(n.attr_s("auto_pad", "NOTSET") == "NOTSET", kRunFailed, "Conv: auto_pad unsupported");
      auto pads = n.attr_ints("pads", {0, 0, 0, 0});
      auto strides = n.attr_ints("strides", {1, 1});
      auto dil = n.attr_ints("dilations", {1, 1});
      nn::ConvParams p;
      p.N = (int)X.shape[0]; p.C = (int)X.shape[1]; p.H = (int)X.shape[2]; p.W = (int)X.shape[3];
      p.K = (int)W.shape[0]; p.R = (int)W.shape[2]; p.S = (int)W.shape[3];
      p.groups = (int)n.attr_i("group", 1);
      OCRS_CHECK(W.shape[1] * p.groups == p.C, kRunFailed, "Conv: channel mismatch");
      p.stride_h = (int)strides[0]; p.stride_w = (int)strides[1];
      p.pad_t = (int)pads[0]; p.pad_l = (int)pads[1];
      p.dil_h = (int)dil[0]; p.dil_w = (int)dil[1];
      p.OH = (int)((p.H + pads[0] + pads[2] - dil[0] * (p.R - 1) - 1) / strides[0] + 1);
      p.OW = (int)((p.W + pads[1] + pads[3] - dil[1] * (p.S - 1) - 1) / strides[1] + 1);
      p.relu = fuse_relu_[ni];
      DTensor Y = alloc_tensor({p.N, p.K, p.OH, p.OW}, st);
      nn::conv2d(X.t.data, W.t.data, has(2) ? in[2].t.data : nullptr, Y.data, p, st);
      flops += 2.0 * p.N * p.K * p.OH * p.OW * (double)(p.C / p.groups) * p.R * p.S;
      out.push_back(dev_value(Y));
    } else if (op == "ConvTranspose") {
      const Value& X = in[0];
      const Value& W = in[1];
      OCRS_CHECK(X.shape.size() == 4 && W.shape.size() == 4, kRunFailed, "ConvTranspose: expected 4-D x/w");
      auto pads = n.attr_ints("pads", {0, 0, 0, 0});
      auto strides = n.attr_ints("strides", {1, 1});
      auto opad = n.attr_ints("output_padding", {0, 0});
      auto dil = n.attr_ints("dilations", {1, 1});
      OCRS_CHECK(dil[0] == 1 && dil[1] == 1, kRunFailed, "ConvTranspose: dilation unsupported");
      nn::ConvTParams p;
      p.N = (int)X.shape[0]; p.C = (int)X.shape[1]; p.H = (int)X.shape[2]; p.W = (int)X.shape[3];
      p.groups = (int)n.attr_i("group", 1);
      p.K = (int)W.shape[1] * p.groups; p.R = (int)W.shape[2]; p.S = (int)W.shape[3];
      OCRS_CHECK(W.shape[0] == p.C, kRunFailed, "ConvTranspose: channel mismatch");
      p.stride_h = (int)strides[0]; p.stride_w = (int)strides[1];
      p.pad_t = (int)pads[0]; p.pad_l = (int)pads[1];
      p.OH = (int)((p.H - 1) * strides[0] - pads[0] - pads[2] + p.R + opad[0]);
      p.OW = (int)((p.W - 1) * strides[1] - pads[1] - pads[3] + p.S + opad[1]);
      p.relu = fuse_relu_[ni];
      const bool k2s2 = p.R == 2 && p.S == 2 && p.stride_h == 2 && p.stride_w == 2 && pads[0] == 0 && pads[1] == 0 &&
                        pads[2] == 0 && pads[3] == 0 && opad[0] == 0 && opad[1] == 0 && p.groups == 1;
      auto hf = impl->head_fuse.find(ni);
      if (hf != impl->head_fuse.end() && k2s2) {
        const Node& cn = g.nodes[hf->second.conv1x1];
        Value W2 = get(cn.inputs[1]);
        const float* b2 = (cn.inputs.size() > 2 && !cn.inputs[2].empty()) ? get(cn.inputs[2]).t.data : nullptr;
        DTensor Y = alloc_tensor({p.N, 1, p.OH, p.OW}, st);
        nn::conv_transpose_2x2s2_head(X.t.data, W.t.data, has(2) ? in[2].t.data : nullptr, W2.t.data, b2, Y.data, p.N,
                                      p.C, p.H, p.W, p.K, st);
        flops += 2.0 * p.N * p.H * p.W * 4.0 * ((double)p.C * p.K + p.K);
        out.push_back(dev_value(Y));
      } else {
        OCRS_CHECK(hf == impl->head_fuse.end(), kInternal, "head fusion planned for an unsupported ConvTranspose");
        DTensor Y = alloc_tensor({p.N, p.K, p.OH, p.OW}, st);
        if (k2s2) nn::conv_transpose_2x2s2(X.t.data, W.t.data, has(2) ? in[2].t.data : nullptr, Y.data, p.N, p.C, p.H, p.W, p.K, p.relu, st);
        else nn::conv_transpose2d(X.t.data, W.t.data, has(2) ? in[2].t.data : nullptr, Y.data, p, st);
        flops += 2.0 * p.N * p.C * p.H * p.W * (double)(p.K / p.groups) * p.R * p.S;
        out.push_back(dev_value(Y));
      }
    } else if (op == "MaxPool" || op == "AveragePool") {
      const Value& X = in[0];
      OCRS_CHECK(X.shape.size() == 4, kRunFailed, op + ": expected 4-D input");
      OCRS_CHECK(n.attr_i("ceil_mode", 0) == 0, kRunFailed, op + ": ceil_mode unsupported");
      auto ks = n.attr_ints("kernel_shape", {});
      OCRS_CHECK(ks.size() == 2, kRunFailed, op + ": kernel_shape must be 2-D");
      auto pads = n.attr_ints("pads", {0, 0, 0, 0});
      auto strides = n.attr_ints("strides", {1, 1});
      nn::PoolParams p;
      p.NC = (int)(X.shape[0] * X.shape[1]); p.H = (int)X.shape[2]; p.W = (int)X.shape[3];
      p.R = (int)ks[0]; p.S = (int)ks[1];
      p.stride_h = (int)strides[0]; p.stride_w = (int)strides[1];
      p.pad_t = (int)pads[0]; p.pad_l = (int)pads[1];
      p.OH = (int)((p.H + pads[0] + pads[2] - p.R) / strides[0] + 1);
      p.OW = (int)((p.W + pads[1] + pads[3] - p.S) / strides[1] + 1);
      p.count_include_pad = (int)n.attr_i("count_include_pad", 0);
      DTensor Y = alloc_tensor({X.shape[0], X.shape[1], p.OH, p.OW}, st);
      if (op == "MaxPool") nn::max_pool2d(X.t.data, Y.data, p, st);
      else nn::avg_pool2d(X.t.data, Y.data, p, st);
      out.push_back(dev_value(Y));
    } else if (op == "Relu" || op == "Sigmoid" || op == "Tanh") {
      const Value& X = in[0];
      DTensor Y = alloc_tensor(X.shape, st);
      if (op == "Relu") nn::relu(X.t.data, Y.data, X.numel(), st);
      else if (op == "Sigmoid") nn::sigmoid(X.t.data, Y.data, X.numel(), st);
      else nn::tanh_op(X.t.data, Y.data, X.numel(), st);
      out.push_back(dev_value(Y));
    } else if (op == "Add") {
      if (in[0].is_int && in[1].is_int) {
        const Value& A = in[0].numel() >= in[1].numel() ? in[0] : in[1];
        const Value& B = in[0].numel() >= in[1].numel() ? in[1] : in[0];
        OCRS_CHECK(B.numel() > 0 && A.numel() % B.numel() == 0, kRunFailed, "Add: int broadcast unsupported");
        std::vector<int64_t> r(A.ivals);
        for (size_t i = 0; i < r.size(); ++i) r[i] += B.ivals[i % B.ivals.size()];
        out.push_back(int_value(r, A.shape));
      } else {
        const Value& A = in[0].numel() >= in[1].numel() ? in[0] : in[1];
        const Value& B = in[0].numel() >= in[1].numel() ? in[1] : in[0];
        OCRS_CHECK(!A.is_int && !B.is_int, kRunFailed, "Add: mixed int/float");
        // B's shape must be a suffix of A's (after stripping leading 1s)
        std::vector<int64_t> bs = B.shape;
        while (!bs.empty() && bs.front() == 1) bs.erase(bs.begin());
        OCRS_CHECK(bs.size() <= A.shape.size() &&
                       std::equal(bs.rbegin(), bs.rend(), A.shape.rbegin()),
                   kRunFailed, "Add: only suffix broadcasting is supported");
        DTensor Y = alloc_tensor(A.shape, st);
        nn::add_bcast_suffix(A.t.data, B.t.data, Y.data, A.numel(), std::max<int64_t>(B.numel(), 1), st);
        out.push_back(dev_value(Y));
      }
    } else if (op == "MatMul") {
      const Value& A = in[0];
      const Value& B = in[1];
      OCRS_CHECK(!A.is_int && !B.is_int && B.shape.size() == 2 && !A.shape.empty(), kRunFailed,
                 "MatMul: expected [..., K] x [K, N]");
      int64_t K = B.shape[0], N = B.shape[1];
      OCRS_CHECK(A.shape.back() == K, kRunFailed, "MatMul: inner dimension mismatch");
      int64_t M = A.numel() / K;
      std::vector<int64_t> oshape(A.shape.begin(), A.shape.end() - 1);
      oshape.push_back(N);
      DTensor Y = alloc_tensor(oshape, st);
      const float* Bt;
      DTensor tmp;
      auto wt = dev_weights_t_.find(n.inputs[1]);
      if (wt != dev_weights_t_.end()) {
        Bt = reinterpret_cast<const float*>(wt->second->ptr);
      } else {
        tmp = alloc_tensor({N, K}, st);
        int perm[2] = {1, 0};
        nn::permute(B.t.data, tmp.data, B.shape.data(), perm, 2, st);
        Bt = tmp.data;
      }
      const float* bias = nullptr;
      if (!fused_bias_[ni].empty()) bias = reinterpret_cast<const float*>(dev_weights_.at(fused_bias_[ni])->ptr);
      nn::sgemm_nt(A.t.data, Bt, bias, Y.data, (int)M, (int)N, (int)K, 0, st);
      flops += 2.0 * M * N * K;
      out.push_back(dev_value(Y));
    } else if (op == "LogSoftmax") {
      const Value& X = in[0];
      int64_t axis = norm_axis(n.attr_i("axis", -1), (int64_t)X.shape.size());
      OCRS_CHECK(axis == (int64_t)X.shape.size() - 1, kRunFailed, "LogSoftmax: only the last axis is supported");
      DTensor Y = alloc_tensor(X.shape, st);
      int cols = (int)X.shape.back();
      nn::log_softmax_lastdim(X.t.data, Y.data, X.numel() / std::max(cols, 1), cols, st);
      out.push_back(dev_value(Y));
    } else if (op == "Concat") {
      OCRS_CHECK(!in.empty(), kRunFailed, "Concat: no inputs");
      if (in[0].is_int) {
        std::vector<int64_t> r;
        for (auto& v : in) {
          OCRS_CHECK(v.is_int && v.shape.size() <= 1, kRunFailed, "Concat: int inputs must be 1-D");
          r.insert(r.end(), v.ivals.begin(), v.ivals.end());
        }
        int64_t len = (int64_t)r.size();
        out.push_back(int_value(std::move(r), {len}));
      } else {
        int64_t nd = (int64_t)in[0].shape.size();
        int64_t axis = norm_axis(n.attr_i("axis", 0), nd);
        std::vector<int64_t> oshape = in[0].shape;
        oshape[axis] = 0;
        for (auto& v : in) {
          OCRS_CHECK(!v.is_int && (int64_t)v.shape.size() == nd, kRunFailed, "Concat: rank mismatch");
          for (int64_t d = 0; d < nd; ++d)
            OCRS_CHECK(d == axis || v.shape[d] == in[0].shape[d], kRunFailed, "Concat: shape mismatch");
          oshape[axis] += v.shape[axis];
        }
        DTensor Y = alloc_tensor(oshape, st);
        int64_t outer = 1, inner = 1;
        for (int64_t d = 0; d < axis; ++d) outer *= oshape[d];
        for (int64_t d = axis + 1; d < nd; ++d) inner *= oshape[d];
        int64_t off = 0;
        for (auto& v : in) {
          nn::concat_copy(v.t.data, Y.data, outer, v.shape[axis], oshape[axis], off, inner, st);
          off += v.shape[axis];
        }
        out.push_back(dev_value(Y));
      }
    } else if (op == "Transpose") {
      const Value& X = in[0];
      int nd = (int)X.shape.size();
      std::vector<int64_t> dflt;
      for (int d = nd - 1; d >= 0; --d) dflt.push_back(d);
      auto perm64 = n.attr_ints("perm", dflt);
      OCRS_CHECK((int)perm64.size() == nd && !X.is_int, kRunFailed, "Transpose: bad perm");
      std::vector<int> perm(perm64.begin(), perm64.end());
      std::vector<int64_t> oshape(nd);
      for (int d = 0; d < nd; ++d) oshape[d] = X.shape[perm[d]];
      DTensor Y = alloc_tensor(oshape, st);
      nn::permute(X.t.data, Y.data, X.shape.data(), perm.data(), nd, st);
      out.push_back(dev_value(Y));
    } else if (op == "Reshape") {
      const Value& X = in[0];
      OCRS_CHECK(in[1].is_int, kRunFailed, "Reshape: shape must be an int tensor");
      std::vector<int64_t> shp = in[1].ivals;
      bool allowzero = n.attr_i("allowzero", 0) != 0;
      int64_t known = 1, infer = -1;
      for (size_t d = 0; d < shp.size(); ++d) {
        if (shp[d] == 0 && !allowzero) {
          OCRS_CHECK(d < X.shape.size(), kRunFailed, "Reshape: 0 dim out of range");
          shp[d] = X.shape[d];
        }
        if (shp[d] == -1) infer = (int64_t)d;
        else known *= shp[d];
      }
      if (infer >= 0) shp[infer] = known ? X.numel() / known : 0;
      Value v = X;
      v.shape = shp;
      int64_t cnt = 1;
      for (auto d : shp) cnt *= d;
      OCRS_CHECK(cnt == X.numel(), kRunFailed, "Reshape: element count mismatch");
      if (!v.is_int) v.t.shape = shp;
      out.push_back(v);
    } else if (op == "Identity") {
      out.push_back(in[0]);
    } else if (op == "Constant") {
      const Attr* a = n.find("value");
      OCRS_CHECK(a && a->kind == Attr::kTensor, kRunFailed, "Constant: only tensor values are supported");
      if (a->t.dtype == onnx::kFloat) {
        DTensor Y = alloc_tensor(a->t.dims, st);
        OCRS_CUDA_CHECK(cudaMemcpyAsync(Y.data, a->t.raw.data(), a->t.raw.size(), cudaMemcpyHostToDevice, st));
        OCRS_CUDA_CHECK(cudaStreamSynchronize(st));
        Value v = dev_value(Y);
        v.host_f32 = &a->t;
        out.push_back(v);
      } else {
        out.push_back(int_value(a->t.as_int64(), a->t.dims));
      }
    } else if (op == "Shape") {
      const Value& X = in[0];
      int64_t nd = (int64_t)X.shape.size();
      int64_t s = n.attr_i("start", 0), e = n.attr_i("end", nd);
      if (s < 0) s += nd;
      if (e < 0) e += nd;
      s = std::min(std::max<int64_t>(s, 0), nd);
      e = std::min(std::max<int64_t>(e, 0), nd);
      std::vector<int64_t> r(X.shape.begin() + s, X.shape.begin() + std::max(s, e));
      int64_t len = (int64_t)r.size();
      out.push_back(int_value(std::move(r), {len}));
    } else if (op == "Gather") {
      OCRS_CHECK(in[0].is_int && in[1].is_int && in[0].shape.size() == 1, kRunFailed,
                 "Gather: only 1-D integer data is supported (shape arithmetic)");
      OCRS_CHECK(norm_axis(n.attr_i("axis", 0), 1) == 0, kRunFailed, "Gather: bad axis");
      std::vector<int64_t> r;
      for (auto idx : in[1].ivals) {
        if (idx < 0) idx += (int64_t)in[0].ivals.size();
        OCRS_CHECK(idx >= 0 && idx < (int64_t)in[0].ivals.size(), kRunFailed, "Gather: index out of range");
        r.push_back(in[0].ivals[(size_t)idx]);
      }
      out.push_back(int_value(std::move(r), in[1].shape));
    } else if (op == "Unsqueeze" || op == "Squeeze") {
      Value v = in[0];
      std::vector<int64_t> axes;
      if (has(1)) axes = in[1].ivals;
      else axes = n.attr_ints("axes", {});
      std::vector<int64_t> shp = v.shape;
      if (op == "Unsqueeze") {
        int64_t nd = (int64_t)shp.size() + (int64_t)axes.size();
        for (auto& a : axes) a = norm_axis(a, nd);
        std::sort(axes.begin(), axes.end());
        for (auto a : axes) shp.insert(shp.begin() + a, 1);
      } else {
        int64_t nd = (int64_t)shp.size();
        if (axes.empty()) {
          std::vector<int64_t> s2;
          for (auto d : shp) if (d != 1) s2.push_back(d);
          shp = s2;
        } else {
          for (auto& a : axes) a = norm_axis(a, nd);
          std::sort(axes.rbegin(), axes.rend());
          for (auto a : axes) {
            OCRS_CHECK(shp[a] == 1, kRunFailed, "Squeeze: dim is not 1");
            shp.erase(shp.begin() + a);
          }
        }
      }
      v.shape = shp;
      if (!v.is_int) v.t.shape = shp;
      out.push_back(v);
    } else if (op == "Slice") {
      const Value& X = in[0];
      OCRS_CHECK(has(1) && has(2) && in[1].is_int && in[2].is_int, kRunFailed, "Slice: starts/ends required");
      size_t k = in[1].ivals.size();
      std::vector<int64_t> axes(k), steps(k, 1);
      for (size_t i = 0; i < k; ++i) axes[i] = (int64_t)i;
      if (has(3)) axes = in[3].ivals;
      if (has(4)) steps = in[4].ivals;
      int64_t nd = (int64_t)X.shape.size();
      std::vector<int64_t> begin(nd, 0), size = X.shape;
      for (size_t i = 0; i < k; ++i) {
        OCRS_CHECK(steps[i] == 1, kRunFailed, "Slice: only step 1 is supported");
        int64_t ax = norm_axis(axes[i], nd), dim = X.shape[ax];
        int64_t s = in[1].ivals[i], e = in[2].ivals[i];
        if (s < 0) s += dim;
        if (e < 0) e += dim;
        s = std::min(std::max<int64_t>(s, 0), dim);
        e = std::min(std::max<int64_t>(e, 0), dim);
        begin[ax] = s;
        size[ax] = std::max<int64_t>(e - s, 0);
      }
      if (X.is_int) {
        OCRS_CHECK(nd == 1, kRunFailed, "Slice: int data must be 1-D");
        std::vector<int64_t> r(X.ivals.begin() + begin[0], X.ivals.begin() + begin[0] + size[0]);
        out.push_back(int_value(std::move(r), {size[0]}));
      } else {
        OCRS_CHECK(nd <= 4, kRunFailed, "Slice: rank > 4");
        int64_t ish[4] = {1, 1, 1, 1}, osh[4] = {1, 1, 1, 1}, bg[4] = {0, 0, 0, 0};
        for (int64_t d = 0; d < nd; ++d) {
          ish[4 - nd + d] = X.shape[d];
          osh[4 - nd + d] = size[d];
          bg[4 - nd + d] = -begin[d];
        }
        DTensor Y = alloc_tensor(size, st);
        nn::pad4d(X.t.data, Y.data, ish, bg, osh, 0.f, st);
        out.push_back(dev_value(Y));
      }
    } else if (op == "Cast") {
      int64_t to = n.attr_i("to", onnx::kFloat);
      const Value& X = in[0];
      if (X.is_int) {
        OCRS_CHECK(to == onnx::kInt64 || to == onnx::kInt32, kRunFailed, "Cast: int -> float unsupported");
        out.push_back(X);
      } else {
        OCRS_CHECK(to == onnx::kFloat, kRunFailed, "Cast: float -> int unsupported");
        out.push_back(X);
      }
    } else if (op == "ConstantOfShape") {
      OCRS_CHECK(in[0].is_int, kRunFailed, "ConstantOfShape: shape must be ints");
      const Attr* a = n.find("value");
      if (a && a->kind == Attr::kTensor && a->t.dtype != onnx::kFloat) {
        int64_t cnt = 1;
        for (auto d : in[0].ivals) cnt *= d;
        auto vals = a->t.as_int64();
        out.push_back(int_value(std::vector<int64_t>((size_t)cnt, vals.empty() ? 0 : vals[0]), in[0].ivals));
      } else {
        float v = (a && a->kind == Attr::kTensor && a->t.numel() > 0) ? a->t.f32()[0] : 0.f;
        DTensor Y = alloc_tensor(in[0].ivals, st);
        nn::fill(Y.data, v, Y.numel(), st);
        out.push_back(dev_value(Y));
      }
    } else if (op == "Pad") {
      const Value& X = in[0];
      OCRS_CHECK(n.attr_s("mode", "constant") == "constant", kRunFailed, "Pad: only constant mode");
      std::vector<int64_t> pads;
      if (has(1)) pads = in[1].ivals;
      else pads = n.attr_ints("pads", {});
      float value = n.attr_f("value", 0.f);
      if (has(2)) {
        OCRS_CHECK(in[2].host_f32 != nullptr, kRunFailed, "Pad: constant_value must be an initializer");
        if (in[2].host_f32->numel() > 0) value = in[2].host_f32->f32()[0];
      }
      int64_t nd = (int64_t)X.shape.size();
      OCRS_CHECK(nd <= 4 && !X.is_int, kRunFailed, "Pad: rank > 4");
      std::vector<int64_t> axes;
      if (has(3)) axes = in[3].ivals;
      else for (int64_t d = 0; d < nd; ++d) axes.push_back(d);
      OCRS_CHECK(pads.size() == 2 * axes.size(), kRunFailed, "Pad: pads size mismatch");
      std::vector<int64_t> begin(nd, 0), end(nd, 0);
      for (size_t i = 0; i < axes.size(); ++i) {
        int64_t ax = norm_axis(axes[i], nd);
        begin[ax] = pads[i];
        end[ax] = pads[i + axes.size()];
      }
      std::vector<int64_t> oshape(nd);
      int64_t ish[4] = {1, 1, 1, 1}, osh[4] = {1, 1, 1, 1}, bg[4] = {0, 0, 0, 0};
      for (int64_t d = 0; d < nd; ++d) {
        oshape[d] = X.shape[d] + begin[d] + end[d];
        OCRS_CHECK(oshape[d] >= 0, kRunFailed, "Pad: negative output dim");
        ish[4 - nd + d] = X.shape[d];
        osh[4 - nd + d] = oshape[d];
        bg[4 - nd + d] = begin[d];
      }
      DTensor Y = alloc_tensor(oshape, st);
      nn::pad4d(X.t.data, Y.data, ish, bg, osh, value, st);
      out.push_back(dev_value(Y));
    } else if (op == "GRU") {
      const Value& X = in[0];
      const Value& W = in[1];
      const Value& R = in[2];
      OCRS_CHECK(X.shape.size() == 3 && W.shape.size() == 3 && R.shape.size() == 3, kRunFailed, "GRU: bad ranks");
      OCRS_CHECK(!has(4), kRunFailed, "GRU: sequence_lens unsupported");
      OCRS_CHECK(n.attr_i("layout", 0) == 0, kRunFailed, "GRU: layout=1 unsupported");
      int T = (int)X.shape[0], N = (int)X.shape[1], I = (int)X.shape[2];
      int D = (int)W.shape[0], H = (int)n.attr_i("hidden_size", W.shape[1] / 3);
      OCRS_CHECK(W.shape[1] == 3 * H && W.shape[2] == I && R.shape[1] == 3 * H && R.shape[2] == H, kRunFailed,
                 "GRU: weight shape mismatch");
      std::string dir = n.attr_s("direction", "forward");
      OCRS_CHECK((dir == "bidirectional") == (D == 2), kRunFailed, "GRU: direction/num_directions mismatch");
      int rev[2] = {dir == "reverse" ? 1 : 0, 1};
      const auto& gb = impl->gru_bias.at(ni);
      const float* Wb = reinterpret_cast<const float*>(gb.first->ptr);
      const float* Rb = reinterpret_cast<const float*>(gb.second->ptr);
      auto tcg = impl->tc_gru.find(ni);
      if (tcg != impl->tc_gru.end()) {
        OCRS_CHECK(!has(5) || in[5].numel() == (int64_t)D * N * H, kRunFailed, "GRU: initial_h shape mismatch");
        DTensor Y = alloc_tensor({T, D, N, H}, st);
        bool want_yh = n.outputs.size() > 1 && !n.outputs[1].empty() && remaining.count(n.outputs[1]);
        DTensor Yh;
        if (want_yh) Yh = alloc_tensor({D, N, H}, st);
        std::vector<std::shared_ptr<Storage>> scratch;
        tc::gru_forward(X.t.data, *tcg->second, has(5) ? in[5].t.data : nullptr, Y.data, want_yh ? Yh.data : nullptr, T, N,
                        rev, [&](size_t bytes) { scratch.push_back(std::make_shared<Storage>(bytes, st)); return scratch.back()->ptr; },
                        st);
        flops += 2.0 * D * T * N * 3.0 * H * (I + H);
        out.push_back(dev_value(Y));
        if (want_yh) out.push_back(dev_value(Yh));
      } else {
      DTensor xw = alloc_tensor({D, (int64_t)T * N, 3 * H}, st);
      for (int d = 0; d < D; ++d)
        nn::sgemm_nt(X.t.data, W.t.data + (int64_t)d * 3 * H * I, Wb + (int64_t)d * 3 * H,
                     xw.data + (int64_t)d * T * N * 3 * H, T * N, 3 * H, I, 0, st);
      DTensor hA = alloc_tensor({D, N, H}, st), hB = alloc_tensor({D, N, H}, st);
      if (has(5)) {
        OCRS_CHECK(in[5].numel() == (int64_t)D * N * H, kRunFailed, "GRU: initial_h shape mismatch");
        OCRS_CUDA_CHECK(cudaMemcpyAsync(hA.data, in[5].t.data, sizeof(float) * D * N * H, cudaMemcpyDeviceToDevice, st));
      } else {
        nn::fill(hA.data, 0.f, (int64_t)D * N * H, st);
      }
      DTensor Y = alloc_tensor({T, D, N, H}, st);
      float* hin = hA.data;
      float* hout = hB.data;
      for (int s = 0; s < T; ++s) {
        nn::gru_step(xw.data, R.t.data, Rb, hin, hout, Y.data, D, T, N, H, s, rev, 1, st);
        std::swap(hin, hout);
      }
      flops += 2.0 * D * T * N * 3.0 * H * (I + H);
      out.push_back(dev_value(Y));
      if (n.outputs.size() > 1 && !n.outputs[1].empty() && remaining.count(n.outputs[1])) {
        DTensor Yh = alloc_tensor({D, N, H}, st);
        OCRS_CUDA_CHECK(cudaMemcpyAsync(Yh.data, hin, sizeof(float) * D * N * H, cudaMemcpyDeviceToDevice, st));
        out.push_back(dev_value(Yh));
      }
      }
    } else {
      throw Error(kRunFailed, "unsupported operator at run time: " + op);
    }

    if (prof) prof->end(ptok, st, flops - flops_before);
    // publish outputs
    for (size_t k = 0; k < out.size() && k < n.outputs.size(); ++k) {
      std::string name = n.outputs[k];
      if (k == 0 && !impl->out_rename[ni].empty()) name = impl->out_rename[ni];
      if (!name.empty()) env[name] = std::move(out[k]);
    }
    if (stop_at) {
      auto sit = env.find(*stop_at);
      if (sit != env.end()) {
        if (cost) { cost->flops = flops; cost->min_bytes = 0; }
        DTensor r = sit->second.t;
        r.shape = sit->second.shape;
        return r;
      }
    }
    // release inputs whose last consumer just ran
    const std::vector<std::string>& eff_inputs = sep_it != impl->sep_plan.end() ? sep_it->second.src : n.inputs;
    for (const auto& name : eff_inputs) {
      if (name.empty()) continue;
      auto it = remaining.find(name);
      if (it != remaining.end() && --it->second <= 0) env.erase(name);
    }
  }
  auto it = env.find(g.outputs[0].name);
  OCRS_CHECK(it != env.end() && !it->second.is_int, kWrongOutput, "graph output was not produced");
  if (cost) {
    cost->flops = flops;
    cost->min_bytes = 4.0 * (double)(input.numel() + it->second.t.numel()) + (double)weight_bytes_;
  }
  DTensor result = it->second.t;
  result.shape = it->second.shape;
  return result;
}

bool Model::has_seq_head() const { return impl_->head.present; }
int Model::seq_head_channels() const { return impl_->head.C; }
int Model::seq_head_classes() const { return impl_->head.classes; }

int Model::tc_token() const {
  const Impl* impl = impl_.get();
  return (impl->tc_enabled && impl->tc_ovf.ptr && impl->tc_conv_on.load(std::memory_order_acquire)) ? 1 : 0;
}

bool Model::take_tc_overflow(int token) const {
  Impl* impl = impl_.get();
  if (token == 0) return false;  // the run started on the fp32 kernels: nothing to check
  // the run started with the tensor-core chains on.  If another run has switched them off in the meantime,
  // this run's output may be saturated too (the flag it raised may already have been consumed): repeat it.
  if (!impl->tc_conv_on.load(std::memory_order_acquire)) return true;
  std::lock_guard<std::mutex> lk(impl->ovf_mu);
  if (!impl->tc_conv_on.load(std::memory_order_acquire)) return true;
  int flag = 0;
  OCRS_CUDA_CHECK(cudaMemcpy(&flag, impl->tc_ovf.ptr, 4, cudaMemcpyDeviceToHost));
  if (!flag) return false;
  impl->tc_conv_on.store(false, std::memory_order_release);
  OCRS_CUDA_CHECK(cudaMemset(impl->tc_ovf.ptr, 0, 4));
  return true;
}

DTensor Model::run_prefix(const DTensor& input, cudaStream_t st, ModelCost* cost, Profiler* prof,
                          const std::string& prof_prefix) const {
  OCRS_CHECK(impl_->head.present, kInternal, "run_prefix: model has no packed sequence head");
  return run(input, st, cost, prof, prof_prefix, &impl_->head.x_name);
}

// The packed conv prefix: graph input -> [stem + tensor-core conv chain + pooling tail] -> X_seq, i.e. the
// whole prefix is one TcChain.  Then every layer runs ONCE over all width groups (ragged launches).
static const TcChain* packed_chain(const Model::Impl* impl, const onnx::Graph& g);

bool Model::has_packed_prefix(int in_h) const {
  const Impl* impl = impl_.get();
  if (!impl->head.present || !impl->tc_conv_on.load(std::memory_order_relaxed)) return false;
  if (std::getenv("OCRS_B200_DISABLE_RAGGED") != nullptr) return false;
  const TcChain* ch = packed_chain(impl, graph_);
  if (!ch) return false;
  int h = in_h / 2;
  for (const TcUnit& u : ch->units) {
    if (u.pool_node >= 0 && !((u.ph == 1 || u.ph == 2) && (u.pw == 1 || u.pw == 2))) return false;
    h /= (u.pool_node >= 0 ? u.ph : 1);
  }
  return h == ch->tail.kh && h >= 1;
}

static const TcChain* packed_chain(const Model::Impl* impl, const onnx::Graph& g) {
  for (const auto& kv : impl->tc_chains) {
    const TcChain& ch = kv.second;
    if (!ch.stem || ch.tail.avg_node < 0 || ch.tail.out_name != impl->head.x_name) continue;
    if (g.nodes[kv.first].inputs.empty() || g.nodes[kv.first].inputs[0] != g.inputs[0].name) continue;
    return &ch;
  }
  return nullptr;
}

DTensor Model::run_prefix_packed(const std::vector<PrefixGroup>& groups, int in_h, cudaStream_t st,
                                 std::vector<PackedGroup>* out_groups, ModelCost* cost, Profiler* prof,
                                 const std::string& prof_prefix) const {
  const Impl* impl = impl_.get();
  const TcChain* chp = packed_chain(impl, graph_);
  OCRS_CHECK(chp != nullptr && has_packed_prefix(in_h), kInternal, "run_prefix_packed: model has no packed conv prefix");
  const TcChain& ch = *chp;
  const int G = (int)groups.size();
  const int U = (int)ch.units.size();
  OCRS_CHECK(G > 0, kInvalidArg, "run_prefix_packed: no groups");
  int* ovf = impl->tc_ovf.as<int>();
  double flops = 0;

  // ---- plan: per layer and group, dims and pixel offsets (layer 0 = stem, 1..U = convs, U+1 = tail) ----
  const int L = U + 2;
  std::vector<tc::RaggedDesc> descs((size_t)L * G);
  std::vector<int64_t> layer_pix((size_t)U + 1, 0);   // output pixels of the stem and of each conv
  std::vector<int> layer_c((size_t)U + 1, 0);         // their channel counts
  std::vector<int> layer_units((size_t)L, 0);         // tiles / blocks per launch
  std::vector<int> gh((size_t)G), gw((size_t)G);
  layer_c[0] = ch.stem->Cout;
  {
    int64_t pix = 0;
    int blocks = 0;
    for (int g = 0; g < G; ++g) {
      OCRS_CHECK(groups[g].N > 0 && groups[g].W > 0, kInvalidArg, "run_prefix_packed: empty group");
      tc::RaggedDesc& d = descs[(size_t)g];
      d = tc::RaggedDesc{};
      d.N = groups[g].N; d.H = in_h; d.W = groups[g].W; d.OH = in_h / 2; d.OW = groups[g].W / 2;
      d.first = blocks;
      d.in_off = groups[g].x_off;
      d.out_off = pix;

      pix += (int64_t)d.N * d.OH * d.OW;
      blocks += (int)ceil_div((int64_t)d.N * d.OH * d.OW, 128);
      gh[g] = d.OH; gw[g] = d.OW;
      flops += 2.0 * d.N * d.H * d.W * (double)ch.stem->Cout * 9.0;
    }
    layer_pix[0] = pix;
    layer_units[0] = blocks;
  }
  std::vector<double> conv_flops((size_t)U, 0), conv_bytes((size_t)U, 0);
  for (int u = 0; u < U; ++u) {
    const TcUnit& un = ch.units[(size_t)u];
    const int fph = un.pool_node >= 0 ? un.ph : 1, fpw = un.pool_node >= 0 ? un.pw : 1;
    int64_t pix = 0;
    int tiles = 0;
    for (int g = 0; g < G; ++g) {
      tc::RaggedDesc& d = descs[(size_t)(u + 1) * G + g];
      d = tc::RaggedDesc{};
      d.N = groups[g].N; d.H = gh[g]; d.W = gw[g]; d.OH = gh[g] / fph; d.OW = gw[g] / fpw;
      tc::conv_fill_tiles(&d);
      d.first = tiles;
      d.in_off = descs[(size_t)u * G + g].out_off;
      d.out_off = pix;
      pix += (int64_t)d.N * d.OH * d.OW;
      tiles += tc::conv_tiles(d.N, d.H, d.W);
      const double cf = 2.0 * d.N * d.H * d.W * (double)un.w->Cout * un.w->Cin * 9.0;
      conv_flops[u] += cf;
      conv_bytes[u] += 4.0 * d.N * ((double)d.H * d.W * un.w->Cin + (double)d.OH * d.OW * un.w->Cout);
      gh[g] = d.OH; gw[g] = d.OW;
    }
    layer_pix[(size_t)u + 1] = pix;
    layer_c[(size_t)u + 1] = un.w->Cout;
    layer_units[(size_t)u + 1] = tiles;
    flops += conv_flops[u];
  }
  const int Cf = layer_c[(size_t)U];
  int64_t rows = 0;
  out_groups->clear();
  {
    int blocks = 0;
    for (int g = 0; g < G; ++g) {
      tc::RaggedDesc& d = descs[(size_t)(U + 1) * G + g];
      d = tc::RaggedDesc{};
      d.N = groups[g].N; d.H = gh[g]; d.W = gw[g]; d.OH = 1; d.OW = gw[g];
      OCRS_CHECK(d.H == ch.tail.kh, kInternal, "run_prefix_packed: pooled height does not match the average pool");
      d.first = blocks;
      d.in_off = descs[(size_t)U * G + g].out_off;
      d.out_off = rows;
      blocks += (int)ceil_div((int64_t)d.N * d.W * (Cf / 8), 256);
      out_groups->push_back(PackedGroup{d.W, d.N, rows});
      rows += (int64_t)d.W * d.N;
    }
    layer_units[(size_t)U + 1] = blocks;
  }

  // ---- activations: two ping-pong buffers (hi + lo planes each) ----
  int64_t cap[2] = {0, 0};
  for (int l = 0; l <= U; ++l) cap[l & 1] = std::max(cap[l & 1], layer_pix[(size_t)l] * layer_c[(size_t)l]);
  std::shared_ptr<Storage> buf_hi[2], buf_lo[2];
  for (int k = 0; k < 2; ++k) {
    buf_hi[k] = std::make_shared<Storage>((size_t)std::max<int64_t>(cap[k], 8) * 2, st);
    buf_lo[k] = std::make_shared<Storage>((size_t)std::max<int64_t>(cap[k], 8) * 2, st);
  }
  auto hi_of = [&](int l) { return reinterpret_cast<tc::act_t*>(buf_hi[l & 1]->ptr); };
  auto lo_of = [&](int l) { return reinterpret_cast<tc::act_t*>(buf_lo[l & 1]->ptr); };

  // ---- one device blob: [descs | counters | tensor maps] ----
  const size_t desc_bytes = round_up((int64_t)(descs.size() * sizeof(tc::RaggedDesc)), 128);
  const size_t ctr_bytes = round_up((int64_t)U * 4, 128);
  const size_t map_bytes = (size_t)U * 2 * G * sizeof(CUtensorMap);
  std::vector<uint8_t> blob(desc_bytes + ctr_bytes + map_bytes, 0);
  std::memcpy(blob.data(), descs.data(), descs.size() * sizeof(tc::RaggedDesc));
  auto* hmaps = reinterpret_cast<CUtensorMap*>(blob.data() + desc_bytes + ctr_bytes);
  for (int u = 0; u < U; ++u) {
    const int Cin = ch.units[(size_t)u].w->Cin;
    OCRS_CHECK(Cin == layer_c[(size_t)u], kRunFailed, "Conv: channel mismatch");
    for (int g = 0; g < G; ++g) {
      const tc::RaggedDesc& d = descs[(size_t)(u + 1) * G + g];
      tc::make_act_maps(hi_of(u) + d.in_off * Cin, lo_of(u) + d.in_off * Cin, d.N, d.H, d.W, Cin, hmaps + ((size_t)u * G + g) * 2,
                        ch.units[(size_t)u].w->Cout);
    }
  }
  auto blob_dev = std::make_shared<Storage>(blob.size() + 128, st);
  uint8_t* dblob = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(blob_dev->ptr) + 127) & ~(uintptr_t)127);
  OCRS_CUDA_CHECK(cudaMemcpyAsync(dblob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
  const auto* d_descs = reinterpret_cast<const tc::RaggedDesc*>(dblob);
  int* d_ctr = reinterpret_cast<int*>(dblob + desc_bytes);
  const auto* d_maps = reinterpret_cast<const CUtensorMap*>(dblob + desc_bytes + ctr_bytes);

  int ptok = prof ? prof->begin(prof_prefix + "ConvChain(total)", st) : -1;
  {
    int tk = prof ? prof->begin(prof_prefix + "stem_kernel", st) : -1;
    tc::stem_ragged(groups[0].x_base, *ch.stem, hi_of(0), lo_of(0), d_descs, G, layer_units[0], ovf, st);
    if (prof) prof->end(tk, st, 0, 0);
  }
  for (int u = 0; u < U; ++u) {
    const TcUnit& un = ch.units[(size_t)u];
    const int fph = un.pool_node >= 0 ? un.ph : 1, fpw = un.pool_node >= 0 ? un.pw : 1;
    static const bool prof_layers = std::getenv("OCRS_B200_PROF_LAYERS") != nullptr;  // per-layer op names (diagnostics)
    int ktok = prof ? prof->begin(prof_prefix + "conv3x3_tc_kernel" +
                                      (prof_layers ? "/" + std::to_string(un.w->Cin) + "->" + std::to_string(un.w->Cout) + "#" + std::to_string(u) : std::string()),
                                  st)
                    : -1;
    tc::conv3x3_ragged(d_maps + (size_t)u * 2 * G, d_descs + (size_t)(u + 1) * G, G, layer_units[(size_t)u + 1], d_ctr + u, *un.w,
                       hi_of(u + 1), lo_of(u + 1), un.relu, fph, fpw, ovf, st);
    if (prof) prof->end(ktok, st, conv_flops[u], conv_bytes[u]);
  }
  DTensor S = alloc_tensor({rows, (int64_t)Cf}, st);
  {
    int tk = prof ? prof->begin(prof_prefix + "AvgPool+Transpose(fused)", st) : -1;
    tc::avg_to_seq_ragged(hi_of(U), lo_of(U), S.data, Cf, d_descs + (size_t)(U + 1) * G, G, layer_units[(size_t)U + 1], st);
    if (prof) prof->end(tk, st, 0, 4.0 * (double)rows * Cf * (ch.tail.kh + 1));
  }
  if (prof) prof->end(ptok, st, flops);
  if (cost) { cost->flops = flops; cost->min_bytes = 0; }
  // buf_*, blob_dev: stream-ordered frees behind the launches above
  return S;
}

DTensor Model::run_seq_head(const float* X, int64_t rows, const std::vector<PackedGroup>& groups, cudaStream_t st,
                            ModelCost* cost, Profiler* prof, const std::string& prof_prefix) const {
  const auto& hd = impl_->head;
  OCRS_CHECK(hd.present, kInternal, "run_seq_head: model has no packed sequence head");
  const int H = hd.H;
  double flops = 0;
  // ragged line list, longest first (stable): tiles of 32 lines run max-T steps
  struct L { int T, N; int64_t row0; };
  std::vector<L> ls;
  for (const auto& gph : groups)
    for (int n = 0; n < gph.N; ++n) ls.push_back(L{gph.T, gph.N, gph.row_off + n});
  std::stable_sort(ls.begin(), ls.end(), [](const L& a, const L& b) { return a.T > b.T; });
  std::vector<std::shared_ptr<Storage>> scratch;
  auto alloc = [&](size_t bytes) { scratch.push_back(std::make_shared<Storage>(bytes, st)); return scratch.back()->ptr; };
  const float* cur = X;
  std::shared_ptr<Storage> cur_store;
  int rev[2] = {0, 1};
  for (size_t li = 0; li < hd.gru_nodes.size(); ++li) {
    const tc::GruWeightsTC& w = *impl_->tc_gru.at(hd.gru_nodes[li]);
    const int64_t Ntot = (int64_t)w.D * 3 * H;
    std::vector<tc::SeqLine> lines(ls.size());
    for (size_t i = 0; i < ls.size(); ++i) {
      lines[i].T = ls[i].T;
      lines[i].valid = 1;
      lines[i].xw_base = ls[i].row0 * Ntot;
      lines[i].xw_tstride = (int64_t)ls[i].N * Ntot;
      lines[i].y_base = ls[i].row0 * 2 * H;
      lines[i].y_tstride = (int64_t)ls[i].N * 2 * H;
    }
    auto y_store = std::make_shared<Storage>((size_t)rows * 2 * H * 4, st);
    // two profiler entries: the input projection of all timesteps (split + GEMM) and the recurrence
    const double f_proj = 2.0 * w.D * (double)rows * 3.0 * H * w.I, f_rec = 2.0 * w.D * (double)rows * 3.0 * H * H;
    int tk = prof ? prof->begin(prof_prefix + "GRU input projection(packed)", st) : -1;
    const std::function<void()> mark = [&] {
      if (prof) {
        prof->end(tk, st, f_proj);
        tk = prof->begin(prof_prefix + "GRU recurrence(packed)", st);
      }
    };
    tc::gru_forward_lines(cur, rows, w, lines.data(), (int)lines.size(), reinterpret_cast<float*>(y_store->ptr), H, rev,
                          alloc, st, &mark);
    flops += f_proj + f_rec;
    if (prof) prof->end(tk, st, f_rec);
    cur_store = y_store;
    cur = reinterpret_cast<const float*>(y_store->ptr);
  }
  // Linear + LogSoftmax over all rows
  const Node& fc = graph_.nodes[hd.fc_node];
  const float* Wt = reinterpret_cast<const float*>(dev_weights_t_.at(fc.inputs[1])->ptr);
  const float* bias = reinterpret_cast<const float*>(dev_weights_.at(fused_bias_[hd.fc_node])->ptr);
  DTensor logits = alloc_tensor({rows, hd.classes}, st);
  int tk = prof ? prof->begin(prof_prefix + "Linear+LogSoftmax(packed)", st) : -1;
  if (impl_->head_fc) {
    // split-bf16 tensor-core GEMM into a 128-column padded buffer, log-softmax over the real classes
    const int npad = impl_->head_fc->Npad;
    DTensor tmp = alloc_tensor({rows, npad}, st);
    tc::linear_forward(cur, rows, *impl_->head_fc, tmp.data, alloc, st);
    nn::log_softmax_rows(tmp.data, npad, logits.data, rows, hd.classes, st);
  } else {
    DTensor tmp = alloc_tensor({rows, hd.classes}, st);
    nn::sgemm_nt(cur, Wt, bias, tmp.data, (int)rows, hd.classes, 2 * H, 0, st);
    nn::log_softmax_lastdim(tmp.data, logits.data, rows, hd.classes, st);
  }
  double f = 2.0 * (double)rows * hd.classes * 2 * H;
  flops += f;
  if (prof) prof->end(tk, st, f);
  if (cost) { cost->flops = flops; cost->min_bytes = 0; }
  return logits;
}

}  // namespace ocrs
