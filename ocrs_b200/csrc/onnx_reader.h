// ONNX model file -> in-memory graph.  Replaces `rten::Model::load_file` / `ModelOptions::load`
// (reference call sites: ocrs-cli/src/models.rs:105, ocrs/src/wasm_api.rs:62-64) for `.onnx`
// input.  Hand-written protobuf wire parser: no protobuf/onnx dependency exists in this image.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace ocrs {
namespace onnx {

enum DType : int { kFloat = 1, kUint8 = 2, kInt8 = 3, kInt32 = 6, kInt64 = 7, kBool = 9 };

struct TensorData {
  std::vector<int64_t> dims;
  int dtype = 0;
  std::vector<uint8_t> raw;  // little-endian element bytes
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
  const float* f32() const { return reinterpret_cast<const float*>(raw.data()); }
  const int64_t* i64() const { return reinterpret_cast<const int64_t*>(raw.data()); }
  std::vector<int64_t> as_int64() const;  // converts int32/int64 data
};

struct Attr {
  enum Kind { kNone, kInt, kFloatK, kString, kTensor, kInts, kFloats } kind = kNone;
  int64_t i = 0;
  float f = 0.f;
  std::string s;
  TensorData t;
  std::vector<int64_t> ints;
  std::vector<float> floats;
};

struct Node {
  std::string op, name;
  std::vector<std::string> inputs, outputs;
  std::map<std::string, Attr> attrs;
  int64_t attr_i(const std::string& k, int64_t dflt) const;
  float attr_f(const std::string& k, float dflt) const;
  std::string attr_s(const std::string& k, const std::string& dflt) const;
  std::vector<int64_t> attr_ints(const std::string& k, const std::vector<int64_t>& dflt) const;
  const Attr* find(const std::string& k) const;
};

struct ValueInfo {
  std::string name;
  int elem_type = 0;
  std::vector<int64_t> dims;            // -1 for symbolic
  std::vector<std::string> dim_params;  // "" for fixed
};

struct Graph {
  std::string name;
  std::vector<Node> nodes;
  std::map<std::string, TensorData> initializers;
  std::vector<ValueInfo> inputs, outputs;  // inputs exclude initializers
  int64_t opset = 0;
};

// Throws ocrs::Error(kModelLoad) on malformed input; the returned graph has passed validate_graph().
// validate_graph: arity of every supported operator, lengths / ranges of the attribute vectors and the
// weight shapes that Model::load and Model::run index -- an untrusted file cannot drive them out of bounds.
void validate_graph(const Graph& g);
Graph parse_model(const uint8_t* bytes, size_t len);

// True if the buffer looks like an rten container ("RTEN" magic).
bool looks_like_rten(const uint8_t* bytes, size_t len);

}  // namespace onnx
}  // namespace ocrs
