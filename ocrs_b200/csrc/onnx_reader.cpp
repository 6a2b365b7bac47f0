// Protobuf wire-format reader for the subset of onnx.proto3 the engine needs.
#include "onnx_reader.h"

#include <cstring>

#include "common.h"

namespace ocrs {
namespace onnx {
namespace {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (true) {
      OCRS_CHECK(p < end && shift < 70, kModelLoad, "onnx: truncated varint");
      uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
    }
  }
  Reader sub() {
    uint64_t n = varint();
    OCRS_CHECK((uint64_t)(end - p) >= n, kModelLoad, "onnx: truncated length-delimited field");
    Reader r{p, p + n};
    p += n;
    return r;
  }
  void skip(int wt) {
    switch (wt) {
      case 0: varint(); break;
      case 1: OCRS_CHECK(end - p >= 8, kModelLoad, "onnx: truncated fixed64"); p += 8; break;
      case 2: sub(); break;
      case 5: OCRS_CHECK(end - p >= 4, kModelLoad, "onnx: truncated fixed32"); p += 4; break;
      default: throw Error(kModelLoad, "onnx: unsupported wire type");
    }
  }
  std::string str() {
    Reader r = sub();
    return std::string(reinterpret_cast<const char*>(r.p), r.end - r.p);
  }
  float fixed32f() {
    OCRS_CHECK(end - p >= 4, kModelLoad, "onnx: truncated float");
    float f;
    std::memcpy(&f, p, 4);
    p += 4;
    return f;
  }
};

size_t dtype_size(int dt) {
  switch (dt) {
    case kFloat: case kInt32: return 4;
    case kInt64: return 8;
    case kUint8: case kInt8: case kBool: return 1;
    default: throw Error(kModelLoad, "onnx: unsupported tensor data type " + std::to_string(dt));
  }
}

// Repeated scalar field that may be packed (wt 2) or not (wt 0).
void read_varints(Reader& r, int wt, std::vector<int64_t>& out) {
  if (wt == 2) {
    Reader s = r.sub();
    while (!s.done()) out.push_back((int64_t)s.varint());
  } else {
    out.push_back((int64_t)r.varint());
  }
}
void read_floats(Reader& r, int wt, std::vector<float>& out) {
  if (wt == 2) {
    Reader s = r.sub();
    while (!s.done()) out.push_back(s.fixed32f());
  } else {
    out.push_back(r.fixed32f());
  }
}

TensorData parse_tensor(Reader r, std::string* name) {
  TensorData t;
  std::vector<float> float_data;
  std::vector<int64_t> int32_data, int64_data;
  bool have_raw = false;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = (int)(key >> 3), wt = (int)(key & 7);
    switch (field) {
      case 1: read_varints(r, wt, t.dims); break;
      case 2: t.dtype = (int)r.varint(); break;
      case 4: read_floats(r, wt, float_data); break;
      case 5: read_varints(r, wt, int32_data); break;
      case 7: read_varints(r, wt, int64_data); break;
      case 8: { std::string s = r.str(); if (name) *name = s; break; }
      case 9: { Reader s = r.sub(); t.raw.assign(s.p, s.end); have_raw = true; break; }
      default: r.skip(wt);
    }
  }
  size_t es = dtype_size(t.dtype);
  // dims come from an untrusted file: no negative extents, no product that wraps (kernels are sized
  // from them); 2^40 elements is far beyond any model this engine loads
  OCRS_CHECK(t.dims.size() <= 8, kModelLoad, "onnx: tensor rank > 8");
  uint64_t n_checked = 1;
  for (int64_t d : t.dims) {
    OCRS_CHECK(d >= 0, kModelLoad, "onnx: negative tensor dimension");
    OCRS_CHECK(d == 0 || n_checked <= ((uint64_t)1 << 40) / (uint64_t)d, kModelLoad, "onnx: tensor too large");
    n_checked *= (uint64_t)d;
  }
  size_t n = (size_t)n_checked;
  if (!have_raw) {
    t.raw.resize(n * es);
    if (t.dtype == kFloat) {
      OCRS_CHECK(float_data.size() == n, kModelLoad, "onnx: float_data size mismatch");
      std::memcpy(t.raw.data(), float_data.data(), n * 4);
    } else if (t.dtype == kInt64) {
      OCRS_CHECK(int64_data.size() == n, kModelLoad, "onnx: int64_data size mismatch");
      std::memcpy(t.raw.data(), int64_data.data(), n * 8);
    } else {
      OCRS_CHECK(int32_data.size() == n, kModelLoad, "onnx: int32_data size mismatch");
      for (size_t i = 0; i < n; ++i) {
        if (es == 4) { int32_t v = (int32_t)int32_data[i]; std::memcpy(&t.raw[i * 4], &v, 4); }
        else t.raw[i] = (uint8_t)int32_data[i];
      }
    }
  }
  OCRS_CHECK(t.raw.size() == n * es, kModelLoad, "onnx: tensor byte size does not match dims");
  return t;
}

Attr parse_attr(Reader r, std::string* name) {
  Attr a;
  int type = 0;
  bool has_i = false, has_f = false, has_s = false, has_t = false;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = (int)(key >> 3), wt = (int)(key & 7);
    switch (field) {
      case 1: *name = r.str(); break;
      case 2: a.f = r.fixed32f(); has_f = true; break;
      case 3: a.i = (int64_t)r.varint(); has_i = true; break;
      case 4: a.s = r.str(); has_s = true; break;
      case 5: a.t = parse_tensor(r.sub(), nullptr); has_t = true; break;
      case 7: read_floats(r, wt, a.floats); break;
      case 8: read_varints(r, wt, a.ints); break;
      case 20: type = (int)r.varint(); break;
      default: r.skip(wt);
    }
  }
  switch (type) {
    case 1: a.kind = Attr::kFloatK; break;
    case 2: a.kind = Attr::kInt; break;
    case 3: a.kind = Attr::kString; break;
    case 4: a.kind = Attr::kTensor; break;
    case 6: a.kind = Attr::kFloats; break;
    case 7: a.kind = Attr::kInts; break;
    default:
      if (has_i) a.kind = Attr::kInt;
      else if (has_f) a.kind = Attr::kFloatK;
      else if (has_s) a.kind = Attr::kString;
      else if (has_t) a.kind = Attr::kTensor;
      else if (!a.ints.empty()) a.kind = Attr::kInts;
      else if (!a.floats.empty()) a.kind = Attr::kFloats;
  }
  return a;
}

Node parse_node(Reader r) {
  Node n;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = (int)(key >> 3), wt = (int)(key & 7);
    switch (field) {
      case 1: n.inputs.push_back(r.str()); break;
      case 2: n.outputs.push_back(r.str()); break;
      case 3: n.name = r.str(); break;
      case 4: n.op = r.str(); break;
      case 5: { std::string nm; Attr a = parse_attr(r.sub(), &nm); n.attrs[nm] = std::move(a); break; }
      default: r.skip(wt);
    }
  }
  return n;
}

ValueInfo parse_value_info(Reader r) {
  ValueInfo v;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = (int)(key >> 3), wt = (int)(key & 7);
    if (field == 1) {
      v.name = r.str();
    } else if (field == 2) {  // TypeProto
      Reader tp = r.sub();
      while (!tp.done()) {
        uint64_t k2 = tp.varint();
        if ((k2 >> 3) == 1) {  // tensor_type
          Reader tt = tp.sub();
          while (!tt.done()) {
            uint64_t k3 = tt.varint();
            int f3 = (int)(k3 >> 3);
            if (f3 == 1) {
              v.elem_type = (int)tt.varint();
            } else if (f3 == 2) {  // TensorShapeProto
              Reader sh = tt.sub();
              while (!sh.done()) {
                uint64_t k4 = sh.varint();
                if ((k4 >> 3) == 1) {
                  Reader dim = sh.sub();
                  int64_t val = -1;
                  std::string param;
                  while (!dim.done()) {
                    uint64_t k5 = dim.varint();
                    int f5 = (int)(k5 >> 3);
                    if (f5 == 1) val = (int64_t)dim.varint();
                    else if (f5 == 2) { param = dim.str(); val = -1; }
                    else dim.skip((int)(k5 & 7));
                  }
                  v.dims.push_back(val);
                  v.dim_params.push_back(param);
                } else {
                  sh.skip((int)(k4 & 7));
                }
              }
            } else {
              tt.skip((int)(k3 & 7));
            }
          }
        } else {
          tp.skip((int)(k2 & 7));
        }
      }
    } else {
      r.skip(wt);
    }
  }
  return v;
}

Graph parse_graph(Reader r) {
  Graph g;
  std::vector<ValueInfo> inputs;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = (int)(key >> 3), wt = (int)(key & 7);
    switch (field) {
      case 1: g.nodes.push_back(parse_node(r.sub())); break;
      case 2: g.name = r.str(); break;
      case 5: { std::string nm; TensorData t = parse_tensor(r.sub(), &nm); g.initializers[nm] = std::move(t); break; }
      case 11: inputs.push_back(parse_value_info(r.sub())); break;
      case 12: g.outputs.push_back(parse_value_info(r.sub())); break;
      default: r.skip(wt);
    }
  }
  for (auto& v : inputs)
    if (!g.initializers.count(v.name)) g.inputs.push_back(v);
  return g;
}

}  // namespace

std::vector<int64_t> TensorData::as_int64() const {
  std::vector<int64_t> out((size_t)numel());
  if (dtype == kInt64) {
    std::memcpy(out.data(), raw.data(), out.size() * 8);
  } else if (dtype == kInt32) {
    for (size_t i = 0; i < out.size(); ++i) { int32_t v; std::memcpy(&v, &raw[i * 4], 4); out[i] = v; }
  } else {
    throw Error(kModelLoad, "onnx: expected an integer tensor");
  }
  return out;
}

const Attr* Node::find(const std::string& k) const {
  auto it = attrs.find(k);
  return it == attrs.end() ? nullptr : &it->second;
}
int64_t Node::attr_i(const std::string& k, int64_t d) const { auto a = find(k); return a ? a->i : d; }
float Node::attr_f(const std::string& k, float d) const { auto a = find(k); return a ? a->f : d; }
std::string Node::attr_s(const std::string& k, const std::string& d) const { auto a = find(k); return a ? a->s : d; }
std::vector<int64_t> Node::attr_ints(const std::string& k, const std::vector<int64_t>& d) const {
  auto a = find(k);
  return a ? a->ints : d;
}

// Structural validation of an untrusted graph: every supported operator has the input / output arity and
// the attribute vector lengths that Model::load and Model::run index without further checks.
void validate_graph(const Graph& g) {
  auto need = [](const Node& n, size_t min_in, size_t max_in) {
    OCRS_CHECK(n.inputs.size() >= min_in && n.inputs.size() <= max_in, kModelLoad,
               "onnx: " + n.op + " node '" + n.name + "' has " + std::to_string(n.inputs.size()) + " inputs");
    for (size_t i = 0; i < min_in; ++i)
      OCRS_CHECK(!n.inputs[i].empty(), kModelLoad, "onnx: " + n.op + " node '" + n.name + "' misses a required input");
  };
  auto ints_len = [](const Node& n, const char* key, size_t len) {
    const Attr* a = n.find(key);
    if (!a) return;
    OCRS_CHECK(a->kind == Attr::kInts && a->ints.size() == len, kModelLoad,
               "onnx: " + n.op + " attribute '" + key + "' must hold " + std::to_string(len) + " ints");
  };
  OCRS_CHECK(g.nodes.size() <= (size_t)1 << 20, kModelLoad, "onnx: too many nodes");
  for (const Node& n : g.nodes) {
    OCRS_CHECK(!n.outputs.empty() && !n.outputs[0].empty(), kModelLoad, "onnx: node '" + n.name + "' (" + n.op + ") has no output");
    OCRS_CHECK(n.outputs.size() <= 4, kModelLoad, "onnx: node '" + n.name + "' has too many outputs");
    const std::string& op = n.op;
    if (op == "Conv" || op == "ConvTranspose") {
      need(n, 2, 3);
      ints_len(n, "pads", 4);
      ints_len(n, "strides", 2);
      ints_len(n, "dilations", 2);
      ints_len(n, "kernel_shape", 2);
      ints_len(n, "output_padding", 2);
      OCRS_CHECK(n.attr_i("group", 1) >= 1, kModelLoad, "onnx: " + op + " group must be >= 1");
      for (int64_t v : n.attr_ints("strides", {1, 1})) OCRS_CHECK(v >= 1 && v <= 64, kModelLoad, "onnx: bad stride");
      for (int64_t v : n.attr_ints("dilations", {1, 1})) OCRS_CHECK(v >= 1 && v <= 64, kModelLoad, "onnx: bad dilation");
      for (int64_t v : n.attr_ints("pads", {0, 0, 0, 0})) OCRS_CHECK(v >= 0 && v <= 4096, kModelLoad, "onnx: bad pad");
    } else if (op == "MaxPool" || op == "AveragePool") {
      need(n, 1, 1);
      const Attr* ks = n.find("kernel_shape");
      OCRS_CHECK(ks && ks->kind == Attr::kInts && ks->ints.size() == 2, kModelLoad, "onnx: " + op + " needs a 2-D kernel_shape");
      for (int64_t v : ks->ints) OCRS_CHECK(v >= 1 && v <= 4096, kModelLoad, "onnx: bad pool kernel");
      ints_len(n, "pads", 4);
      ints_len(n, "strides", 2);
      for (int64_t v : n.attr_ints("strides", {1, 1})) OCRS_CHECK(v >= 1 && v <= 4096, kModelLoad, "onnx: bad stride");
      for (int64_t v : n.attr_ints("pads", {0, 0, 0, 0})) OCRS_CHECK(v >= 0 && v <= 4096, kModelLoad, "onnx: bad pad");
    } else if (op == "GRU") {
      need(n, 3, 6);
      OCRS_CHECK(n.attr_i("hidden_size", 1) >= 1, kModelLoad, "onnx: GRU hidden_size must be >= 1");
    } else if (op == "MatMul" || op == "Add" || op == "Gather" || op == "Reshape") {
      need(n, 2, 2);
    } else if (op == "Slice") {
      need(n, 3, 5);
    } else if (op == "Pad") {
      need(n, 1, 4);
    } else if (op == "Concat") {
      need(n, 1, 1 << 16);
    } else if (op == "Unsqueeze" || op == "Squeeze") {
      need(n, 1, 2);
    } else if (op == "Constant") {
      need(n, 0, 0);
    } else if (op == "Relu" || op == "Sigmoid" || op == "Tanh" || op == "LogSoftmax" || op == "Transpose" ||
               op == "Identity" || op == "Shape" || op == "Cast" || op == "ConstantOfShape") {
      need(n, 1, 1);
    }
    // unknown operators are reported by Model::load / model_inspect with their name
  }
  // weights whose shape the loader indexes
  for (const Node& n : g.nodes) {
    auto init = [&](size_t i) -> const TensorData* {
      if (n.inputs.size() <= i || n.inputs[i].empty()) return nullptr;
      auto it = g.initializers.find(n.inputs[i]);
      return it == g.initializers.end() ? nullptr : &it->second;
    };
    if (n.op == "Conv" || n.op == "ConvTranspose") {
      if (const TensorData* w = init(1)) {
        OCRS_CHECK(w->dtype == kFloat && w->dims.size() == 4, kModelLoad, "onnx: " + n.op + " weight must be a 4-D float tensor");
        for (int64_t d : w->dims) OCRS_CHECK(d >= 1, kModelLoad, "onnx: " + n.op + " weight has an empty dimension");
        if (const TensorData* b = init(2))
          OCRS_CHECK(b->dtype == kFloat && b->dims.size() == 1 &&
                         b->dims[0] == (n.op == "Conv" ? w->dims[0] : w->dims[1] * n.attr_i("group", 1)),
                     kModelLoad, "onnx: " + n.op + " bias does not match its weight");
      }
    } else if (n.op == "GRU") {
      const TensorData* w = init(1);
      const TensorData* r = init(2);
      if (w && r) {
        OCRS_CHECK(w->dtype == kFloat && r->dtype == kFloat && w->dims.size() == 3 && r->dims.size() == 3, kModelLoad,
                   "onnx: GRU W / R must be 3-D float tensors");
        OCRS_CHECK(w->dims[0] >= 1 && w->dims[0] <= 2 && r->dims[0] == w->dims[0] && w->dims[1] % 3 == 0 && w->dims[1] >= 3 &&
                       r->dims[1] == w->dims[1] && r->dims[2] == w->dims[1] / 3 && w->dims[2] >= 1,
                   kModelLoad, "onnx: GRU weight shapes are inconsistent");
        if (const TensorData* b = init(3))
          OCRS_CHECK(b->dtype == kFloat && b->dims.size() == 2 && b->dims[0] == w->dims[0] && b->dims[1] == 2 * w->dims[1],
                     kModelLoad, "onnx: GRU bias shape is inconsistent");
      }
    } else if (n.op == "MatMul") {
      if (const TensorData* w = init(1)) OCRS_CHECK(w->dtype == kFloat, kModelLoad, "onnx: MatMul weight must be float");
    }
  }
}

bool looks_like_rten(const uint8_t* b, size_t len) {
  return (len >= 4 && std::memcmp(b, "RTEN", 4) == 0) || (len >= 8 && std::memcmp(b + 4, "RTEN", 4) == 0);
}

Graph parse_model(const uint8_t* bytes, size_t len) {
  OCRS_CHECK(bytes != nullptr && len > 0, kModelLoad, "empty model buffer");
  if (looks_like_rten(bytes, len)) {
    throw Error(kModelLoad,
                "model is in .rten (FlatBuffers) format; this build loads ONNX only -- the rten schema is not "
                "available offline (SURVEY.md section 8f, row N1). Convert with `rten-convert`'s inverse or "
                "supply the .onnx export.");
  }
  Reader r{bytes, bytes + len};
  Graph g;
  bool have_graph = false;
  int64_t opset = 0;
  while (!r.done()) {
    uint64_t key = r.varint();
    int field = (int)(key >> 3), wt = (int)(key & 7);
    if (field == 7 && wt == 2) {
      g = parse_graph(r.sub());
      have_graph = true;
    } else if (field == 8 && wt == 2) {
      Reader o = r.sub();
      std::string domain;
      int64_t version = 0;
      while (!o.done()) {
        uint64_t k2 = o.varint();
        if ((k2 >> 3) == 1) domain = o.str();
        else if ((k2 >> 3) == 2) version = (int64_t)o.varint();
        else o.skip((int)(k2 & 7));
      }
      if (domain.empty() || domain == "ai.onnx") opset = version;
    } else {
      r.skip(wt);
    }
  }
  OCRS_CHECK(have_graph, kModelLoad, "onnx: ModelProto has no graph");
  g.opset = opset;
  validate_graph(g);
  return g;
}

}  // namespace onnx
}  // namespace ocrs
