// extern "C" boundary (include/ocrs_b200.h).  Exceptions never cross it.
#include "../../include/ocrs_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>

#include "engine.h"
#include "executor.h"
#include "layout.h"
#include "onnx_reader.h"
#include "pool.h"
#include "text_output.h"

using namespace ocrs;

namespace {
thread_local std::string g_last_error;

template <typename F>
int guard(F&& f) {
  try {
    g_last_error.clear();
    f();
    return OCRS_B200_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    g_last_error = "out of host memory";
    return OCRS_B200_ERR_INTERNAL;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return OCRS_B200_ERR_INTERNAL;
  } catch (...) {
    g_last_error = "unknown error";
    return OCRS_B200_ERR_INTERNAL;
  }
}

template <typename T>
T* cmalloc(size_t n) {
  T* p = static_cast<T*>(std::malloc(std::max<size_t>(n, 1) * sizeof(T)));
  if (!p) throw std::bad_alloc();
  return p;
}

void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  OCRS_CHECK(e == cudaSuccess && n > 0, kNoDevice,
             std::string("no CUDA device available: ocrs_b200 has no CPU fallback (") + cudaGetErrorString(e) + ")");
}

ocrs_b200_text_result* make_text_result(const std::vector<TextLine>& lines) {
  auto* r = cmalloc<ocrs_b200_text_result>(1);
  size_t n = lines.size(), total = 0;
  for (const auto& l : lines) total += l.present ? l.chars.size() : 0;
  r->n_lines = (int32_t)n;
  r->line_present = cmalloc<uint8_t>(n);
  r->char_offsets = cmalloc<int64_t>(n + 1);
  r->chars = cmalloc<uint32_t>(total);
  r->char_rects = cmalloc<ocrs_b200_rect>(total);
  size_t k = 0;
  for (size_t i = 0; i < n; ++i) {
    r->line_present[i] = lines[i].present ? 1 : 0;
    r->char_offsets[i] = (int64_t)k;
    if (!lines[i].present) continue;
    for (const auto& c : lines[i].chars) {
      r->chars[k] = c.ch;
      r->char_rects[k] = ocrs_b200_rect{c.rect.top, c.rect.left, c.rect.bottom, c.rect.right};
      ++k;
    }
  }
  r->char_offsets[n] = (int64_t)k;
  return r;
}

std::vector<geom::RotatedRect> to_rects(const ocrs_b200_rotated_rect* p, size_t n) {
  std::vector<geom::RotatedRect> v(n);
  static_assert(sizeof(geom::RotatedRect) == sizeof(ocrs_b200_rotated_rect), "layout mismatch");
  if (n) std::memcpy(v.data(), p, n * sizeof(geom::RotatedRect));
  return v;
}
}  // namespace

namespace {
char* join_text(const std::vector<TextLine>& lines) {  // OcrEngine::get_text (lib.rs:290-300)
  std::string text;
  bool first = true;
  for (const auto& l : lines) {
    if (!l.present) continue;
    if (!first) text.push_back('\n');
    first = false;
    std::vector<uint32_t> cps;
    cps.reserve(l.chars.size());
    for (const auto& c : l.chars) cps.push_back(c.ch);
    text += codepoints_to_utf8(cps);
  }
  char* out = cmalloc<char>(text.size() + 1);
  std::memcpy(out, text.c_str(), text.size() + 1);
  return out;
}

EngineParams to_engine_params(const ocrs_b200_engine_params* p) {
  EngineParams ep;
  ep.detection_model = p->detection_model;
  ep.detection_model_len = p->detection_model_len;
  ep.recognition_model = p->recognition_model;
  ep.recognition_model_len = p->recognition_model_len;
  ep.debug = p->debug != 0;
  OCRS_CHECK(p->decode_method == OCRS_B200_DECODE_GREEDY || p->decode_method == OCRS_B200_DECODE_BEAM, kInvalidArg,
             "unknown decode method");
  ep.decode_method = p->decode_method == OCRS_B200_DECODE_BEAM ? DecodeMethod::kBeamSearch : DecodeMethod::kGreedy;
  ep.beam_width = p->beam_width;
  if (p->alphabet_utf8) { ep.has_alphabet = true; ep.alphabet_utf8 = p->alphabet_utf8; }
  if (p->allowed_chars_utf8) { ep.has_allowed_chars = true; ep.allowed_chars_utf8 = p->allowed_chars_utf8; }
  ep.device = p->device;
  return ep;
}
}  // namespace

struct ocrs_b200_pool {
  std::unique_ptr<Pool> pool;
};

struct ocrs_b200_model {
  std::unique_ptr<Model> model;
  double last_flops = 0;
  std::mutex mu;
};
struct ocrs_b200_engine {
  std::shared_ptr<Engine> engine;
};
// An input keeps its engine (stream, memory pool) alive: handles may be destroyed in any order.
struct ocrs_b200_input {
  std::unique_ptr<OcrInput> input;
  std::shared_ptr<Engine> owner;
  ~ocrs_b200_input() { input.reset(); }
};

extern "C" {

const char* ocrs_b200_last_error(void) { return g_last_error.c_str(); }
const char* ocrs_b200_version(void) { return "ocrs_b200 0.1.0 (sm_100a; reference ocrs 0.12.2 @ 4bccf6b)"; }

int ocrs_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void ocrs_b200_free(void* p) { std::free(p); }

int ocrs_b200_model_load(const uint8_t* bytes, size_t len, int device, ocrs_b200_model** out) {
  return guard([&] {
    OCRS_CHECK(out != nullptr && bytes != nullptr, kInvalidArg, "null argument");
    *out = nullptr;
    require_device();
    auto* m = new ocrs_b200_model();
    try {
      m->model = Model::load(bytes, len, device);
    } catch (...) {
      delete m;
      throw;
    }
    *out = m;
  });
}

int ocrs_b200_model_load_file(const char* path, int device, ocrs_b200_model** out) {
  return guard([&] {
    OCRS_CHECK(path != nullptr && out != nullptr, kInvalidArg, "null argument");
    std::ifstream f(path, std::ios::binary);
    OCRS_CHECK(f.good(), kModelLoad, std::string("cannot open model file: ") + path);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    int rc = ocrs_b200_model_load(buf.data(), buf.size(), device, out);
    if (rc != 0) throw Error(rc, g_last_error);
  });
}

int ocrs_b200_model_input_shape(const ocrs_b200_model* m, int64_t dims[8], int* ndim) {
  return guard([&] {
    OCRS_CHECK(m && dims && ndim, kInvalidArg, "null argument");
    const auto& s = m->model->input_shape();
    OCRS_CHECK(s.size() <= 8, kInternal, "input rank > 8");
    *ndim = (int)s.size();
    for (size_t i = 0; i < s.size(); ++i) dims[i] = s[i];
  });
}

int ocrs_b200_model_inspect(const uint8_t* bytes, size_t len, char** json) {
  return guard([&] {
    OCRS_CHECK(bytes && json, kInvalidArg, "null argument");
    OCRS_CHECK(!onnx::looks_like_rten(bytes, len), kModelLoad,
               ".rten containers are not supported (FlatBuffers schema unavailable offline); export the model as .onnx");
    onnx::Graph g = onnx::parse_model(bytes, len);
    // JSON string; names come from an untrusted file: bytes that are not well-formed UTF-8 become U+FFFD
    auto esc = [](const std::string& in) {
      std::string o = "\"";
      const size_t n = in.size();
      for (size_t i = 0; i < n;) {
        const unsigned char ch = (unsigned char)in[i];
        if (ch < 0x80) {
          if (ch == '"' || ch == '\\') { o.push_back('\\'); o.push_back((char)ch); }
          else if (ch < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
          else o.push_back((char)ch);
          ++i;
          continue;
        }
        int len = ch >= 0xF0 && ch <= 0xF4 ? 4 : ch >= 0xE0 ? 3 : ch >= 0xC2 && ch < 0xE0 ? 2 : 0;
        bool ok = len > 0 && i + (size_t)len <= n;
        for (int k = 1; ok && k < len; ++k) ok = ((unsigned char)in[i + k] & 0xC0) == 0x80;
        if (ok && len == 3) {
          const unsigned char c1 = (unsigned char)in[i + 1];
          ok = !(ch == 0xE0 && c1 < 0xA0) && !(ch == 0xED && c1 >= 0xA0);  // overlong, surrogates
        }
        if (ok && len == 4) {
          const unsigned char c1 = (unsigned char)in[i + 1];
          ok = !(ch == 0xF0 && c1 < 0x90) && !(ch == 0xF4 && c1 >= 0x90);
        }
        if (ok) {
          o.append(in, i, (size_t)len);
          i += (size_t)len;
        } else {
          o += "\\ufffd";
          ++i;
        }
      }
      return o + "\"";
    };
    auto values = [&](const std::vector<onnx::ValueInfo>& vs) {
      std::string o = "[";
      for (size_t i = 0; i < vs.size(); ++i) {
        o += std::string(i ? "," : "") + "{\"name\":" + esc(vs[i].name) + ",\"dims\":[";
        for (size_t d = 0; d < vs[i].dims.size(); ++d) o += std::string(d ? "," : "") + std::to_string(vs[i].dims[d]);
        o += "]}";
      }
      return o + "]";
    };
    std::map<std::string, int> ops;
    for (const auto& n : g.nodes) ops[n.op] += 1;
    size_t init_bytes = 0;
    for (const auto& kv : g.initializers) init_bytes += kv.second.raw.size();
    std::string o = "{\"format\":\"onnx\",\"opset\":" + std::to_string(g.opset) + ",\"inputs\":" + values(g.inputs) +
                    ",\"outputs\":" + values(g.outputs) + ",\"nodes\":" + std::to_string(g.nodes.size()) + ",\"ops\":{";
    bool first = true;
    for (const auto& kv : ops) {
      o += std::string(first ? "" : ",") + esc(kv.first) + ":" + std::to_string(kv.second);
      first = false;
    }
    o += "},\"unsupported_ops\":[";
    first = true;
    for (const auto& kv : ops)
      if (!supported_ops().count(kv.first)) {
        o += std::string(first ? "" : ",") + esc(kv.first);
        first = false;
      }
    o += "],\"initializers\":" + std::to_string(g.initializers.size()) + ",\"initializer_bytes\":" +
         std::to_string(init_bytes) + "}";
    char* out = cmalloc<char>(o.size() + 1);
    std::memcpy(out, o.data(), o.size());
    out[o.size()] = 0;
    *json = out;
  });
}

int ocrs_b200_model_run(const ocrs_b200_model* cm, const float* in, const int64_t* in_shape, int in_ndim, float** out,
                        int64_t out_shape[8], int* out_ndim) {
  return guard([&] {
    OCRS_CHECK(cm && in && in_shape && out && out_shape && out_ndim, kInvalidArg, "null argument");
    OCRS_CHECK(in_ndim > 0 && in_ndim <= 8, kInvalidArg, "bad input rank");
    auto* m = const_cast<ocrs_b200_model*>(cm);
    OCRS_CUDA_CHECK(cudaSetDevice(m->model->device()));
    // one private stream per call: re-entrant on a shared handle
    cudaStream_t st;
    OCRS_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    try {
      std::vector<int64_t> shape(in_shape, in_shape + in_ndim);
      int64_t n = 1;
      for (auto d : shape) {
        OCRS_CHECK(d >= 0, kInvalidArg, "negative dimension");
        n *= d;
      }
      DTensor x;
      x.shape = shape;
      x.storage = std::make_shared<Storage>((size_t)n * 4, st);
      x.data = reinterpret_cast<float*>(x.storage->ptr);
      OCRS_CUDA_CHECK(cudaMemcpyAsync(x.data, in, (size_t)n * 4, cudaMemcpyHostToDevice, st));
      ModelCost cost;
      const int tc_tok = m->model->tc_token();
      DTensor y = m->model->run(x, st, &cost);
      OCRS_CHECK(y.shape.size() <= 8, kWrongOutput, "output rank > 8");
      float* host = cmalloc<float>((size_t)y.numel());
      cudaError_t e = cudaMemcpyAsync(host, y.data, (size_t)y.numel() * 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e == cudaSuccess && m->model->take_tc_overflow(tc_tok)) {
        // split-fp16 range overflow: the model now runs its convolutions in fp32; repeat the call
        y = m->model->run(x, st, &cost);
        e = cudaMemcpyAsync(host, y.data, (size_t)y.numel() * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      }
      if (e != cudaSuccess) {
        std::free(host);
        throw Error(kCuda, std::string("CUDA error ") + cudaGetErrorString(e));
      }
      *out = host;
      *out_ndim = (int)y.shape.size();
      for (size_t i = 0; i < y.shape.size(); ++i) out_shape[i] = y.shape[i];
      {
        std::lock_guard<std::mutex> lk(m->mu);
        m->last_flops = cost.flops;
      }
      x = DTensor();
      y = DTensor();
      cudaStreamSynchronize(st);
    } catch (...) {
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
      throw;
    }
    cudaStreamDestroy(st);
  });
}

double ocrs_b200_model_last_flops(const ocrs_b200_model* m) { return m ? m->last_flops : 0.0; }

void ocrs_b200_model_destroy(ocrs_b200_model* m) {
  guard([&] { delete m; });
}

int ocrs_b200_engine_create(const ocrs_b200_engine_params* p, ocrs_b200_engine** out) {
  return guard([&] {
    OCRS_CHECK(p && out, kInvalidArg, "null argument");
    *out = nullptr;
    require_device();
    EngineParams ep = to_engine_params(p);
    auto* e = new ocrs_b200_engine();
    try {
      e->engine = std::make_shared<Engine>(ep);
    } catch (...) {
      delete e;
      throw;
    }
    *out = e;
  });
}

void ocrs_b200_engine_destroy(ocrs_b200_engine* e) {
  guard([&] { delete e; });
}

int ocrs_b200_engine_prepare_input_bytes(ocrs_b200_engine* e, const uint8_t* bytes, size_t len, uint32_t width,
                                         uint32_t height, ocrs_b200_input** out) {
  return guard([&] {
    OCRS_CHECK(e && out, kInvalidArg, "null argument");
    *out = nullptr;
    // ImageSource::from_bytes (preprocess.rs:81-102)
    size_t channel_len = (size_t)width * height;
    OCRS_CHECK(channel_len != 0, kUnsupportedChannelCount, "channel count is not 1, 3 or 4");
    OCRS_CHECK(len % channel_len == 0, kInvalidDataLength, "data length is not a multiple of `width * height`");
    size_t channels = len / channel_len;
    OCRS_CHECK(channels == 1 || channels == 3 || channels == 4, kUnsupportedChannelCount,
               "channel count is not 1, 3 or 4");
    auto in = e->engine->prepare_input(bytes, 0, 0, (int)height, (int)width, (int)channels);
    auto* h = new ocrs_b200_input();
    h->input = std::move(in);
    h->owner = e->engine;
    *out = h;
  });
}

static int prepare_input_common(ocrs_b200_engine* e, const void* pixels, int dtype, int order, int height, int width,
                                int channels, bool on_device, ocrs_b200_input** out) {
  return guard([&] {
    OCRS_CHECK(e && out, kInvalidArg, "null argument");
    *out = nullptr;
    auto in = e->engine->prepare_input(pixels, dtype, order, height, width, channels, on_device);
    auto* h = new ocrs_b200_input();
    h->input = std::move(in);
    h->owner = e->engine;
    *out = h;
  });
}

int ocrs_b200_engine_prepare_input(ocrs_b200_engine* e, const void* pixels, int dtype, int order, int height,
                                   int width, int channels, ocrs_b200_input** out) {
  return prepare_input_common(e, pixels, dtype, order, height, width, channels, false, out);
}
int ocrs_b200_engine_prepare_input_device(ocrs_b200_engine* e, const void* pixels, int dtype, int order, int height,
                                          int width, int channels, ocrs_b200_input** out) {
  return prepare_input_common(e, pixels, dtype, order, height, width, channels, true, out);
}

int ocrs_b200_input_shape(const ocrs_b200_input* in, int* height, int* width) {
  return guard([&] {
    OCRS_CHECK(in && height && width, kInvalidArg, "null argument");
    *height = in->input->H;
    *width = in->input->W;
  });
}

int ocrs_b200_input_read(ocrs_b200_engine* e, const ocrs_b200_input* in, float* out) {
  return guard([&] {
    OCRS_CHECK(e && in && out, kInvalidArg, "null argument");
    e->engine->synchronize();
    OCRS_CUDA_CHECK(cudaMemcpy(out, in->input->grey(), (size_t)in->input->H * in->input->W * 4, cudaMemcpyDeviceToHost));
  });
}

void ocrs_b200_input_destroy(ocrs_b200_input* in) {
  guard([&] { delete in; });
}

int ocrs_b200_engine_detect_text_pixels(ocrs_b200_engine* e, const ocrs_b200_input* in, float* out) {
  return guard([&] {
    OCRS_CHECK(e && in && out, kInvalidArg, "null argument");
    std::vector<float> p = e->engine->detect_text_pixels(*in->input);
    std::memcpy(out, p.data(), p.size() * 4);
  });
}

int ocrs_b200_engine_detect_words(ocrs_b200_engine* e, const ocrs_b200_input* in, ocrs_b200_rotated_rect** rects,
                                  size_t* n) {
  return guard([&] {
    OCRS_CHECK(e && in && rects && n, kInvalidArg, "null argument");
    auto r = e->engine->detect_words({in->input.get()});
    *n = r[0].size();
    *rects = cmalloc<ocrs_b200_rotated_rect>(*n);
    if (*n) std::memcpy(*rects, r[0].data(), *n * sizeof(ocrs_b200_rotated_rect));
  });
}

int ocrs_b200_engine_detect_words_batch(ocrs_b200_engine* e, const ocrs_b200_input* const* inputs, size_t n_pages,
                                        ocrs_b200_rotated_rect** rects, size_t** offsets) {
  return guard([&] {
    OCRS_CHECK(e && (inputs || n_pages == 0) && rects && offsets, kInvalidArg, "null argument");
    std::vector<const OcrInput*> pages(n_pages);
    for (size_t i = 0; i < n_pages; ++i) {
      OCRS_CHECK(inputs[i] != nullptr, kInvalidArg, "null input");
      pages[i] = inputs[i]->input.get();
    }
    auto r = e->engine->detect_words(pages);
    size_t total = 0;
    for (auto& v : r) total += v.size();
    *rects = cmalloc<ocrs_b200_rotated_rect>(total);
    *offsets = cmalloc<size_t>(n_pages + 1);
    size_t k = 0;
    for (size_t i = 0; i < n_pages; ++i) {
      (*offsets)[i] = k;
      if (!r[i].empty()) std::memcpy(*rects + k, r[i].data(), r[i].size() * sizeof(ocrs_b200_rotated_rect));
      k += r[i].size();
    }
    (*offsets)[n_pages] = k;
  });
}

int ocrs_b200_find_text_lines(const ocrs_b200_rotated_rect* words, size_t n_words, ocrs_b200_rotated_rect** out_words,
                              size_t** line_offsets, size_t* n_lines) {
  return guard([&] {
    OCRS_CHECK((words || n_words == 0) && out_words && line_offsets && n_lines, kInvalidArg, "null argument");
    auto lines = layout::find_text_lines(to_rects(words, n_words));
    size_t total = 0;
    for (auto& l : lines) total += l.size();
    *out_words = cmalloc<ocrs_b200_rotated_rect>(total);
    *line_offsets = cmalloc<size_t>(lines.size() + 1);
    size_t k = 0;
    for (size_t i = 0; i < lines.size(); ++i) {
      (*line_offsets)[i] = k;
      std::memcpy(*out_words + k, lines[i].data(), lines[i].size() * sizeof(ocrs_b200_rotated_rect));
      k += lines[i].size();
    }
    (*line_offsets)[lines.size()] = k;
    *n_lines = lines.size();
  });
}

int ocrs_b200_engine_find_text_lines(ocrs_b200_engine* e, const ocrs_b200_input* in,
                                     const ocrs_b200_rotated_rect* words, size_t n_words,
                                     ocrs_b200_rotated_rect** out_words, size_t** line_offsets, size_t* n_lines) {
  (void)in;  // unused by the reference as well (lib.rs:224)
  if (!e) {
    g_last_error = "null argument";
    return OCRS_B200_ERR_INVALID_ARG;
  }
  return ocrs_b200_find_text_lines(words, n_words, out_words, line_offsets, n_lines);
}

namespace {
std::vector<TextLine> from_text_result(const ocrs_b200_text_result* r) {
  OCRS_CHECK(r && r->n_lines >= 0 && (r->n_lines == 0 || (r->line_present && r->char_offsets)), kInvalidArg, "null argument");
  std::vector<TextLine> lines((size_t)r->n_lines);
  for (int32_t i = 0; i < r->n_lines; ++i) {
    lines[i].present = r->line_present[i] != 0;
    if (!lines[i].present) continue;
    OCRS_CHECK(r->char_offsets[i + 1] > r->char_offsets[i], kInvalidArg, "Text lines must not be empty");  // text_items.rs:71
    for (int64_t k = r->char_offsets[i]; k < r->char_offsets[i + 1]; ++k) {
      const ocrs_b200_rect& cr = r->char_rects[k];
      lines[i].chars.push_back(TextChar{r->chars[k], geom::RectI{cr.top, cr.left, cr.bottom, cr.right}});
    }
  }
  return lines;
}
char* dup_string(const std::string& s) {
  char* out = cmalloc<char>(s.size() + 1);
  std::memcpy(out, s.data(), s.size());
  out[s.size()] = 0;
  return out;
}
}  // namespace

int ocrs_b200_text_item_rotated_rect(const ocrs_b200_rect* char_rects, size_t n, ocrs_b200_rotated_rect* out) {
  return guard([&] {
    OCRS_CHECK(char_rects && out && n > 0, kInvalidArg, "null argument or empty item");
    std::vector<geom::RectI> rects(n);
    for (size_t i = 0; i < n; ++i) rects[i] = geom::RectI{char_rects[i].top, char_rects[i].left, char_rects[i].bottom, char_rects[i].right};
    geom::RotatedRect rr = textout::item_rotated_rect(rects.data(), n);
    std::memcpy(out, &rr, sizeof rr);
  });
}

int ocrs_b200_rotated_rect_vertices(const ocrs_b200_rotated_rect* r, int32_t xy[8]) {
  return guard([&] {
    OCRS_CHECK(r && xy, kInvalidArg, "null argument");
    geom::RotatedRect rr;
    std::memcpy(&rr, r, sizeof rr);
    textout::rounded_vertices(rr, xy);
  });
}

int ocrs_b200_format_text_output(const ocrs_b200_text_result* lines, char** utf8) {
  return guard([&] {
    OCRS_CHECK(utf8, kInvalidArg, "null argument");
    *utf8 = dup_string(textout::format_text(from_text_result(lines)));
  });
}

int ocrs_b200_format_json_output(const ocrs_b200_text_result* lines, const char* input_path, int image_height,
                                 int image_width, char** utf8) {
  return guard([&] {
    OCRS_CHECK(utf8 && input_path, kInvalidArg, "null argument");
    *utf8 = dup_string(textout::format_json(from_text_result(lines), input_path, image_height, image_width));
  });
}

int ocrs_b200_engine_recognize_text(ocrs_b200_engine* e, const ocrs_b200_input* in,
                                    const ocrs_b200_rotated_rect* words, const size_t* line_offsets, size_t n_lines,
                                    ocrs_b200_text_result** out) {
  return guard([&] {
    OCRS_CHECK(e && in && out && (n_lines == 0 || (words && line_offsets)), kInvalidArg, "null argument");
    std::vector<std::vector<geom::RotatedRect>> lines(n_lines);
    for (size_t i = 0; i < n_lines; ++i) {
      OCRS_CHECK(line_offsets[i + 1] >= line_offsets[i], kInvalidArg, "line offsets must be non-decreasing");
      lines[i] = to_rects(words + line_offsets[i], line_offsets[i + 1] - line_offsets[i]);
    }
    auto r = e->engine->recognize_text({in->input.get()}, {lines});
    *out = make_text_result(r[0]);
  });
}

void ocrs_b200_text_result_free(ocrs_b200_text_result* r) {
  if (!r) return;
  std::free(r->line_present);
  std::free(r->char_offsets);
  std::free(r->chars);
  std::free(r->char_rects);
  std::free(r);
}

int ocrs_b200_engine_prepare_recognition_input(ocrs_b200_engine* e, const ocrs_b200_input* in,
                                               const ocrs_b200_rotated_rect* line_words, size_t n_words, float** out,
                                               int* out_h, int* out_w) {
  return guard([&] {
    OCRS_CHECK(e && in && line_words && out && out_h && out_w, kInvalidArg, "null argument");
    std::vector<float> img = e->engine->prepare_recognition_input(*in->input, to_rects(line_words, n_words), out_h, out_w);
    *out = cmalloc<float>(img.size());
    std::memcpy(*out, img.data(), img.size() * 4);
  });
}

float ocrs_b200_engine_detection_threshold(const ocrs_b200_engine* e) {
  return e ? e->engine->detection_threshold() : 0.2f;  // TextDetectorParams::default (detection.rs:34)
}

int ocrs_b200_engine_get_text(ocrs_b200_engine* e, const ocrs_b200_input* in, char** utf8) {
  return guard([&] {
    OCRS_CHECK(e && in && utf8, kInvalidArg, "null argument");
    auto r = e->engine->ocr_pages({in->input.get()});
    std::string text;
    bool first = true;
    for (const auto& l : r[0]) {
      if (!l.present) continue;
      if (!first) text.push_back('\n');
      first = false;
      std::vector<uint32_t> cps;
      for (const auto& c : l.chars) cps.push_back(c.ch);
      text += codepoints_to_utf8(cps);
    }
    *utf8 = cmalloc<char>(text.size() + 1);
    std::memcpy(*utf8, text.c_str(), text.size() + 1);
  });
}

int ocrs_b200_engine_ocr_batch(ocrs_b200_engine* e, const ocrs_b200_input* const* inputs, size_t n_pages,
                               ocrs_b200_text_result** results) {
  return guard([&] {
    OCRS_CHECK(e && (inputs || n_pages == 0) && results, kInvalidArg, "null argument");
    std::vector<const OcrInput*> pages(n_pages);
    for (size_t i = 0; i < n_pages; ++i) {
      OCRS_CHECK(inputs[i] != nullptr, kInvalidArg, "null input");
      pages[i] = inputs[i]->input.get();
      results[i] = nullptr;
    }
    auto r = e->engine->ocr_pages(pages);
    for (size_t i = 0; i < n_pages; ++i) results[i] = make_text_result(r[i]);
  });
}

int ocrs_b200_engine_ocr_batch_text(ocrs_b200_engine* e, const ocrs_b200_input* const* inputs, size_t n_pages,
                                    char** texts) {
  return guard([&] {
    OCRS_CHECK(e && (inputs || n_pages == 0) && texts, kInvalidArg, "null argument");
    std::vector<const OcrInput*> pages(n_pages);
    for (size_t i = 0; i < n_pages; ++i) {
      OCRS_CHECK(inputs[i] != nullptr, kInvalidArg, "null input");
      pages[i] = inputs[i]->input.get();
      texts[i] = nullptr;
    }
    auto r = e->engine->ocr_pages(pages);
    for (size_t i = 0; i < n_pages; ++i) texts[i] = join_text(r[i]);
  });
}

// ---- batched debug outputs -------------------------------------------------------------------------
int ocrs_b200_engine_detect_text_pixels_batch(ocrs_b200_engine* e, const ocrs_b200_input* const* inputs, size_t n_pages,
                                              float** maps, uint8_t** masks) {
  return guard([&] {
    OCRS_CHECK(e && (inputs || n_pages == 0), kInvalidArg, "null argument");
    std::vector<const OcrInput*> pages(n_pages);
    for (size_t i = 0; i < n_pages; ++i) {
      OCRS_CHECK(inputs[i] != nullptr, kInvalidArg, "null input");
      pages[i] = inputs[i]->input.get();
      if (maps) maps[i] = nullptr;
      if (masks) masks[i] = nullptr;
    }
    auto r = e->engine->detect_text_pixels_batch(pages, maps != nullptr, masks != nullptr);
    for (size_t i = 0; i < n_pages; ++i) {
      if (maps) {
        maps[i] = cmalloc<float>(r[i].map.size());
        std::memcpy(maps[i], r[i].map.data(), r[i].map.size() * 4);
      }
      if (masks) {
        masks[i] = cmalloc<uint8_t>(r[i].mask.size());
        std::memcpy(masks[i], r[i].mask.data(), r[i].mask.size());
      }
    }
  });
}

int ocrs_b200_engine_prepare_recognition_inputs(ocrs_b200_engine* e, const ocrs_b200_input* in,
                                                const ocrs_b200_rotated_rect* words, const size_t* line_offsets,
                                                size_t n_lines, float** images, int* height, int** widths,
                                                size_t** offsets) {
  return guard([&] {
    OCRS_CHECK(e && in && (line_offsets || n_lines == 0) && images && height && widths && offsets, kInvalidArg, "null argument");
    std::vector<std::vector<geom::RotatedRect>> lines(n_lines);
    for (size_t i = 0; i < n_lines; ++i) {
      OCRS_CHECK(line_offsets[i + 1] >= line_offsets[i], kInvalidArg, "line offsets must be non-decreasing");
      OCRS_CHECK(words != nullptr || line_offsets[i + 1] == line_offsets[i], kInvalidArg, "words is null");
      lines[i] = to_rects(words + line_offsets[i], line_offsets[i + 1] - line_offsets[i]);
    }
    auto r = e->engine->prepare_recognition_inputs(*in->input, lines);
    *height = r.height;
    *images = cmalloc<float>(r.images.size());
    std::memcpy(*images, r.images.data(), r.images.size() * 4);
    *widths = cmalloc<int>(n_lines);
    *offsets = cmalloc<size_t>(n_lines);
    for (size_t i = 0; i < n_lines; ++i) {
      (*widths)[i] = r.widths[i];
      (*offsets)[i] = r.offsets[i];
    }
  });
}

// ---- engine pool ---------------------------------------------------------------------------------
int ocrs_b200_pool_create(const ocrs_b200_pool_params* p, ocrs_b200_pool** out) {
  return guard([&] {
    OCRS_CHECK(p && out, kInvalidArg, "null argument");
    *out = nullptr;
    require_device();
    PoolParams pp;
    pp.engine = to_engine_params(&p->engine);
    OCRS_CHECK(p->n_devices >= 0 && (p->device_ids != nullptr || p->n_devices == 0), kInvalidArg, "bad device list");
    for (int i = 0; i < p->n_devices; ++i) pp.devices.push_back(p->device_ids[i]);
    if (p->in_flight != 0) pp.in_flight = p->in_flight;
    pp.pin_numa = p->pin_numa >= 0;
    if (p->layout_threads != 0) pp.layout_threads = p->layout_threads;
    auto* h = new ocrs_b200_pool();
    try {
      h->pool = std::make_unique<Pool>(pp);
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
  });
}

void ocrs_b200_pool_destroy(ocrs_b200_pool* p) {
  guard([&] { delete p; });
}

int ocrs_b200_pool_submit(ocrs_b200_pool* p, const ocrs_b200_page* pages, size_t n_pages, uint64_t* ticket) {
  return guard([&] {
    OCRS_CHECK(p && ticket && (pages || n_pages == 0), kInvalidArg, "null argument");
    std::vector<PoolPage> v(n_pages);
    for (size_t i = 0; i < n_pages; ++i) {
      const ocrs_b200_page& s = pages[i];
      // ImageSource validation (preprocess.rs:81-123) happens up front so that a bad page fails the submit
      OCRS_CHECK(s.pixels != nullptr, kInvalidArg, "pixels is null");
      OCRS_CHECK(s.dtype == 0 || s.dtype == 1, kInvalidArg, "dtype must be 0 (u8) or 1 (f32)");
      OCRS_CHECK(s.order == 0 || s.order == 1, kInvalidArg, "order must be 0 (HWC) or 1 (CHW)");
      OCRS_CHECK(s.height >= 0 && s.width >= 0, kInvalidArg, "negative image size");
      OCRS_CHECK(s.channels == 1 || s.channels == 3 || s.channels == 4, kUnsupportedChannelCount,
                 "channel count is not 1, 3 or 4");
      v[i].pixels = s.pixels; v[i].dtype = s.dtype; v[i].order = s.order;
      v[i].H = s.height; v[i].W = s.width; v[i].C = s.channels; v[i].on_device = s.on_device != 0;
    }
    *ticket = p->pool->submit(v.data(), v.size());
  });
}

int ocrs_b200_pool_wait(ocrs_b200_pool* p, uint64_t ticket, ocrs_b200_text_result** results, size_t n_pages) {
  return guard([&] {
    OCRS_CHECK(p && (results || n_pages == 0), kInvalidArg, "null argument");
    for (size_t i = 0; i < n_pages; ++i) results[i] = nullptr;
    auto r = p->pool->wait(ticket);
    OCRS_CHECK(r.size() == n_pages, kInvalidArg, "n_pages does not match the submitted batch");
    for (size_t i = 0; i < n_pages; ++i) results[i] = make_text_result(r[i]);
  });
}

int ocrs_b200_pool_wait_text(ocrs_b200_pool* p, uint64_t ticket, char** texts, size_t n_pages) {
  return guard([&] {
    OCRS_CHECK(p && (texts || n_pages == 0), kInvalidArg, "null argument");
    for (size_t i = 0; i < n_pages; ++i) texts[i] = nullptr;
    auto r = p->pool->wait(ticket);
    OCRS_CHECK(r.size() == n_pages, kInvalidArg, "n_pages does not match the submitted batch");
    for (size_t i = 0; i < n_pages; ++i) texts[i] = join_text(r[i]);
  });
}

int ocrs_b200_pool_done(ocrs_b200_pool* p, uint64_t ticket, int* done) {
  return guard([&] {
    OCRS_CHECK(p && done, kInvalidArg, "null argument");
    *done = p->pool->done(ticket) ? 1 : 0;
  });
}

int ocrs_b200_pool_shape(const ocrs_b200_pool* p, int* n_devices, int* in_flight) {
  return guard([&] {
    OCRS_CHECK(p && n_devices && in_flight, kInvalidArg, "null argument");
    *n_devices = p->pool->n_devices();
    *in_flight = p->pool->in_flight();
  });
}

int ocrs_b200_pool_engine(ocrs_b200_pool* p, int dev_index, int slot, ocrs_b200_engine** out) {
  return guard([&] {
    OCRS_CHECK(p && out, kInvalidArg, "null argument");
    *out = nullptr;
    auto eng = p->pool->engine(dev_index, slot);
    auto* e = new ocrs_b200_engine();
    e->engine = std::move(eng);
    *out = e;
  });
}

int ocrs_b200_pool_describe(ocrs_b200_pool* p, char** text) {
  return guard([&] {
    OCRS_CHECK(p && text, kInvalidArg, "null argument");
    std::string s = p->pool->numa_report();
    *text = cmalloc<char>(s.size() + 1);
    std::memcpy(*text, s.c_str(), s.size() + 1);
  });
}

int ocrs_b200_engine_stats(ocrs_b200_engine* e, double out[8], int reset) {
  return guard([&] {
    OCRS_CHECK(e && out, kInvalidArg, "null argument");
    auto s = e->engine->stats();
    out[0] = s.det_flops;
    out[1] = s.rec_flops;
    out[2] = (double)s.n_words;
    out[3] = (double)s.n_lines;
    out[4] = (double)s.n_timesteps;
    out[5] = (double)s.rec_batches;
    out[6] = 0;
    out[7] = 0;
    if (reset) e->engine->reset_stats();
  });
}

int ocrs_b200_engine_set_profiling(ocrs_b200_engine* e, int enable) {
  return guard([&] {
    OCRS_CHECK(e, kInvalidArg, "null argument");
    e->engine->set_profiling(enable != 0);
  });
}

int ocrs_b200_engine_profile_json(ocrs_b200_engine* e, char** json, int reset) {
  return guard([&] {
    OCRS_CHECK(e && json, kInvalidArg, "null argument");
    std::string j = e->engine->profile_json(reset != 0);
    *json = cmalloc<char>(j.size() + 1);
    std::memcpy(*json, j.c_str(), j.size() + 1);
  });
}

int ocrs_b200_engine_timer_start(ocrs_b200_engine* e) {
  return guard([&] {
    OCRS_CHECK(e, kInvalidArg, "null argument");
    e->engine->timer_start();
  });
}

int ocrs_b200_engine_timer_stop(ocrs_b200_engine* e, float* ms) {
  return guard([&] {
    OCRS_CHECK(e && ms, kInvalidArg, "null argument");
    *ms = e->engine->timer_stop();
  });
}

int ocrs_b200_engine_transfer_bytes(ocrs_b200_engine* e, int64_t out[2]) {
  return guard([&] {
    OCRS_CHECK(e && out, kInvalidArg, "null argument");
    out[0] = e->engine->h2d_bytes();
    out[1] = e->engine->d2h_bytes();
  });
}

int64_t ocrs_b200_kernel_launch_count(void) { return g_kernel_launches.load(); }

}  // extern "C"
