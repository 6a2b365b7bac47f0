// OcrEngine orchestration on one GPU.  See engine.h.
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <exception>
#include <map>
#include <deque>
#include <thread>

#include "layout.h"

namespace ocrs {

using geom::PointI;
using geom::RectI;
using geom::RotatedRect;

// lib.rs:34 (the character before "ABCDE" is a plain 'E' in the reference source at this commit)
static const char kDefaultAlphabet[] =
    " 0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~EABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz";

std::vector<uint32_t> utf8_to_codepoints(const std::string& s) {
  std::vector<uint32_t> out;
  size_t i = 0, n = s.size();
  while (i < n) {
    uint8_t c = (uint8_t)s[i];
    uint32_t cp;
    int extra;
    if (c < 0x80) { cp = c; extra = 0; }
    else if ((c >> 5) == 0x6) { cp = c & 0x1F; extra = 1; }
    else if ((c >> 4) == 0xE) { cp = c & 0x0F; extra = 2; }
    else if ((c >> 3) == 0x1E) { cp = c & 0x07; extra = 3; }
    else { cp = 0xFFFD; extra = 0; }
    ++i;
    for (int k = 0; k < extra && i < n; ++k, ++i) cp = (cp << 6) | ((uint8_t)s[i] & 0x3F);
    out.push_back(cp);
  }
  return out;
}

std::string codepoints_to_utf8(const std::vector<uint32_t>& cps) {
  std::string s;
  for (uint32_t cp : cps) {
    if (cp < 0x80) s.push_back((char)cp);
    else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {
      s.push_back((char)(0xE0 | (cp >> 12)));
      s.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      s.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      s.push_back((char)(0xF0 | (cp >> 18)));
      s.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      s.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      s.push_back((char)(0x80 | (cp & 0x3F)));
    }
  }
  return s;
}

// RAII host-section timer feeding profile_json ("host/<name>")
struct Engine::HostTimer {
  Engine* e;
  const char* name;
  std::chrono::steady_clock::time_point t0;
  HostTimer(Engine* eng, const char* n) : e(eng), name(n), t0(std::chrono::steady_clock::now()) {}
  ~HostTimer() {
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::lock_guard<std::mutex> lk(e->host_mu_);  // timers run inside and outside mu_ (ocr_pages): own lock
    auto& slot = e->host_ms_[name];
    slot.first += ms;
    slot.second += 1;
  }
};

struct Engine::PageScratch {
  int cap_hw = 0;
  DeviceBuffer mask, bits, wstart, prob, labels, comp_roots, counters, pts, simp_idx, stack, fpts, hull, rects, rect_root;
  img::ComponentBuffers bufs{};
};

Engine::PageScratch& Engine::scratch_for(int slot, int H, int W) {
  while ((int)scratch_.size() <= slot) scratch_.push_back(std::make_unique<PageScratch>());
  PageScratch& s = *scratch_[slot];
  int64_t hw = (int64_t)H * W;
  if (hw > s.cap_hw || s.bufs.labels == nullptr) {
    int64_t pool = 2 * hw + 16;
    int32_t max_comps = (int32_t)(((int64_t)(H + 1) / 2) * ((W + 1) / 2) + 1);
    s.mask.reserve((size_t)hw);
    s.bits.reserve((size_t)H * ((W + 31) / 32) * 4 + 4);
    s.wstart.reserve((size_t)H * ((W + 31) / 32) * 2 + 4);
    s.labels.reserve((size_t)(hw + 1) * 4);
    s.comp_roots.reserve((size_t)max_comps * 4);
    s.counters.reserve(8 * 4);
    s.pts.reserve((size_t)pool * 2 * sizeof(int16_t));
    s.simp_idx.reserve((size_t)pool * 4);
    s.stack.reserve((size_t)pool * 3 * 4);
    s.fpts.reserve((size_t)pool * 2 * 4);
    s.hull.reserve((size_t)pool * 2 * 4);
    s.rects.reserve((size_t)max_comps * sizeof(RotatedRect));
    s.rect_root.reserve((size_t)max_comps * 4);
    s.cap_hw = (int)hw;
    s.bufs.labels = s.labels.as<int32_t>();
    s.bufs.comp_roots = s.comp_roots.as<int32_t>();
    s.bufs.counters = s.counters.as<int32_t>();
    s.bufs.pts = s.pts.as<int16_t>();
    s.bufs.simp_idx = s.simp_idx.as<int32_t>();
    s.bufs.stack = s.stack.as<int32_t>();
    s.bufs.fpts = s.fpts.as<float>();
    s.bufs.hull = s.hull.as<float>();
    s.bufs.rects = s.rects.as<RotatedRect>();
    s.bufs.rect_root = s.rect_root.as<int32_t>();
    s.bufs.pool_cap = pool;
    s.bufs.max_comps = max_comps;
  }
  return s;
}

Engine::Engine(const EngineParams& p) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  OCRS_CHECK(e == cudaSuccess && ndev > 0, kNoDevice,
             "no CUDA device available: ocrs_b200 has no CPU fallback (cudaGetDeviceCount: " +
                 std::string(cudaGetErrorString(e)) + ")");
  OCRS_CHECK(p.device >= 0 && p.device < ndev, kInvalidArg, "device index out of range");
  device_ = p.device;
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  configure_device_pool(device_);
  OCRS_CUDA_CHECK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
  if (p.detection_model) det_ = Model::load(p.detection_model, p.detection_model_len, device_);
  if (p.recognition_model) rec_ = Model::load(p.recognition_model, p.recognition_model_len, device_);
  debug_ = p.debug;
  decode_method_ = p.decode_method;
  beam_width_ = p.beam_width;
  OCRS_CHECK(decode_method_ != DecodeMethod::kBeamSearch || (beam_width_ >= 1 && beam_width_ <= (uint32_t)img::kMaxBeamWidth),
             kInvalidArg, "beam width must be in [1, " + std::to_string(img::kMaxBeamWidth) + "]");
  alphabet_ = utf8_to_codepoints(p.has_alphabet ? p.alphabet_utf8 : std::string(kDefaultAlphabet));  // lib.rs:149-151
  if (p.has_allowed_chars) {  // lib.rs:153-170
    std::vector<uint32_t> allowed = utf8_to_codepoints(p.allowed_chars_utf8);
    excluded_mask_.assign(alphabet_.size() + 1, 0);
    for (size_t i = 0; i < alphabet_.size(); ++i) {
      bool ok = std::find(allowed.begin(), allowed.end(), alphabet_[i]) != allowed.end();
      if (!ok) excluded_mask_[i + 1] = 1;
    }
    has_excluded_ = true;
    d_excluded_.reserve(excluded_mask_.size());
    OCRS_CUDA_CHECK(cudaMemcpy(d_excluded_.ptr, excluded_mask_.data(), excluded_mask_.size(), cudaMemcpyHostToDevice));
  }
}

Engine::~Engine() {
  cudaSetDevice(device_);
  if (st_) {
    cudaStreamSynchronize(st_);
  }
  det_.reset();
  rec_.reset();
  if (ev_a_) cudaEventDestroy(ev_a_);
  if (ev_copy_) cudaEventDestroy(ev_copy_);
  if (ev_wait_) cudaEventDestroy(ev_wait_);
  if (ev_fork_) cudaEventDestroy(ev_fork_);
  for (auto e : aux_done_) cudaEventDestroy(e);
  for (auto s2 : aux_) cudaStreamDestroy(s2);
  if (ev_b_) cudaEventDestroy(ev_b_);
  if (st_) cudaStreamDestroy(st_);
}

void Engine::ensure_aux(int n) {
  if (!ev_fork_) OCRS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  while ((int)aux_.size() < n) {
    cudaStream_t s;
    cudaEvent_t e;
    OCRS_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    OCRS_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    aux_.push_back(s);
    aux_done_.push_back(e);
  }
}

void Engine::synchronize() {
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  wait_stream();
}

// Host wait for everything queued on the engine's stream.  Spinning (cudaStreamSynchronize) has the lowest latency
// and is right for a single engine; the pool's workers -- up to ten per GPU, most of their time waiting for their
// batch -- block on an event created with cudaEventBlockingSync instead, so that a waiting worker does not hold a
// host core (eight ranks x ten workers would otherwise spin on 80 of the box's cores).
void Engine::wait_stream() {
  if (!blocking_sync_) {
    OCRS_CUDA_CHECK(cudaStreamSynchronize(st_));
    return;
  }
  if (!ev_wait_) OCRS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_wait_, cudaEventBlockingSync | cudaEventDisableTiming));
  OCRS_CUDA_CHECK(cudaEventRecord(ev_wait_, st_));
  OCRS_CUDA_CHECK(cudaEventSynchronize(ev_wait_));
}

void Engine::set_profiling(bool on) {
  std::lock_guard<std::mutex> lk(mu_);
  prof_.enabled = on;
}

std::string Engine::profile_json(bool reset) {
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  wait_stream();
  prof_.collect();
  std::string j = "{";
  bool first = true;
  for (const auto& kv : prof_.results()) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.6f, \"calls\": %lld, \"launches\": %lld, \"flops\": %.6e, \"bytes\": %.6e}",
             first ? "" : ", ", kv.first.c_str(), kv.second.ms, (long long)kv.second.calls,
             (long long)kv.second.launches, kv.second.flops, kv.second.bytes);
    j += buf;
    first = false;
  }
  std::lock_guard<std::mutex> hlk(host_mu_);
  for (const auto& kv : host_ms_) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"host/%s\": {\"ms\": %.6f, \"calls\": %lld, \"launches\": 0, \"flops\": 0, \"bytes\": 0}",
             first ? "" : ", ", kv.first.c_str(), kv.second.first, (long long)kv.second.second);
    j += buf;
    first = false;
  }
  j += "}";
  if (reset) {
    prof_.reset();
    host_ms_.clear();
  }
  return j;
}

void Engine::timer_start() {
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  if (!ev_a_) {
    OCRS_CUDA_CHECK(cudaEventCreate(&ev_a_));
    OCRS_CUDA_CHECK(cudaEventCreate(&ev_b_));
  }
  OCRS_CUDA_CHECK(cudaEventRecord(ev_a_, st_));
}

float Engine::timer_stop() {
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  OCRS_CHECK(ev_a_ != nullptr, kInvalidArg, "timer_stop without timer_start");
  OCRS_CUDA_CHECK(cudaEventRecord(ev_b_, st_));
  OCRS_CUDA_CHECK(cudaEventSynchronize(ev_b_));
  float ms = 0;
  OCRS_CUDA_CHECK(cudaEventElapsedTime(&ms, ev_a_, ev_b_));
  return ms;
}

uint32_t Engine::rec_input_height() const {  // recognition.rs:332-337
  OCRS_CHECK(rec_ != nullptr, kModelNotLoaded, "Recognition model not loaded");
  const auto& s = rec_->input_shape();
  if (s.size() > 2 && s[2] >= 0) return (uint32_t)s[2];
  return 50;
}

std::unique_ptr<OcrInput> Engine::prepare_input(const void* pixels, int dtype, int order, int H, int W, int C,
                                                bool on_device) {
  // ImageSource::from_bytes / from_tensor validation (preprocess.rs:81-123)
  OCRS_CHECK(pixels != nullptr, kInvalidArg, "pixels is null");
  OCRS_CHECK(dtype == 0 || dtype == 1, kInvalidArg, "dtype must be 0 (u8) or 1 (f32)");
  OCRS_CHECK(order == 0 || order == 1, kInvalidArg, "order must be 0 (HWC) or 1 (CHW)");
  OCRS_CHECK(H >= 0 && W >= 0, kInvalidArg, "negative image size");
  OCRS_CHECK(C == 1 || C == 3 || C == 4, kUnsupportedChannelCount, "channel count is not 1, 3 or 4");
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  auto in = std::make_unique<OcrInput>();
  in->H = H;
  in->W = W;
  in->device = device_;
  size_t hw = (size_t)H * W;
  in->store = std::make_shared<Storage>(std::max<size_t>(hw, 1) * sizeof(float), st_);
  if (hw == 0) return in;
  size_t bytes = hw * C * (dtype == 0 ? 1 : 4);
  const void* dpx = pixels;
  std::shared_ptr<Storage> staging;  // freed (stream-ordered) after the conversion kernel
  if (!on_device) {
    staging = std::make_shared<Storage>(bytes, st_);
    OCRS_CUDA_CHECK(cudaMemcpyAsync(staging->ptr, pixels, bytes, cudaMemcpyHostToDevice, st_));
    dpx = staging->ptr;
    h2d_bytes_ += (int64_t)bytes;
    // the caller's buffer is only borrowed for this call (ImageSource<'a>): wait for the copy,
    // not for the kernel
    if (!ev_copy_) OCRS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_copy_, cudaEventDisableTiming));
    OCRS_CUDA_CHECK(cudaEventRecord(ev_copy_, st_));
  }
  int tk = prof_.begin("stage/prepare_image", st_);
  img::prepare_image(dpx, dtype, order, H, W, C, in->grey(), st_);
  prof_.end(tk, st_, 0, (double)bytes + (double)hw * 4);
  if (!on_device) OCRS_CUDA_CHECK(cudaEventSynchronize(ev_copy_));
  return in;
}

std::vector<std::unique_ptr<OcrInput>> Engine::prepare_inputs(const std::vector<PageSpec>& pages) {
  // batched fast path: u8 HWC RGB pages of one shape -> one staging buffer, one conversion launch, one wait.
  // Anything else goes page by page through prepare_input (same kernels, same bits).
  const int N = (int)pages.size();
  bool fast = N > 1;
  for (int i = 0; i < N && fast; ++i) {
    const PageSpec& p = pages[i];
    fast = p.pixels != nullptr && p.dtype == 0 && p.order == 0 && p.C == 3 && p.H == pages[0].H && p.W == pages[0].W &&
           (int64_t)p.H * p.W > 0 && ((int64_t)p.H * p.W) % 4 == 0 && (!p.on_device || reinterpret_cast<uintptr_t>(p.pixels) % 4 == 0);
  }
  std::vector<std::unique_ptr<OcrInput>> out;
  if (!fast) {
    for (const PageSpec& p : pages) out.push_back(prepare_input(p.pixels, p.dtype, p.order, p.H, p.W, p.C, p.on_device));
    return out;
  }
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  const int H = pages[0].H, W = pages[0].W;
  const size_t hw = (size_t)H * W, bytes = hw * 3;
  const size_t slot = (bytes + 255) / 256 * 256;
  int n_host = 0;
  for (const PageSpec& p : pages) n_host += p.on_device ? 0 : 1;
  std::shared_ptr<Storage> staging;
  if (n_host) staging = std::make_shared<Storage>(slot * (size_t)n_host, st_);
  std::vector<img::PagePrepare> tab((size_t)N);
  int k = 0;
  for (int i = 0; i < N; ++i) {
    auto in = std::make_unique<OcrInput>();
    in->H = H; in->W = W; in->device = device_;
    in->store = std::make_shared<Storage>(hw * sizeof(float), st_);
    const void* src = pages[i].pixels;
    if (!pages[i].on_device) {
      void* d = static_cast<char*>(staging->ptr) + slot * (size_t)k++;
      OCRS_CUDA_CHECK(cudaMemcpyAsync(d, pages[i].pixels, bytes, cudaMemcpyHostToDevice, st_));
      h2d_bytes_ += (int64_t)bytes;
      src = d;
    }
    tab[i] = img::PagePrepare{src, in->grey()};
    out.push_back(std::move(in));
  }
  if (n_host) {
    if (!ev_copy_) OCRS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_copy_, cudaEventDisableTiming));
    OCRS_CUDA_CHECK(cudaEventRecord(ev_copy_, st_));
  }
  tab_prep_.reserve(tab.size() * sizeof(img::PagePrepare));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(tab_prep_.ptr, tab.data(), tab.size() * sizeof(img::PagePrepare), cudaMemcpyHostToDevice, st_));
  int tk = prof_.begin("stage/prepare_image", st_);
  img::prepare_image_rgb8_batch(tab_prep_.as<img::PagePrepare>(), N, H, W, st_);
  prof_.end(tk, st_, 0, (double)N * ((double)bytes + (double)hw * 4));
  // the callers' buffers are only borrowed: wait for the copies, not for the kernel
  if (n_host) OCRS_CUDA_CHECK(cudaEventSynchronize(ev_copy_));
  return out;
}

namespace {
DTensor wrap_tensor(float* ptr, std::vector<int64_t> shape) {
  DTensor t;
  int64_t n = 1;
  for (auto d : shape) n *= d;
  t.storage = std::make_shared<Storage>(ptr, (size_t)n * 4);
  t.data = ptr;
  t.shape = std::move(shape);
  return t;
}
}  // namespace

std::vector<float> Engine::detect_text_pixels(const OcrInput& in) {
  OCRS_CHECK(det_ != nullptr, kModelNotLoaded, "Detection model not loaded");  // lib.rs:211
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  const auto& shp = det_->input_shape();
  OCRS_CHECK(shp.size() == 4 && shp[2] >= 0 && shp[3] >= 0, kRunFailed, "failed to get model dims");  // detection.rs:143
  int in_h = (int)shp[2], in_w = (int)shp[3];
  int H = in.H, W = in.W;
  int pad_bottom = std::max(in_h - H, 0), pad_right = std::max(in_w - W, 0);
  det_in_.reserve((size_t)in_h * in_w * 4);
  img::resize_padded(in.grey(), H, W, H + pad_bottom, W + pad_right, img::kBlackValue, det_in_.as<float>(),
                     in_h, in_w, 1, 0, 0, st_);
  int tok = det_->tc_token();
  DTensor out = det_->run(wrap_tensor(det_in_.as<float>(), {1, shp[1] < 0 ? 1 : shp[1], in_h, in_w}), st_);
  if (tok) {  // tensor-core chains in the detection model: check the fp16 range flag, rerun in fp32 if raised
    wait_stream();
    if (det_->take_tc_overflow(tok)) out = det_->run(wrap_tensor(det_in_.as<float>(), {1, shp[1] < 0 ? 1 : shp[1], in_h, in_w}), st_);
  }
  OCRS_CHECK(out.shape.size() == 4 && out.numel() == (int64_t)in_h * in_w, kWrongOutput,
             "detection output must be [1,1,H,W]");
  PageScratch& s = scratch_for(0, H, W);
  s.prob.reserve((size_t)H * W * 4 + 4);
  img::resize_threshold(out.data, in_h, in_w, in_h - pad_bottom, in_w - pad_right, s.prob.as<float>(),
                        s.mask.as<uint8_t>(), H, W, text_threshold_, st_);
  std::vector<float> host((size_t)H * W);
  OCRS_CUDA_CHECK(cudaMemcpyAsync(host.data(), s.prob.ptr, host.size() * 4, cudaMemcpyDeviceToHost, st_));
  wait_stream();
  return host;
}

std::vector<std::vector<RotatedRect>> Engine::detect_words(const std::vector<const OcrInput*>& pages) {
  OCRS_CHECK(det_ != nullptr, kModelNotLoaded, "Detection model not loaded");  // lib.rs:197
  std::lock_guard<std::mutex> lk(mu_);
  const int tok = det_->tc_token();
  auto result = detect_words_locked(pages);
  // a detection model with tensor-core-eligible convolutions whose activations left the fp16 range: the
  // model has switched itself to the fp32 kernels, repeat the call once (as recognize_text does)
  if (det_->take_tc_overflow(tok)) result = detect_words_locked(pages);
  return result;
}

std::vector<std::vector<RotatedRect>> Engine::detect_words_locked(const std::vector<const OcrInput*>& pages) {
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  const int N = (int)pages.size();
  std::vector<std::vector<RotatedRect>> result((size_t)N);
  if (N == 0) return result;
  HostTimer ht_all(this, "detect_words_total");
  const auto& shp = det_->input_shape();
  OCRS_CHECK(shp.size() == 4 && shp[2] >= 0 && shp[3] >= 0, kRunFailed, "failed to get model dims");
  const int in_h = (int)shp[2], in_w = (int)shp[3];
  const int64_t plane = (int64_t)in_h * in_w;
  det_in_.reserve((size_t)N * plane * 4);
  int tk = prof_.begin("stage/resize_in", st_);
  double rs_bytes = 0;
  {
    std::vector<img::PageResizeIn> tab((size_t)N);
    for (int i = 0; i < N; ++i) {
      const OcrInput& in = *pages[i];
      rs_bytes += 4.0 * ((double)in.H * in.W + (double)plane);
      OCRS_CHECK(in.device == device_, kInvalidArg, "input lives on another device");
      int pb = std::max(in_h - in.H, 0), pr = std::max(in_w - in.W, 0);
      tab[i] = img::PageResizeIn{in.grey(), in.H, in.W, in.H + pb, in.W + pr, img::resize_scale(in.H + pb, in_h), img::resize_scale(in.W + pr, in_w)};
    }
    tab_in_.reserve(tab.size() * sizeof(img::PageResizeIn));
    OCRS_CUDA_CHECK(cudaMemcpyAsync(tab_in_.ptr, tab.data(), tab.size() * sizeof(img::PageResizeIn), cudaMemcpyHostToDevice, st_));
    img::resize_padded_batch(tab_in_.as<img::PageResizeIn>(), N, img::kBlackValue, det_in_.as<float>(), in_h, in_w, plane, st_);
  }
  prof_.end(tk, st_, 0, rs_bytes);
  ModelCost cost;
  tk = prof_.begin("stage/det_net", st_);
  DTensor out = det_->run(wrap_tensor(det_in_.as<float>(), {N, shp[1] < 0 ? 1 : shp[1], in_h, in_w}), st_, &cost,
                          &prof_, "det/");
  prof_.end(tk, st_, cost.flops, cost.min_bytes);
  stats_.det_flops += cost.flops;
  OCRS_CHECK(out.shape.size() == 4 && out.numel() == (int64_t)N * plane, kWrongOutput,
             "detection output must be [N,1,H,W]");
  h_pin_.reserve((size_t)N * 8 * 4);
  int32_t* h_counters = h_pin_.as<int32_t>();
  // detection epilogue (slice + resize + threshold) of all pages in one launch
  {
    std::vector<img::PageResizeOut> tab((size_t)N);
    int max_h = 0, max_w = 0;
    double rt_bytes = 0;
    for (int i = 0; i < N; ++i) {
      const OcrInput& in = *pages[i];
      PageScratch& s = scratch_for(i, in.H, in.W);
      int pb = std::max(in_h - in.H, 0), pr = std::max(in_w - in.W, 0);
      tab[i] = img::PageResizeOut{out.data + i * plane, nullptr, s.mask.as<uint8_t>(), s.bits.as<uint32_t>(), in_h - pb, in_w - pr, in.H, in.W,
                                   img::resize_scale(in_h - pb, in.H), img::resize_scale(in_w - pr, in.W)};
      max_h = std::max(max_h, in.H);
      max_w = std::max(max_w, in.W);
      rt_bytes += 4.0 * (in_h - pb) * (in_w - pr) + (double)in.H * in.W;
    }
    tab_out_.reserve(tab.size() * sizeof(img::PageResizeOut));
    OCRS_CUDA_CHECK(cudaMemcpyAsync(tab_out_.ptr, tab.data(), tab.size() * sizeof(img::PageResizeOut), cudaMemcpyHostToDevice, st_));
    int t1 = prof_.begin("stage/resize_threshold", st_);
    img::resize_threshold_batch(tab_out_.as<img::PageResizeOut>(), N, in_w, max_h, max_w, text_threshold_, st_);
    prof_.end(t1, st_, 0, rt_bytes);
  }
  // labelling + contours -> rects of all pages: four launches for the batch
  {
    std::vector<img::CclPage> ctab((size_t)N);
    double cc_bytes = 0;
    for (int i = 0; i < N; ++i) {
      const OcrInput& in = *pages[i];
      PageScratch& s = *scratch_[i];
      img::CclPage& c = ctab[(size_t)i];
      c.bits = s.bits.as<uint32_t>();
      c.wstart = s.wstart.as<uint16_t>();
      c.labels = s.bufs.labels;
      c.H = in.H; c.W = in.W; c.wd = (in.W + 31) / 32;
      c.bufs = s.bufs;
      cc_bytes += (double)in.H * c.wd * 4.0 * 3.0;  // packed mask: init, union, flatten passes
    }
    tab_ccl_.reserve(ctab.size() * sizeof(img::CclPage));
    OCRS_CUDA_CHECK(cudaMemcpyAsync(tab_ccl_.ptr, ctab.data(), ctab.size() * sizeof(img::CclPage), cudaMemcpyHostToDevice, st_));
    int t2 = prof_.begin("stage/components_to_rects", st_);
    img::find_component_rects_batch(tab_ccl_.as<img::CclPage>(), ctab.data(), N, 2.0f /* detection.rs:50 */,
                                    3.0f /* detection.rs:116 */, min_area_, st_);
    prof_.end(t2, st_, 0, cc_bytes);
    for (int i = 0; i < N; ++i) {
      OCRS_CUDA_CHECK(cudaMemcpyAsync(h_counters + 8 * i, scratch_[i]->bufs.counters, 8 * 4, cudaMemcpyDeviceToHost, st_));
      d2h_bytes_ += 32;
    }
  }
  wait_stream();
  std::vector<std::vector<int32_t>> roots((size_t)N);
  for (int i = 0; i < N; ++i) {
    const int32_t* c = h_counters + 8 * i;
    OCRS_CHECK(c[2] == 0, kRunFailed,
               c[2] == 1 ? "component table overflow" : "contour point pool overflow (mask too fragmented)");
    int n = c[3];
    result[i].resize((size_t)n);
    roots[i].resize((size_t)n);
    if (n > 0) {
      PageScratch& s = *scratch_[i];
      OCRS_CUDA_CHECK(cudaMemcpyAsync(result[i].data(), s.bufs.rects, (size_t)n * sizeof(RotatedRect),
                                      cudaMemcpyDeviceToHost, st_));
      OCRS_CUDA_CHECK(cudaMemcpyAsync(roots[i].data(), s.bufs.rect_root, (size_t)n * 4, cudaMemcpyDeviceToHost, st_));
      d2h_bytes_ += (int64_t)n * (sizeof(RotatedRect) + 4);
    }
  }
  wait_stream();
  for (int i = 0; i < N; ++i) {
    // restore contour discovery (raster) order: layout depends on it (layout_analysis.rs:116-119)
    size_t n = result[i].size();
    std::vector<size_t> order(n);
    for (size_t k = 0; k < n; ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return roots[i][a] < roots[i][b]; });
    std::vector<RotatedRect> sorted(n);
    for (size_t k = 0; k < n; ++k) sorted[k] = result[i][order[k]];
    result[i].swap(sorted);
    stats_.n_words += (int64_t)n;
  }
  return result;
}

std::vector<std::vector<RotatedRect>> Engine::find_text_lines(const std::vector<RotatedRect>& words) const {
  return layout::find_text_lines(words);
}

namespace {
struct RecLine {
  int page, index;
  std::vector<PointI> poly;
  RectI rect;  // polygon bounding rect
  uint32_t resized_width;
  int group_width;
};
}  // namespace

std::vector<std::vector<TextLine>> Engine::recognize_text(
    const std::vector<const OcrInput*>& pages,
    const std::vector<std::vector<std::vector<RotatedRect>>>& lines_per_page) {
  OCRS_CHECK(rec_ != nullptr, kModelNotLoaded, "Recognition model not loaded");  // lib.rs:254
  OCRS_CHECK(pages.size() == lines_per_page.size(), kInvalidArg, "pages / lines size mismatch");
  std::lock_guard<std::mutex> lk(mu_);
  const int tok = rec_->tc_token();
  auto result = recognize_text_locked(pages, lines_per_page);
  // split-fp16 range overflow in the tensor-core convolutions: the model has switched itself to
  // the fp32 kernels, repeat the call once
  if (rec_->take_tc_overflow(tok)) result = recognize_text_locked(pages, lines_per_page);
  return result;
}

std::vector<std::vector<TextLine>> Engine::recognize_text_locked(
    const std::vector<const OcrInput*>& pages,
    const std::vector<std::vector<std::vector<RotatedRect>>>& lines_per_page) {
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  const int n_pages = (int)pages.size();
  const int rec_h = (int)rec_input_height();
  std::vector<std::vector<TextLine>> result((size_t)n_pages);
  HostTimer ht_all(this, "recognize_text_total");
  auto ht_prep = std::make_unique<HostTimer>(this, "rec_line_geometry");

  // ---- host: per-line geometry (recognition.rs:429-446) ----
  std::vector<RecLine> lines;
  for (int p = 0; p < n_pages; ++p) {
    result[p].resize(lines_per_page[p].size());
    for (size_t li = 0; li < lines_per_page[p].size(); ++li) {
      const auto& words = lines_per_page[p][li];
      OCRS_CHECK(!words.empty(), kInvalidArg, "line has no words");  // recognition.rs:433 (expect)
      RectI line_rect;
      layout::line_integral_rect(words, &line_rect);
      RecLine rl;
      rl.page = p;
      rl.index = (int)li;
      rl.resized_width = layout::resized_line_width(geom::rwidth(line_rect), geom::rheight(line_rect), rec_h);
      rl.group_width = (int)(((rl.resized_width + 49u) / 50u) * 50u);  // next_multiple_of(50) (:437)
      rl.poly = layout::line_polygon(words);
      rl.rect = layout::polygon_bounding_rect(rl.poly);
      lines.push_back(std::move(rl));
    }
  }
  const int n_lines = (int)lines.size();
  if (n_lines == 0) return result;
  OCRS_CHECK(n_lines <= 65535, kInvalidArg, "too many lines in one call (max 65535)");

  // group by padded width (ascending); keep line order inside a group
  std::vector<int> order((size_t)n_lines);
  for (int i = 0; i < n_lines; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lines[a].group_width < lines[b].group_width; });

  // ---- device line descriptors ----
  std::vector<img::LineDesc> descs((size_t)n_lines);
  std::vector<int32_t> poly_xy;
  int64_t dst_total = 0, cross_total = 0;
  int max_gw = 0, max_rows = 0;
  for (int k = 0; k < n_lines; ++k) {
    const RecLine& rl = lines[order[k]];
    img::LineDesc& d = descs[k];
    d.poly_off = (int32_t)(poly_xy.size() / 2);
    d.poly_n = (int32_t)rl.poly.size();
    int non_horizontal = 0;
    for (size_t v = 0; v < rl.poly.size(); ++v) {
      poly_xy.push_back(rl.poly[v].x);
      poly_xy.push_back(rl.poly[v].y);
      if (rl.poly[v].y != rl.poly[(v + 1) % rl.poly.size()].y) ++non_horizontal;
    }
    d.top = rl.rect.top;
    d.left = rl.rect.left;
    d.lh = std::max(geom::rheight(rl.rect), 0);
    d.lw = std::max(geom::rwidth(rl.rect), 0);
    d.resized_width = (int32_t)rl.resized_width;
    d.group_width = rl.group_width;
    d.dst_off = dst_total;
    d.cross_off = cross_total;
    d.max_cross = non_horizontal;
    d.page = rl.page;
    dst_total += (int64_t)rec_h * rl.group_width;
    cross_total += (int64_t)d.lh * (non_horizontal + 1);
    max_gw = std::max(max_gw, rl.group_width);
    max_rows = std::max(max_rows, d.lh);
  }
  // page table
  std::vector<const float*> page_ptrs((size_t)n_pages);
  std::vector<int> page_hw((size_t)2 * n_pages);
  for (int p = 0; p < n_pages; ++p) {
    OCRS_CHECK(pages[p]->device == device_, kInvalidArg, "input lives on another device");
    page_ptrs[p] = pages[p]->grey();
    page_hw[p] = pages[p]->H;
    page_hw[n_pages + p] = pages[p]->W;
  }
  size_t tab_bytes = (size_t)n_pages * (sizeof(float*) + 2 * sizeof(int));
  page_tab_.reserve(tab_bytes);
  line_desc_.reserve(descs.size() * sizeof(img::LineDesc));
  poly_.reserve(poly_xy.size() * 4 + 4);
  cross_.reserve((size_t)cross_total * 4 + 4);
  rec_batch_.reserve((size_t)dst_total * 4);
  OCRS_CUDA_CHECK(cudaMemcpyAsync(page_tab_.ptr, page_ptrs.data(), n_pages * sizeof(float*), cudaMemcpyHostToDevice, st_));
  int* d_page_hw = reinterpret_cast<int*>(page_tab_.as<char>() + n_pages * sizeof(float*));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(d_page_hw, page_hw.data(), 2 * n_pages * sizeof(int), cudaMemcpyHostToDevice, st_));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(line_desc_.ptr, descs.data(), descs.size() * sizeof(img::LineDesc), cudaMemcpyHostToDevice, st_));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(poly_.ptr, poly_xy.data(), poly_xy.size() * 4, cudaMemcpyHostToDevice, st_));
  h2d_bytes_ += (int64_t)(descs.size() * sizeof(img::LineDesc) + poly_xy.size() * 4 + tab_bytes);
  int tkc = prof_.begin("stage/crop_lines", st_);
  img::crop_lines(page_tab_.as<const float*>(), d_page_hw, d_page_hw + n_pages, line_desc_.as<img::LineDesc>(),
                  n_lines, poly_.as<int32_t>(), cross_.as<int32_t>(), rec_batch_.as<float>(), rec_h, max_gw, max_rows,
                  st_);
  prof_.end(tkc, st_, 0, 2.0 * 4.0 * (double)dst_total);
  ht_prep.reset();
  // host vectors must outlive the async copies
  {
    HostTimer ht(this, "rec_sync_after_crop");
    wait_stream();
  }
  auto ht_launch = std::make_unique<HostTimer>(this, "rec_enqueue_networks");

  // ---- recognition network + CTC per (group, chunk) ----
  struct Chunk { int first, count, gw, T; int64_t out_off; };
  std::vector<Chunk> chunks;
  const int64_t kMaxActBytes = (int64_t)3 << 30;  // bound for the largest activation of one run
  int64_t out_total = 0;
  for (int k = 0; k < n_lines;) {
    int gw = descs[k].group_width;
    int e = k;
    while (e < n_lines && descs[e].group_width == gw) ++e;
    int64_t per_line = (int64_t)32 * rec_h * gw * 4;
    int max_b = (int)std::max<int64_t>(1, kMaxActBytes / std::max<int64_t>(per_line, 1));
    for (int s = k; s < e; s += max_b) chunks.push_back(Chunk{s, std::min(max_b, e - s), gw, 0, 0});
    k = e;
  }
  const size_t n_classes = alphabet_.size() + 1;
  // device output layout per chunk: labels [B*Tmax], pos [B*Tmax], counts [B]; Tmax = gw (>= T)
  for (auto& c : chunks) {
    c.out_off = out_total;
    out_total += (int64_t)c.count * c.gw * 2 + c.count;
  }
  ctc_out_.reserve((size_t)out_total * 4);
  h_pin_.reserve((size_t)out_total * 4);
  auto check_logits = [&](const std::vector<int64_t>& shape, int count) {
    OCRS_CHECK(shape.size() == 3, kWrongOutput,
               "expected recognition output to have 3 dims but it has " + std::to_string(shape.size()));  // :350
    OCRS_CHECK(shape[1] == count, kWrongOutput, "recognition output batch dim mismatch");
    OCRS_CHECK((size_t)shape[2] == n_classes, kWrongOutput,
               "output column count (" + std::to_string(shape[2]) + ") does not match alphabet size (" +
                   std::to_string(n_classes) + ")");  // recognition.rs:487-493
  };
  // CTC decode of a set of lines sharing one logits buffer: greedy = one argmax launch + one
  // collapse launch; beam search = one block per line (recognition.rs:509-515)
  std::deque<std::vector<img::CtcLine>> cl_keep;  // descriptors must outlive their async upload
  auto decode_lines = [&](const float* logits_ptr, int64_t rows, std::vector<img::CtcLine>&& cl_in) {
    cl_keep.push_back(std::move(cl_in));
    std::vector<img::CtcLine>& cl = cl_keep.back();
    const bool beam = decode_method_ == DecodeMethod::kBeamSearch;
    int64_t n_nodes = 0;
    if (beam)
      for (auto& L : cl) {
        L.node_off = n_nodes;
        n_nodes += img::ctc_beam_nodes_per_line(L.T, (int)beam_width_);
      }
    const size_t lab_bytes = ((size_t)rows * 4 + 15) / 16 * 16;
    const size_t desc_bytes = (cl.size() * sizeof(img::CtcLine) + 15) / 16 * 16;
    ctc_scratch_.reserve(lab_bytes + desc_bytes + (size_t)n_nodes * 24 + 64);
    int32_t* row_labels = ctc_scratch_.as<int32_t>();
    auto* d_cl = reinterpret_cast<img::CtcLine*>(ctc_scratch_.as<char>() + lab_bytes);
    int32_t* d_nodes = reinterpret_cast<int32_t*>(ctc_scratch_.as<char>() + lab_bytes + desc_bytes);
    OCRS_CUDA_CHECK(cudaMemcpyAsync(d_cl, cl.data(), cl.size() * sizeof(img::CtcLine), cudaMemcpyHostToDevice, st_));
    h2d_bytes_ += (int64_t)(cl.size() * sizeof(img::CtcLine));
    const uint8_t* excl = has_excluded_ ? d_excluded_.as<uint8_t>() : nullptr;
    int tkt = prof_.begin(beam ? "stage/ctc_beam" : "stage/ctc_greedy", st_);
    if (beam)
      img::ctc_beam_search(logits_ptr, (int)n_classes, excl, d_cl, (int)cl.size(), (int)beam_width_, d_nodes,
                           ctc_out_.as<int32_t>(), st_);
    else
      img::ctc_greedy_packed(logits_ptr, rows, (int)n_classes, excl, row_labels, d_cl, (int)cl.size(),
                             ctc_out_.as<int32_t>(), st_);
    prof_.end(tkt, st_, 0, 4.0 * (double)rows * (double)n_classes);
  };
  auto chunk_lines = [&](const Chunk& c, int64_t row_off, std::vector<img::CtcLine>* cl) {
    OCRS_CHECK(c.T <= c.gw, kWrongOutput, "recognition output longer than its input");
    stats_.n_timesteps += (int64_t)c.T * c.count;
    for (int b = 0; b < c.count; ++b) {
      img::CtcLine L{};
      L.base = row_off + b;
      L.stride = c.count;
      L.T = c.T;
      L.lab_off = c.out_off + (int64_t)b * c.T;
      L.pos_off = c.out_off + (int64_t)c.count * c.gw + (int64_t)b * c.T;
      L.cnt_off = c.out_off + (int64_t)c.count * c.gw * 2 + b;
      cl->push_back(L);
    }
  };
  std::vector<DTensor> feats;  // per-group feature sequences (packed path)
  if (rec_->has_seq_head()) {
    // ---- packed path: conv prefix per width group, then ONE ragged GRU/Linear pass over all lines ----
    OCRS_CHECK((size_t)rec_->seq_head_classes() == n_classes, kWrongOutput,
               "output column count (" + std::to_string(rec_->seq_head_classes()) + ") does not match alphabet size (" +
                   std::to_string(n_classes) + ")");
    const int Cf = rec_->seq_head_channels();
    if (rec_->has_packed_prefix(rec_h)) {
      // ---- ragged path: every layer of the conv prefix runs once over all width groups; the packed
      // sequence head follows.  Passes bound the activation memory of huge calls (normally one pass). ----
      size_t ci = 0;
      while (ci < chunks.size()) {
        size_t ce = ci;
        int64_t bytes = 0;
        std::vector<Model::PrefixGroup> pg;
        while (ce < chunks.size()) {
          const Chunk& c = chunks[ce];
          const int64_t b = (int64_t)32 * rec_h * c.gw * 4 * c.count;
          if (!pg.empty() && bytes + b > kMaxActBytes) break;
          bytes += b;
          Model::PrefixGroup g;
          g.x_base = rec_batch_.as<float>();
          g.x_off = descs[c.first].dst_off;
          g.N = c.count;
          g.W = c.gw;
          pg.push_back(g);
          ++ce;
        }
        std::vector<Model::PackedGroup> groups;
        ModelCost cost;
        int tkr = prof_.begin("stage/rec_prefix", st_);
        DTensor x = rec_->run_prefix_packed(pg, rec_h, st_, &groups, &cost, &prof_, "rec/");
        prof_.end(tkr, st_, cost.flops, 0);
        stats_.rec_flops += cost.flops;
        stats_.rec_batches += 1;
        OCRS_CHECK(x.shape.size() == 2 && x.shape[1] == Cf, kWrongOutput, "recognition feature rows must be [rows, C]");
        const int64_t rows = x.shape[0];
        ModelCost hcost;
        int tkh = prof_.begin("stage/rec_seq_head", st_);
        DTensor logits = rec_->run_seq_head(x.data, rows, groups, st_, &hcost, &prof_, "rec/");
        prof_.end(tkh, st_, hcost.flops, 0);
        stats_.rec_flops += hcost.flops;
        std::vector<img::CtcLine> cl;
        for (size_t k = ci; k < ce; ++k) {
          chunks[k].T = groups[k - ci].T;
          chunk_lines(chunks[k], groups[k - ci].row_off, &cl);
        }
        decode_lines(logits.data, rows, std::move(cl));
        feats.push_back(std::move(x));       // alive until the final synchronisation below
        feats.push_back(std::move(logits));
        ci = ce;
      }
    } else {
    std::vector<Model::PackedGroup> groups;
    int64_t rows = 0;
    // the per-group conv prefixes are independent and individually too small to fill 148 SMs:
    // spread them over side streams, join before the packed head
    const int n_aux = prof_.enabled ? 1 : (int)std::min<size_t>(chunks.size(), 6);
    ensure_aux(n_aux);
    OCRS_CUDA_CHECK(cudaEventRecord(ev_fork_, st_));
    for (int a = 0; a < n_aux; ++a) OCRS_CUDA_CHECK(cudaStreamWaitEvent(aux_[a], ev_fork_, 0));
    int ci = 0;
    for (auto& c : chunks) {
      cudaStream_t sa = aux_[ci++ % n_aux];
      float* in_ptr = rec_batch_.as<float>() + descs[c.first].dst_off;
      ModelCost cost;
      int tkr = prof_.begin("stage/rec_prefix", sa);
      DTensor x = rec_->run_prefix(wrap_tensor(in_ptr, {c.count, 1, rec_h, c.gw}), sa, &cost, &prof_, "rec/");
      prof_.end(tkr, sa, cost.flops, 0);
      stats_.rec_flops += cost.flops;
      stats_.rec_batches += 1;
      OCRS_CHECK(x.shape.size() == 3 && x.shape[1] == c.count && x.shape[2] == Cf, kWrongOutput,
                 "recognition feature sequence must be [T, N, C]");
      c.T = (int)x.shape[0];
      groups.push_back(Model::PackedGroup{c.T, c.count, rows});
      rows += (int64_t)c.T * c.count;
      feats.push_back(std::move(x));
    }
    for (int a = 0; a < n_aux; ++a) {
      OCRS_CUDA_CHECK(cudaEventRecord(aux_done_[a], aux_[a]));
      OCRS_CUDA_CHECK(cudaStreamWaitEvent(st_, aux_done_[a], 0));
    }
    auto packed = std::make_shared<Storage>((size_t)rows * Cf * 4, st_);
    for (size_t g = 0; g < feats.size(); ++g)
      OCRS_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<float*>(packed->ptr) + groups[g].row_off * Cf, feats[g].data,
                                      (size_t)feats[g].numel() * 4, cudaMemcpyDeviceToDevice, st_));
    // `feats` (allocated on the side streams) stay alive until the final synchronisation below
    ModelCost cost;
    int tkh = prof_.begin("stage/rec_seq_head", st_);
    DTensor logits = rec_->run_seq_head(reinterpret_cast<const float*>(packed->ptr), rows, groups, st_, &cost, &prof_, "rec/");
    prof_.end(tkh, st_, cost.flops, 0);
    stats_.rec_flops += cost.flops;
    std::vector<img::CtcLine> cl;
    cl.reserve((size_t)n_lines);
    for (size_t g = 0; g < chunks.size(); ++g) chunk_lines(chunks[g], groups[g].row_off, &cl);
    decode_lines(logits.data, rows, std::move(cl));
    }
  } else {
    for (auto& c : chunks) {
      float* in_ptr = rec_batch_.as<float>() + descs[c.first].dst_off;
      ModelCost cost;
      int tkr = prof_.begin("stage/rec_net", st_);
      DTensor logits = rec_->run(wrap_tensor(in_ptr, {c.count, 1, rec_h, c.gw}), st_, &cost, &prof_, "rec/");
      prof_.end(tkr, st_, cost.flops, cost.min_bytes);
      stats_.rec_flops += cost.flops;
      stats_.rec_batches += 1;
      check_logits(logits.shape, c.count);
      c.T = (int)logits.shape[0];
      std::vector<img::CtcLine> cl;
      chunk_lines(c, 0, &cl);
      decode_lines(logits.data, (int64_t)c.T * c.count, std::move(cl));
    }
  }
  OCRS_CUDA_CHECK(cudaMemcpyAsync(h_pin_.ptr, ctc_out_.ptr, (size_t)out_total * 4, cudaMemcpyDeviceToHost, st_));
  d2h_bytes_ += out_total * 4;
  ht_launch.reset();
  {
    HostTimer ht(this, "rec_sync_wait_gpu");
    wait_stream();
  }
  HostTimer ht_asm(this, "rec_assemble_text");

  // ---- host: CTC steps -> characters with boxes (recognition.rs:241-311) ----
  const int32_t* h = h_pin_.as<int32_t>();
  for (const auto& c : chunks) {
    const int32_t* lab = h + c.out_off;
    const int32_t* pos = lab + (int64_t)c.count * c.gw;
    const int32_t* cnt = pos + (int64_t)c.count * c.gw;
    for (int b = 0; b < c.count; ++b) {
      const RecLine& rl = lines[order[c.first + b]];
      const RectI line_rect = rl.rect;
      float x_scale = (float)geom::rwidth(line_rect) / (float)rl.resized_width;
      uint32_t downsample = geom::f2u(roundf((float)c.gw / (float)c.T));  // :254-255
      int n_steps = cnt[b];
      TextLine tl;
      for (int i = 0; i < n_steps; ++i) {
        uint32_t label = (uint32_t)lab[(int64_t)b * c.T + i];
        uint32_t start_u = (uint32_t)pos[(int64_t)b * c.T + i] * downsample;
        uint32_t end_u = (i + 1 < n_steps) ? (uint32_t)pos[(int64_t)b * c.T + i + 1] * downsample : rl.resized_width;
        int start_x = line_rect.left + geom::f2i((float)start_u * x_scale);  // :271-272
        int end_x = line_rect.left + geom::f2i((float)end_u * x_scale);
        if (start_x >= line_rect.right) continue;  // :278
        uint32_t ch = (label >= 1 && label - 1 < alphabet_.size()) ? alphabet_[label - 1] : (uint32_t)'?';
        RectI r;
        bool ok = layout::polygon_slice_bounding_rect(rl.poly, start_x, end_x, &r);
        OCRS_CHECK(ok, kInternal, "invalid X coords");  // recognition.rs:299 (expect)
        tl.chars.push_back(TextChar{ch, r});
      }
      tl.present = !tl.chars.empty();
      result[rl.page][rl.index] = std::move(tl);
      stats_.n_lines += 1;
    }
  }
  return result;
}

std::vector<float> Engine::prepare_recognition_input(const OcrInput& in, const std::vector<RotatedRect>& line,
                                                     int* out_h, int* out_w) {
  OCRS_CHECK(rec_ != nullptr, kModelNotLoaded, "Recognition model not loaded");  // lib.rs:274
  OCRS_CHECK(!line.empty(), kInvalidArg, "line has no words");
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  const int rec_h = (int)rec_input_height();
  RectI line_rect;
  layout::line_integral_rect(line, &line_rect);
  uint32_t rw = layout::resized_line_width(geom::rwidth(line_rect), geom::rheight(line_rect), rec_h);
  std::vector<PointI> poly = layout::line_polygon(line);
  RectI pr = layout::polygon_bounding_rect(poly);
  img::LineDesc d{};
  std::vector<int32_t> poly_xy;
  int non_horizontal = 0;
  for (size_t v = 0; v < poly.size(); ++v) {
    poly_xy.push_back(poly[v].x);
    poly_xy.push_back(poly[v].y);
    if (poly[v].y != poly[(v + 1) % poly.size()].y) ++non_horizontal;
  }
  d.poly_off = 0; d.poly_n = (int32_t)poly.size();
  d.top = pr.top; d.left = pr.left;
  d.lh = std::max(geom::rheight(pr), 0); d.lw = std::max(geom::rwidth(pr), 0);
  d.resized_width = (int32_t)rw; d.group_width = (int32_t)rw;
  d.dst_off = 0; d.cross_off = 0; d.max_cross = non_horizontal; d.page = 0;
  const float* page_ptr = in.grey();
  int hw[2] = {in.H, in.W};
  page_tab_.reserve(sizeof(float*) + 2 * sizeof(int));
  line_desc_.reserve(sizeof(d));
  poly_.reserve(poly_xy.size() * 4 + 4);
  cross_.reserve((size_t)d.lh * (non_horizontal + 1) * 4 + 4);
  rec_batch_.reserve((size_t)rec_h * rw * 4 + 4);
  OCRS_CUDA_CHECK(cudaMemcpyAsync(page_tab_.ptr, &page_ptr, sizeof(float*), cudaMemcpyHostToDevice, st_));
  int* d_hw = reinterpret_cast<int*>(page_tab_.as<char>() + sizeof(float*));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(d_hw, hw, sizeof(hw), cudaMemcpyHostToDevice, st_));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(line_desc_.ptr, &d, sizeof(d), cudaMemcpyHostToDevice, st_));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(poly_.ptr, poly_xy.data(), poly_xy.size() * 4, cudaMemcpyHostToDevice, st_));
  img::crop_lines(page_tab_.as<const float*>(), d_hw, d_hw + 1, line_desc_.as<img::LineDesc>(), 1, poly_.as<int32_t>(),
                  cross_.as<int32_t>(), rec_batch_.as<float>(), rec_h, (int)rw, d.lh, st_);
  std::vector<float> out((size_t)rec_h * rw);
  OCRS_CUDA_CHECK(cudaMemcpyAsync(out.data(), rec_batch_.ptr, out.size() * 4, cudaMemcpyDeviceToHost, st_));
  wait_stream();
  *out_h = rec_h;
  *out_w = (int)rw;
  return out;
}

std::vector<Engine::TextPixels> Engine::detect_text_pixels_batch(const std::vector<const OcrInput*>& pages, bool want_map,
                                                                 bool want_mask) {
  OCRS_CHECK(det_ != nullptr, kModelNotLoaded, "Detection model not loaded");  // lib.rs:211
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  const int N = (int)pages.size();
  std::vector<TextPixels> result((size_t)N);
  if (N == 0) return result;
  const auto& shp = det_->input_shape();
  OCRS_CHECK(shp.size() == 4 && shp[2] >= 0 && shp[3] >= 0, kRunFailed, "failed to get model dims");
  const int in_h = (int)shp[2], in_w = (int)shp[3];
  const int64_t plane = (int64_t)in_h * in_w;
  det_in_.reserve((size_t)N * plane * 4);
  {
    std::vector<img::PageResizeIn> tab((size_t)N);
    for (int i = 0; i < N; ++i) {
      const OcrInput& in = *pages[i];
      OCRS_CHECK(in.device == device_, kInvalidArg, "input lives on another device");
      int pb = std::max(in_h - in.H, 0), pr = std::max(in_w - in.W, 0);
      tab[i] = img::PageResizeIn{in.grey(), in.H, in.W, in.H + pb, in.W + pr, img::resize_scale(in.H + pb, in_h), img::resize_scale(in.W + pr, in_w)};
    }
    tab_in_.reserve(tab.size() * sizeof(img::PageResizeIn));
    OCRS_CUDA_CHECK(cudaMemcpyAsync(tab_in_.ptr, tab.data(), tab.size() * sizeof(img::PageResizeIn), cudaMemcpyHostToDevice, st_));
    img::resize_padded_batch(tab_in_.as<img::PageResizeIn>(), N, img::kBlackValue, det_in_.as<float>(), in_h, in_w, plane, st_);
  }
  auto run_net = [&] { return det_->run(wrap_tensor(det_in_.as<float>(), {N, shp[1] < 0 ? 1 : shp[1], in_h, in_w}), st_); };
  const int tok = det_->tc_token();
  DTensor out = run_net();
  if (tok) {
    wait_stream();
    if (det_->take_tc_overflow(tok)) out = run_net();
  }
  OCRS_CHECK(out.shape.size() == 4 && out.numel() == (int64_t)N * plane, kWrongOutput, "detection output must be [N,1,H,W]");
  for (int i = 0; i < N; ++i) {
    const OcrInput& in = *pages[i];
    PageScratch& s = scratch_for(i, in.H, in.W);
    s.prob.reserve((size_t)in.H * in.W * 4 + 4);
    int pb = std::max(in_h - in.H, 0), pr = std::max(in_w - in.W, 0);
    img::resize_threshold(out.data + i * plane, in_h, in_w, in_h - pb, in_w - pr, s.prob.as<float>(), s.mask.as<uint8_t>(), in.H,
                          in.W, text_threshold_, st_);
    result[i].H = in.H;
    result[i].W = in.W;
    const size_t hw = (size_t)in.H * in.W;
    if (want_map) {
      result[i].map.resize(hw);
      OCRS_CUDA_CHECK(cudaMemcpyAsync(result[i].map.data(), s.prob.ptr, hw * 4, cudaMemcpyDeviceToHost, st_));
      d2h_bytes_ += (int64_t)hw * 4;
    }
    if (want_mask) {
      result[i].mask.resize(hw);
      OCRS_CUDA_CHECK(cudaMemcpyAsync(result[i].mask.data(), s.mask.ptr, hw, cudaMemcpyDeviceToHost, st_));
      d2h_bytes_ += (int64_t)hw;
    }
  }
  wait_stream();
  return result;
}

Engine::LineImages Engine::prepare_recognition_inputs(const OcrInput& in,
                                                      const std::vector<std::vector<RotatedRect>>& lines) {
  OCRS_CHECK(rec_ != nullptr, kModelNotLoaded, "Recognition model not loaded");  // lib.rs:274
  std::lock_guard<std::mutex> lk(mu_);
  OCRS_CUDA_CHECK(cudaSetDevice(device_));
  OCRS_CHECK(in.device == device_, kInvalidArg, "input lives on another device");
  const int rec_h = (int)rec_input_height();
  LineImages out;
  out.height = rec_h;
  const int n = (int)lines.size();
  if (n == 0) return out;
  std::vector<img::LineDesc> descs((size_t)n);
  std::vector<int32_t> poly_xy;
  int64_t dst_total = 0, cross_total = 0;
  int max_w = 0, max_rows = 0;
  for (int k = 0; k < n; ++k) {
    OCRS_CHECK(!lines[k].empty(), kInvalidArg, "line has no words");
    RectI line_rect;
    layout::line_integral_rect(lines[k], &line_rect);
    const uint32_t rw = layout::resized_line_width(geom::rwidth(line_rect), geom::rheight(line_rect), rec_h);
    std::vector<PointI> poly = layout::line_polygon(lines[k]);
    const RectI pr = layout::polygon_bounding_rect(poly);
    img::LineDesc& d = descs[(size_t)k];
    d.poly_off = (int32_t)(poly_xy.size() / 2);
    d.poly_n = (int32_t)poly.size();
    int non_horizontal = 0;
    for (size_t v = 0; v < poly.size(); ++v) {
      poly_xy.push_back(poly[v].x);
      poly_xy.push_back(poly[v].y);
      if (poly[v].y != poly[(v + 1) % poly.size()].y) ++non_horizontal;
    }
    d.top = pr.top; d.left = pr.left;
    d.lh = std::max(geom::rheight(pr), 0); d.lw = std::max(geom::rwidth(pr), 0);
    d.resized_width = (int32_t)rw; d.group_width = (int32_t)rw;  // no batch padding: exactly [height, w']
    d.dst_off = dst_total; d.cross_off = cross_total; d.max_cross = non_horizontal; d.page = 0;
    out.widths.push_back((int)rw);
    out.offsets.push_back((size_t)dst_total);
    dst_total += (int64_t)rec_h * rw;
    cross_total += (int64_t)d.lh * (non_horizontal + 1);
    max_w = std::max(max_w, (int)rw);
    max_rows = std::max(max_rows, d.lh);
  }
  const float* page_ptr = in.grey();
  int hw[2] = {in.H, in.W};
  page_tab_.reserve(sizeof(float*) + 2 * sizeof(int));
  line_desc_.reserve(descs.size() * sizeof(img::LineDesc));
  poly_.reserve(poly_xy.size() * 4 + 4);
  cross_.reserve((size_t)cross_total * 4 + 4);
  rec_batch_.reserve((size_t)dst_total * 4 + 4);
  OCRS_CUDA_CHECK(cudaMemcpyAsync(page_tab_.ptr, &page_ptr, sizeof(float*), cudaMemcpyHostToDevice, st_));
  int* d_hw = reinterpret_cast<int*>(page_tab_.as<char>() + sizeof(float*));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(d_hw, hw, sizeof(hw), cudaMemcpyHostToDevice, st_));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(line_desc_.ptr, descs.data(), descs.size() * sizeof(img::LineDesc), cudaMemcpyHostToDevice, st_));
  OCRS_CUDA_CHECK(cudaMemcpyAsync(poly_.ptr, poly_xy.data(), poly_xy.size() * 4, cudaMemcpyHostToDevice, st_));
  img::crop_lines(page_tab_.as<const float*>(), d_hw, d_hw + 1, line_desc_.as<img::LineDesc>(), n, poly_.as<int32_t>(),
                  cross_.as<int32_t>(), rec_batch_.as<float>(), rec_h, max_w, max_rows, st_);
  out.images.resize((size_t)dst_total);
  OCRS_CUDA_CHECK(cudaMemcpyAsync(out.images.data(), rec_batch_.ptr, out.images.size() * 4, cudaMemcpyDeviceToHost, st_));
  d2h_bytes_ += dst_total * 4;
  wait_stream();
  return out;
}

std::vector<std::vector<TextLine>> Engine::ocr_pages(const std::vector<const OcrInput*>& pages) {
  auto words = detect_words(pages);
  std::vector<std::vector<std::vector<RotatedRect>>> lines(pages.size());
  {
    // layout analysis is pure host code and independent per page: a few threads share the pages
    HostTimer ht(this, "find_text_lines");
    const size_t nt = std::min<size_t>(pages.size(), (size_t)std::max(1, layout_threads_));
    if (nt <= 1) {
      for (size_t p = 0; p < pages.size(); ++p) lines[p] = layout::find_text_lines(words[p]);
    } else {
      std::vector<std::thread> pool;
      std::vector<std::exception_ptr> errs(pages.size());
      std::atomic<size_t> next{0};
      for (size_t t = 0; t < nt; ++t)
        pool.emplace_back([&] {
          for (size_t p = next.fetch_add(1); p < pages.size(); p = next.fetch_add(1)) {
            try {
              lines[p] = layout::find_text_lines(words[p]);
            } catch (...) {
              errs[p] = std::current_exception();
            }
          }
        });
      for (auto& t : pool) t.join();
      for (auto& e : errs)
        if (e) std::rethrow_exception(e);
    }
  }
  return recognize_text(pages, lines);
}

}  // namespace ocrs
