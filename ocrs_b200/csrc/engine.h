// OcrEngine on one B200: the reference's public pipeline surface (ocrs/src/lib.rs:111-301) with
// device-resident pages and batched stages.
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"
#include "executor.h"
#include "geom.h"
#include "image_kernels.h"

namespace ocrs {

enum class DecodeMethod { kGreedy = 0, kBeamSearch = 1 };  // recognition.rs:199-205

struct EngineParams {  // lib.rs:37-71
  const uint8_t* detection_model = nullptr;
  size_t detection_model_len = 0;
  const uint8_t* recognition_model = nullptr;
  size_t recognition_model_len = 0;
  bool debug = false;
  DecodeMethod decode_method = DecodeMethod::kGreedy;
  uint32_t beam_width = 100;
  bool has_alphabet = false;
  std::string alphabet_utf8;
  bool has_allowed_chars = false;
  std::string allowed_chars_utf8;
  int device = 0;
};

// `OcrInput` (lib.rs:125-128): greyscale page in [-0.5, 0.5], resident in HBM.
struct OcrInput {
  std::shared_ptr<Storage> store;  // f32 [H*W], stream-ordered allocation on the engine's stream
  int H = 0, W = 0;
  int device = 0;
  float* grey() const { return reinterpret_cast<float*>(store->ptr); }
};

struct TextChar {  // text_items.rs:47-53
  uint32_t ch;     // Unicode scalar value
  geom::RectI rect;
};
struct TextLine {  // Option<TextLine>: present == false <-> None
  bool present = false;
  std::vector<TextChar> chars;
};

struct StageTimes {  // milliseconds, CUDA-event timed when `debug`
  float detect_ms = 0, recog_ms = 0;
};

class Engine {
 public:
  explicit Engine(const EngineParams& p);
  ~Engine();

  bool has_detector() const { return det_ != nullptr; }
  bool has_recognizer() const { return rec_ != nullptr; }
  int device() const { return device_; }
  float detection_threshold() const { return text_threshold_; }  // lib.rs:282-287
  const Model* detection_model() const { return det_.get(); }
  const Model* recognition_model() const { return rec_.get(); }

  // lib.rs:183 / preprocess.rs:149.  `pixels` is a host pointer (or a device pointer when
  // `pixels_on_device`).  dtype 0 = u8, 1 = f32; order 0 = HWC, 1 = CHW.
  std::unique_ptr<OcrInput> prepare_input(const void* pixels, int dtype, int order, int H, int W, int C,
                                          bool pixels_on_device = false);

  // prepare_input for a batch of pages; pages of one shape (u8 HWC RGB) are converted by ONE launch.
  struct PageSpec {
    const void* pixels = nullptr;
    int dtype = 0, order = 0, H = 0, W = 0, C = 0;
    bool on_device = false;
  };
  std::vector<std::unique_ptr<OcrInput>> prepare_inputs(const std::vector<PageSpec>& pages);

  // lib.rs:207 / detection.rs:131-200 (host copy of the H x W probability map)
  std::vector<float> detect_text_pixels(const OcrInput& in);
  // lib.rs:193 / detection.rs:104-122, batched over pages; rects in contour-discovery order.
  std::vector<std::vector<geom::RotatedRect>> detect_words(const std::vector<const OcrInput*>& pages);
  // lib.rs:222
  std::vector<std::vector<geom::RotatedRect>> find_text_lines(const std::vector<geom::RotatedRect>& words) const;
  // lib.rs:237 / recognition.rs:404-540, batched over pages.
  std::vector<std::vector<TextLine>> recognize_text(
      const std::vector<const OcrInput*>& pages,
      const std::vector<std::vector<std::vector<geom::RotatedRect>>>& lines_per_page);
  std::vector<std::vector<geom::RotatedRect>> detect_words_locked(const std::vector<const OcrInput*>& pages);  // caller holds mu_
  std::vector<std::vector<TextLine>> recognize_text_locked(
      const std::vector<const OcrInput*>& pages,
      const std::vector<std::vector<std::vector<geom::RotatedRect>>>& lines_per_page);  // caller holds mu_
  // lib.rs:268 / recognition.rs:366-393: returns [input_height, resized_width] row-major.
  std::vector<float> prepare_recognition_input(const OcrInput& in, const std::vector<geom::RotatedRect>& line,
                                               int* out_h, int* out_w);

  // Debug outputs of ocrs-cli (`--text-map`, `--text-mask`: main.rs:423-436) for a batch of pages in one
  // detection pass: per page the H x W probability map and / or the thresholded mask (x > threshold).
  struct TextPixels {
    std::vector<float> map;     // empty unless requested
    std::vector<uint8_t> mask;  // empty unless requested
    int H = 0, W = 0;
  };
  std::vector<TextPixels> detect_text_pixels_batch(const std::vector<const OcrInput*>& pages, bool want_map, bool want_mask);
  // `--text-line-images` (main.rs:441-443 -> write_preprocessed_text_line_images): the recognition inputs of ALL
  // lines of a page in one crop launch.  Line i is [height, widths[i]] row-major at offsets[i] of `images`.
  struct LineImages {
    std::vector<float> images;
    std::vector<int> widths;
    std::vector<size_t> offsets;
    int height = 0;
  };
  LineImages prepare_recognition_inputs(const OcrInput& in, const std::vector<std::vector<geom::RotatedRect>>& lines);

  // Whole pipeline on a batch of resident pages (detect -> layout -> recognise).
  std::vector<std::vector<TextLine>> ocr_pages(const std::vector<const OcrInput*>& pages);

  const std::vector<uint32_t>& alphabet() const { return alphabet_; }
  uint32_t rec_input_height() const;  // recognition.rs:332-337

  // statistics of the last recognize_text / detect_words call (for bench / roofline math)
  struct Stats {
    double det_flops = 0, rec_flops = 0;
    int64_t n_lines = 0, n_words = 0, n_timesteps = 0, rec_batches = 0;
    int64_t kernel_launches = 0;
  };
  Stats stats() const { return stats_; }
  void reset_stats() { stats_ = Stats(); }
  void synchronize();
  // host threads used by ocr_pages for layout analysis (pages of a batch in parallel); default 8
  void set_layout_threads(int n) { layout_threads_ = n < 1 ? 1 : n; }
  // host waits block on an event instead of spinning (the pool sets this for its workers)
  void set_blocking_sync(bool on) { blocking_sync_ = on; }

  // CUDA-event profiling of stages and of every operator of the two networks
  void set_profiling(bool on);
  std::string profile_json(bool reset);
  // Event timer on the engine's stream (bench: device-side timing of the timed region)
  void timer_start();
  float timer_stop();
  int64_t d2h_bytes() const { return d2h_bytes_; }
  int64_t h2d_bytes() const { return h2d_bytes_; }

 private:
  struct PageScratch;  // per concurrent page: mask, labels, pools
  PageScratch& scratch_for(int slot, int H, int W);

  int device_ = 0;
  cudaStream_t st_ = nullptr;
  std::unique_ptr<Model> det_, rec_;
  float text_threshold_ = 0.2f;  // detection.rs:34
  float min_area_ = 100.0f;      // detection.rs:31
  bool debug_ = false;
  DecodeMethod decode_method_ = DecodeMethod::kGreedy;
  uint32_t beam_width_ = 100;
  std::vector<uint32_t> alphabet_;
  bool has_excluded_ = false;
  std::vector<uint8_t> excluded_mask_;  // per class label (index 0 = blank)
  DeviceBuffer d_excluded_;
  std::vector<std::unique_ptr<PageScratch>> scratch_;
  DeviceBuffer det_in_, staging_, line_desc_, poly_, cross_, rec_batch_, ctc_scratch_, ctc_out_, page_tab_;
  DeviceBuffer tab_in_, tab_out_, tab_prep_, tab_ccl_;  // per-batch page tables of the batched pixel kernels
  PinnedBuffer h_pin_;
  std::mutex mu_;
  Stats stats_;
  Profiler prof_;
  cudaEvent_t ev_a_ = nullptr, ev_b_ = nullptr, ev_copy_ = nullptr;
  std::vector<cudaStream_t> aux_;       // per-page side streams for detection post-processing
  std::vector<cudaEvent_t> aux_done_;
  cudaEvent_t ev_fork_ = nullptr;
  cudaEvent_t ev_wait_ = nullptr;
  bool blocking_sync_ = false;
  void wait_stream();
  void ensure_aux(int n);
  int64_t d2h_bytes_ = 0, h2d_bytes_ = 0;
  int layout_threads_ = 8;
  std::map<std::string, std::pair<double, int64_t>> host_ms_;  // host-side section timers (ms, calls)
  std::mutex host_mu_;                                         // guards host_ms_ only
  struct HostTimer;
};

std::vector<uint32_t> utf8_to_codepoints(const std::string& s);
std::string codepoints_to_utf8(const std::vector<uint32_t>& cps);

}  // namespace ocrs
