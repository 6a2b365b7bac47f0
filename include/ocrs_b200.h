/* ocrs_b200 -- C ABI of the B200-native OCR hot path.
 *
 * Drop-in boundary for robertknight/ocrs @ 4bccf6b (v0.12.2).  Two seams are exported:
 *
 *  (1) the inner seam = the reference's `Model` trait (ocrs/src/model.rs:6-17), which
 *      `TextDetector` / `TextRecognizer` box as `dyn Model` (detection.rs:67, recognition.rs:316):
 *        ocrs_b200_model_load         <- rten::Model::load_file / ModelOptions::load
 *                                        (ocrs-cli/src/models.rs:105, ocrs/src/wasm_api.rs:62-64)
 *        ocrs_b200_model_input_shape  <- Model::input_shape   (model.rs:9; rten impl :20-31)
 *        ocrs_b200_model_run          <- Model::run           (model.rs:12-16; rten impl :33-40)
 *
 *  (2) the outer seam = the public `OcrEngine` methods (ocrs/src/lib.rs:130-300), needed to keep
 *      pages resident in HBM between stages and to batch pages:
 *        ocrs_b200_engine_create               <- OcrEngine::new                 (lib.rs:132)
 *        ocrs_b200_engine_prepare_input        <- OcrEngine::prepare_input       (lib.rs:183)
 *                                                 + ImageSource::from_bytes/from_tensor
 *                                                 (preprocess.rs:81,105)
 *        ocrs_b200_engine_detect_words         <- OcrEngine::detect_words        (lib.rs:193)
 *        ocrs_b200_engine_detect_text_pixels   <- OcrEngine::detect_text_pixels  (lib.rs:207)
 *        ocrs_b200_engine_find_text_lines      <- OcrEngine::find_text_lines     (lib.rs:222)
 *        ocrs_b200_engine_recognize_text       <- OcrEngine::recognize_text      (lib.rs:237)
 *        ocrs_b200_engine_prepare_recognition_input <- OcrEngine::prepare_recognition_input (lib.rs:268)
 *        ocrs_b200_engine_detection_threshold  <- OcrEngine::detection_threshold (lib.rs:282)
 *        ocrs_b200_engine_get_text             <- OcrEngine::get_text            (lib.rs:290)
 *        ocrs_b200_engine_ocr_batch            <- get_text over a batch of pages (new: configs 2-5)
 *
 * All pointers are plain host pointers unless a function says otherwise.  Functions return 0 on
 * success or a negative ocrs_b200_status; the message is available from ocrs_b200_last_error()
 * (thread-local).  Nothing throws or unwinds across this boundary.  There is no CPU fallback:
 * without a CUDA device every entry point that needs one fails with OCRS_B200_ERR_NO_DEVICE.
 */
#ifndef OCRS_B200_H_
#define OCRS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ocrs_b200_status {
  OCRS_B200_OK = 0,
  OCRS_B200_ERR_INVALID_ARG = -1,
  OCRS_B200_ERR_UNSUPPORTED_CHANNEL_COUNT = -2, /* ImageSourceError::UnsupportedChannelCount (preprocess.rs:41) */
  OCRS_B200_ERR_INVALID_DATA_LENGTH = -3,       /* ImageSourceError::InvalidDataLength (preprocess.rs:44) */
  OCRS_B200_ERR_MODEL_NOT_LOADED = -4,          /* "Detection/Recognition model not loaded" (lib.rs:197,254) */
  OCRS_B200_ERR_MODEL_LOAD = -5,
  OCRS_B200_ERR_RUN_FAILED = -6,                /* ModelRunError::RunFailed (errors.rs:8) */
  OCRS_B200_ERR_WRONG_OUTPUT = -7,              /* ModelRunError::WrongOutput (errors.rs:11) */
  OCRS_B200_ERR_CUDA = -8,
  OCRS_B200_ERR_NO_DEVICE = -9,
  OCRS_B200_ERR_INTERNAL = -10
} ocrs_b200_status;

typedef struct ocrs_b200_model ocrs_b200_model;   /* impl Model (model.rs) */
typedef struct ocrs_b200_engine ocrs_b200_engine; /* OcrEngine (lib.rs:111) */
typedef struct ocrs_b200_input ocrs_b200_input;   /* OcrInput (lib.rs:125), device resident */

/* rten_imageproc::RotatedRect: centre, unit up axis, width (perpendicular to up), height. */
typedef struct ocrs_b200_rotated_rect {
  float cx, cy, ux, uy, w, h;
} ocrs_b200_rotated_rect;

/* rten_imageproc::Rect<i32> */
typedef struct ocrs_b200_rect {
  int32_t top, left, bottom, right;
} ocrs_b200_rect;

/* Vec<Option<TextLine>> flattened (text_items.rs:47-66).  Line i is `None` when
 * line_present[i] == 0; its characters are [char_offsets[i], char_offsets[i+1]). */
typedef struct ocrs_b200_text_result {
  int32_t n_lines;
  uint8_t* line_present;      /* [n_lines] */
  int64_t* char_offsets;      /* [n_lines + 1] */
  uint32_t* chars;            /* Unicode scalar values */
  ocrs_b200_rect* char_rects; /* TextChar::rect */
} ocrs_b200_text_result;

enum { OCRS_B200_DTYPE_U8 = 0, OCRS_B200_DTYPE_F32 = 1 };        /* ImagePixels (preprocess.rs:9-14) */
enum { OCRS_B200_ORDER_HWC = 0, OCRS_B200_ORDER_CHW = 1 };       /* DimOrder (preprocess.rs:50-57) */
enum { OCRS_B200_DECODE_GREEDY = 0, OCRS_B200_DECODE_BEAM = 1 }; /* DecodeMethod (recognition.rs:199-205) */

/* OcrEngineParams (lib.rs:37-71).  Model buffers hold an .onnx file image; NULL = not loaded. */
typedef struct ocrs_b200_engine_params {
  const uint8_t* detection_model;
  size_t detection_model_len;
  const uint8_t* recognition_model;
  size_t recognition_model_len;
  int32_t debug;
  int32_t decode_method;
  uint32_t beam_width;
  const char* alphabet_utf8;      /* NULL = DEFAULT_ALPHABET (lib.rs:34) */
  const char* allowed_chars_utf8; /* NULL = every character allowed */
  int32_t device;
} ocrs_b200_engine_params;

/* ---- general -------------------------------------------------------------------------------- */
const char* ocrs_b200_last_error(void);
int ocrs_b200_device_count(void);
void ocrs_b200_free(void* p);
const char* ocrs_b200_version(void);

/* ---- inner seam: trait Model ---------------------------------------------------------------- */
int ocrs_b200_model_load(const uint8_t* bytes, size_t len, int device, ocrs_b200_model** out);
int ocrs_b200_model_load_file(const char* path, int device, ocrs_b200_model** out);
/* dims[i] = -1 for a symbolic dimension (rten::Dimension::Symbolic). */
int ocrs_b200_model_input_shape(const ocrs_b200_model* m, int64_t dims[8], int* ndim);
/* Thread-safe on one handle (recognition.rs:465-485 calls run() from rayon workers).
 * *out is malloc'ed by the library: release with ocrs_b200_free. */
int ocrs_b200_model_run(const ocrs_b200_model* m, const float* in, const int64_t* in_shape, int in_ndim, float** out,
                        int64_t out_shape[8], int* out_ndim);
/* Parses a model file WITHOUT touching a GPU (host code) and describes it as JSON:
 * {"format","opset","inputs":[{"name","dims"}],"outputs":[...],"nodes":n,"ops":{op:count},
 *  "unsupported_ops":[...],"initializers":n,"initializer_bytes":n}; dims: -1 = symbolic.  What
 * `rten::Model::load` would reject (ocrs-cli/src/models.rs:105) is reported here as
 * OCRS_B200_ERR_MODEL_LOAD with a message; a malformed file never crashes the caller.
 * *json is malloc'ed, NUL terminated. */
int ocrs_b200_model_inspect(const uint8_t* bytes, size_t len, char** json);
/* FLOPs executed by the last run on this handle (2*MACs of Conv/ConvTranspose/MatMul/GRU). */
double ocrs_b200_model_last_flops(const ocrs_b200_model* m);
void ocrs_b200_model_destroy(ocrs_b200_model* m);

/* ---- outer seam: OcrEngine ------------------------------------------------------------------ */
int ocrs_b200_engine_create(const ocrs_b200_engine_params* params, ocrs_b200_engine** out);
void ocrs_b200_engine_destroy(ocrs_b200_engine* e);

/* ImageSource::from_bytes: `len` bytes of HWC u8; channels = len / (width*height). */
int ocrs_b200_engine_prepare_input_bytes(ocrs_b200_engine* e, const uint8_t* bytes, size_t len, uint32_t width,
                                         uint32_t height, ocrs_b200_input** out);
/* ImageSource::from_tensor: u8 or f32, HWC or CHW, 1/3/4 channels. */
int ocrs_b200_engine_prepare_input(ocrs_b200_engine* e, const void* pixels, int dtype, int order, int height,
                                   int width, int channels, ocrs_b200_input** out);
/* Same, but `pixels` already lives in device memory of the engine's GPU. */
int ocrs_b200_engine_prepare_input_device(ocrs_b200_engine* e, const void* device_pixels, int dtype, int order,
                                          int height, int width, int channels, ocrs_b200_input** out);
int ocrs_b200_input_shape(const ocrs_b200_input* in, int* height, int* width);
/* Copies the [1,H,W] greyscale page to `out` (H*W floats). */
int ocrs_b200_input_read(ocrs_b200_engine* e, const ocrs_b200_input* in, float* out);
void ocrs_b200_input_destroy(ocrs_b200_input* in);

/* out: H*W floats. */
int ocrs_b200_engine_detect_text_pixels(ocrs_b200_engine* e, const ocrs_b200_input* in, float* out);
/* *rects is malloc'ed (ocrs_b200_free).  Order = contour discovery order, as in the reference. */
int ocrs_b200_engine_detect_words(ocrs_b200_engine* e, const ocrs_b200_input* in, ocrs_b200_rotated_rect** rects,
                                  size_t* n);
/* lines: *out_words holds the words regrouped line by line; line i = [offsets[i], offsets[i+1]). */
int ocrs_b200_engine_find_text_lines(ocrs_b200_engine* e, const ocrs_b200_input* in,
                                     const ocrs_b200_rotated_rect* words, size_t n_words,
                                     ocrs_b200_rotated_rect** out_words, size_t** line_offsets, size_t* n_lines);
/* Same computation without an engine handle (pure host code; usable without a GPU). */
int ocrs_b200_find_text_lines(const ocrs_b200_rotated_rect* words, size_t n_words, ocrs_b200_rotated_rect** out_words,
                              size_t** line_offsets, size_t* n_lines);
/* ---- text items (ocrs/src/text_items.rs) and CLI output formats (ocrs-cli/src/output.rs) -------
 * Pure host code, usable without a GPU.
 * `TextItem::rotated_rect` (text_items.rs:18-30): min-area rect of all corners of the items'
 * character rects, oriented towards "up" = (x 0, y -1).  rects: n >= 1 character boxes. */
int ocrs_b200_text_item_rotated_rect(const ocrs_b200_rect* char_rects, size_t n, ocrs_b200_rotated_rect* out);
/* `RotatedRect::corners` rounded like `rounded_vertex_coords` (output.rs:24-27): xy[2*i], xy[2*i+1]
 * = round-half-away(x), round-half-away(y) of corner i. */
int ocrs_b200_rotated_rect_vertices(const ocrs_b200_rotated_rect* r, int32_t xy[8]);
/* `format_text_output` (output.rs:87-94) and `format_json_output` (output.rs:97-100, HierText-style
 * document of output.rs:34-76) of one page's recognition result.  *utf8 is malloc'ed, NUL terminated.
 * The JSON is pretty-printed with 2-space indentation like serde_json::to_string_pretty. */
int ocrs_b200_format_text_output(const ocrs_b200_text_result* lines, char** utf8);
int ocrs_b200_format_json_output(const ocrs_b200_text_result* lines, const char* input_path, int image_height,
                                 int image_width, char** utf8);

int ocrs_b200_engine_recognize_text(ocrs_b200_engine* e, const ocrs_b200_input* in,
                                    const ocrs_b200_rotated_rect* words, const size_t* line_offsets, size_t n_lines,
                                    ocrs_b200_text_result** out);
void ocrs_b200_text_result_free(ocrs_b200_text_result* r);
/* *out is malloc'ed [*out_h, *out_w] floats. */
int ocrs_b200_engine_prepare_recognition_input(ocrs_b200_engine* e, const ocrs_b200_input* in,
                                               const ocrs_b200_rotated_rect* line_words, size_t n_words, float** out,
                                               int* out_h, int* out_w);
float ocrs_b200_engine_detection_threshold(const ocrs_b200_engine* e);
/* *utf8 is malloc'ed, NUL terminated; lines joined with '\n' (lib.rs:290-300). */
int ocrs_b200_engine_get_text(ocrs_b200_engine* e, const ocrs_b200_input* in, char** utf8);

/* Batched pipeline over resident pages: detect -> layout -> recognise for every page.
 * results[i] is a malloc'ed ocrs_b200_text_result for page i (ocrs_b200_text_result_free). */
int ocrs_b200_engine_ocr_batch(ocrs_b200_engine* e, const ocrs_b200_input* const* inputs, size_t n_pages,
                               ocrs_b200_text_result** results);
/* Same pipeline, results as text only: texts[i] is a malloc'ed NUL-terminated UTF-8 string, the
 * recognised lines of page i joined with '\n' (== OcrEngine::get_text, lib.rs:290-300). */
int ocrs_b200_engine_ocr_batch_text(ocrs_b200_engine* e, const ocrs_b200_input* const* inputs, size_t n_pages,
                                    char** texts);
/* Batched detect_words: rects of page i are (*rects)[offsets[i] .. offsets[i+1]). */
int ocrs_b200_engine_detect_words_batch(ocrs_b200_engine* e, const ocrs_b200_input* const* inputs, size_t n_pages,
                                        ocrs_b200_rotated_rect** rects, size_t** offsets);

/* ---- debug outputs of ocrs-cli, batched on the device ------------------
 * `--text-map` / `--text-mask` (ocrs-cli/src/main.rs:423-436: detect_text_pixels, then x > detection_threshold)
 * for a batch of pages in ONE detection pass.  maps / masks: caller arrays of n_pages entries (either may be
 * NULL = not wanted), filled with malloc'ed buffers of H_i*W_i floats / bytes (1 = text). */
int ocrs_b200_engine_detect_text_pixels_batch(ocrs_b200_engine* e, const ocrs_b200_input* const* inputs, size_t n_pages,
                                              float** maps, uint8_t** masks);
/* `--text-line-images` (main.rs:441-443, write_preprocessed_text_line_images): the recognition inputs
 * (prepare_recognition_input, lib.rs:268) of ALL lines of a page in one crop launch.  *images is one malloc'ed
 * buffer; line i is a [*height, (*widths)[i]] row-major image at float offset (*offsets)[i].  *widths and
 * *offsets are malloc'ed arrays of n_lines entries. */
int ocrs_b200_engine_prepare_recognition_inputs(ocrs_b200_engine* e, const ocrs_b200_input* in,
                                                const ocrs_b200_rotated_rect* words, const size_t* line_offsets,
                                                size_t n_lines, float** images, int* height, int** widths,
                                                size_t** offsets);

/* ---- engine pool: pipelining and multi-GPU fan-out inside the library ---------------------------
 * The reference runs get_text page by page on the calling thread (ocrs/src/lib.rs:290-300,
 * ocrs-cli/src/main.rs:438-446).  A pool owns, per device, `in_flight` worker threads with one engine
 * each; batches of pages are submitted asynchronously and collected by ticket, so the host phases of
 * one batch (layout analysis, result assembly) overlap the kernels of another, and all GPUs of a box
 * are fed from one process (SURVEY.md section 8b "device_ids[]", section 8e).  Workers are pinned to
 * the CPUs of their GPU's NUMA node. */
typedef struct ocrs_b200_pool ocrs_b200_pool;

/* ImageSource (preprocess.rs:61-124).  `pixels` is borrowed until the batch's ticket has been waited
 * for.  on_device != 0: `pixels` is device memory of one of the pool's GPUs (the batch runs there). */
typedef struct ocrs_b200_page {
  const void* pixels;
  int32_t dtype, order; /* OCRS_B200_DTYPE_*, OCRS_B200_ORDER_* */
  int32_t height, width, channels;
  int32_t on_device;
} ocrs_b200_page;

typedef struct ocrs_b200_pool_params {
  ocrs_b200_engine_params engine; /* models, decode method, alphabet ...; `device` is ignored */
  const int32_t* device_ids;      /* NULL = every visible device */
  int32_t n_devices;
  int32_t in_flight;              /* batches in flight (worker threads, engines) per device; 0 = 2 */
  int32_t pin_numa;               /* 0 = default (pin), 1 = pin, -1 = do not pin */
  int32_t layout_threads;         /* host threads per worker for layout analysis; 0 = 4 */
} ocrs_b200_pool_params;

int ocrs_b200_pool_create(const ocrs_b200_pool_params* params, ocrs_b200_pool** out);
void ocrs_b200_pool_destroy(ocrs_b200_pool* p); /* finishes queued batches first */
/* Enqueues one batch (detect -> layout -> recognise of every page); never blocks on the GPU. */
int ocrs_b200_pool_submit(ocrs_b200_pool* p, const ocrs_b200_page* pages, size_t n_pages, uint64_t* ticket);
/* Blocks until the batch is done.  results / texts: caller arrays of n_pages entries, filled with
 * malloc'ed objects (ocrs_b200_text_result_free / ocrs_b200_free).  A ticket is consumed by its wait. */
int ocrs_b200_pool_wait(ocrs_b200_pool* p, uint64_t ticket, ocrs_b200_text_result** results, size_t n_pages);
int ocrs_b200_pool_wait_text(ocrs_b200_pool* p, uint64_t ticket, char** texts, size_t n_pages);
int ocrs_b200_pool_done(ocrs_b200_pool* p, uint64_t ticket, int* done);
int ocrs_b200_pool_shape(const ocrs_b200_pool* p, int* n_devices, int* in_flight);
/* Worker engine `slot` of device index `dev_index` as an engine handle (statistics / profiling hooks);
 * release with ocrs_b200_engine_destroy. */
int ocrs_b200_pool_engine(ocrs_b200_pool* p, int dev_index, int slot, ocrs_b200_engine** out);
/* Human-readable worker placement (device, PCI bus id, NUMA node, CPUs); *text is malloc'ed. */
int ocrs_b200_pool_describe(ocrs_b200_pool* p, char** text);

/* Counters accumulated since the last reset: [0] detection FLOPs, [1] recognition FLOPs,
 * [2] words, [3] lines, [4] CTC timesteps, [5] recognition batches. */
int ocrs_b200_engine_stats(ocrs_b200_engine* e, double out[8], int reset);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------ */
/* CUDA-event profiling of pipeline stages and of every operator of the two networks. */
int ocrs_b200_engine_set_profiling(ocrs_b200_engine* e, int enable);
/* *json is malloc'ed: {"stage/...": {"ms", "calls", "launches", "flops", "bytes"}, "rec/Conv": ...} */
int ocrs_b200_engine_profile_json(ocrs_b200_engine* e, char** json, int reset);
/* cudaEvent timer on the engine's stream: start records, stop records + waits + returns ms. */
int ocrs_b200_engine_timer_start(ocrs_b200_engine* e);
int ocrs_b200_engine_timer_stop(ocrs_b200_engine* e, float* ms);
/* Host<->device traffic of the engine since creation: out[0] = H2D bytes, out[1] = D2H bytes. */
int ocrs_b200_engine_transfer_bytes(ocrs_b200_engine* e, int64_t out[2]);
/* Kernels launched by the library in this process so far. */
int64_t ocrs_b200_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OCRS_B200_H_ */
