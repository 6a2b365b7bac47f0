/* Plain-C consumer of include/ocrs_b200.h: what a non-Python host (the reference's Rust through
 * `extern "C"`) sees.  Built by tests/test_c_abi.py with gcc, linked against
 * ocrs_b200/libocrs_b200.so only (no CUDA headers, no torch).
 *
 *   abi_smoke host                    host-only entry points (no GPU needed)
 *   abi_smoke gpu <det.onnx> <rec.onnx>   full pipeline on one synthetic page, prints the text
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ocrs_b200.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != OCRS_B200_OK) {                                                        \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, ocrs_b200_last_error());    \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

static int host_only(void) {
  printf("version %s\n", ocrs_b200_version());

  /* layout_analysis.rs:158: two rows of three words each -> two lines, left to right */
  ocrs_b200_rotated_rect words[6];
  for (int i = 0; i < 6; ++i) {
    int row = i / 3, col = i % 3;
    words[i].cx = 30.0f + 60.0f * (float)(2 - col); /* given right-to-left */
    words[i].cy = 20.0f + 40.0f * (float)row;
    words[i].ux = 0.0f;
    words[i].uy = 1.0f;
    words[i].w = 50.0f;
    words[i].h = 20.0f;
  }
  ocrs_b200_rotated_rect* out_words = NULL;
  size_t* offs = NULL;
  size_t n_lines = 0;
  CHECK(ocrs_b200_find_text_lines(words, 6, &out_words, &offs, &n_lines));
  printf("lines %zu:", n_lines);
  for (size_t l = 0; l < n_lines; ++l) {
    printf(" [");
    for (size_t k = offs[l]; k < offs[l + 1]; ++k) printf("%s%.0f", k > offs[l] ? " " : "", out_words[k].cx);
    printf("]");
  }
  printf("\n");
  ocrs_b200_free(out_words);
  ocrs_b200_free(offs);

  /* output.rs:218-236 with one line "hi yo" of 10-px characters */
  const char* text = "hi yo";
  uint8_t present[1] = {1};
  int64_t char_offs[2] = {0, 5};
  uint32_t chars[5];
  ocrs_b200_rect rects[5];
  for (int i = 0; i < 5; ++i) {
    chars[i] = (uint32_t)text[i];
    rects[i].top = 0;
    rects[i].left = 10 * i;
    rects[i].bottom = 25;
    rects[i].right = 10 * i + 10;
  }
  ocrs_b200_text_result res = {1, present, char_offs, chars, rects};
  char* s = NULL;
  CHECK(ocrs_b200_format_text_output(&res, &s));
  printf("text <%s>\n", s);
  ocrs_b200_free(s);
  CHECK(ocrs_b200_format_json_output(&res, "page.png", 25, 50, &s));
  printf("json bytes %zu first-line %.1s\n", strlen(s), s);
  ocrs_b200_free(s);

  ocrs_b200_rotated_rect rr;
  int32_t xy[8];
  CHECK(ocrs_b200_text_item_rotated_rect(rects, 5, &rr));
  CHECK(ocrs_b200_rotated_rect_vertices(&rr, xy));
  printf("vertices %d,%d %d,%d %d,%d %d,%d\n", xy[0], xy[1], xy[2], xy[3], xy[4], xy[5], xy[6], xy[7]);

  /* errors never unwind across the boundary: a bad argument is a code plus a message */
  int rc = ocrs_b200_text_item_rotated_rect(rects, 0, &rr);
  printf("empty item -> %d (%s)\n", rc, rc ? ocrs_b200_last_error() : "ok");
  return 0;
}

static unsigned char* read_file(const char* path, size_t* len) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  unsigned char* buf = (unsigned char*)malloc((size_t)n);
  if (fread(buf, 1, (size_t)n, f) != (size_t)n) {
    fclose(f);
    free(buf);
    return NULL;
  }
  fclose(f);
  *len = (size_t)n;
  return buf;
}

static int gpu(const char* det_path, const char* rec_path, const char* page_path, int w, int h) {
  size_t det_len = 0, rec_len = 0, page_len = 0;
  unsigned char* det = read_file(det_path, &det_len);
  unsigned char* rec = read_file(rec_path, &rec_len);
  unsigned char* page = read_file(page_path, &page_len);
  if (!det || !rec || !page) {
    fprintf(stderr, "cannot read inputs\n");
    return 1;
  }
  ocrs_b200_engine_params p;
  memset(&p, 0, sizeof p);
  p.detection_model = det;
  p.detection_model_len = det_len;
  p.recognition_model = rec;
  p.recognition_model_len = rec_len;
  p.decode_method = OCRS_B200_DECODE_GREEDY;
  p.beam_width = 100;
  ocrs_b200_engine* e = NULL;
  CHECK(ocrs_b200_engine_create(&p, &e));
  ocrs_b200_input* in = NULL;
  CHECK(ocrs_b200_engine_prepare_input_bytes(e, page, page_len, (uint32_t)w, (uint32_t)h, &in)); /* ImageSource::from_bytes */
  char* text = NULL;
  CHECK(ocrs_b200_engine_get_text(e, in, &text)); /* lib.rs:290 */
  printf("%s\n", text);
  ocrs_b200_free(text);
  ocrs_b200_input_destroy(in);
  ocrs_b200_engine_destroy(e);
  free(det);
  free(rec);
  free(page);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && strcmp(argv[1], "host") == 0) return host_only();
  if (argc >= 7 && strcmp(argv[1], "gpu") == 0) return gpu(argv[2], argv[3], argv[4], atoi(argv[5]), atoi(argv[6]));
  fprintf(stderr, "usage: abi_smoke host | gpu <det.onnx> <rec.onnx> <page.rgb> <width> <height>\n");
  return 2;
}
