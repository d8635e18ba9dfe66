// f4 wire format: PFM files (reference datasets/data_io.py:8-73 read_pfm / save_pfm) — the depth / confidence maps the eval
// script writes per reference view (eval_rcmvsnet_dtu.py:199-253) and the fusion step reads back.  Host code (no kernels):
// at >200 ref-views/s per GPU the Python writer (np.flipud + tofile per map) is the serial bottleneck; these routines
// stream the rows of a (pinned) host buffer straight to / from the file in bottom-up order.
// Format: "Pf\n" (1 channel) | "PF\n" (3 channels), "<width> <height>\n", "<scale>\n" (negative = little endian), raw
// float32 rows from the BOTTOM row up.
#include <errno.h>

#include "common.cuh"

namespace {
static bool host_little_endian() {
  const uint32_t one = 1;
  return *reinterpret_cast<const unsigned char*>(&one) == 1;
}
static int read_line(FILE* f, char* buf, int n) {
  int i = 0, c;
  while (i < n - 1 && (c = fgetc(f)) != EOF) {
    buf[i++] = (char)c;
    if (c == '\n') break;
  }
  buf[i] = 0;
  return i;
}
}  // namespace

extern "C" int rcmvs_pfm_read_header(const char* path, int* width, int* height, int* channels, float* scale) {
  using namespace rcmvs;
  RCMVS_REQUIRE(path && width && height && channels && scale, RCMVS_ERR_BAD_ARG, "rcmvs_pfm_read_header: null pointer");
  FILE* f = fopen(path, "rb");
  RCMVS_REQUIRE(f != nullptr, RCMVS_ERR_BAD_ARG, "rcmvs_pfm_read_header: cannot open %s: %s", path, strerror(errno));
  char line[128];
  read_line(f, line, sizeof line);
  int rc = RCMVS_OK;
  if (!strcmp(line, "PF\n") || !strcmp(line, "PF\r\n")) *channels = 3;
  else if (!strcmp(line, "Pf\n") || !strcmp(line, "Pf\r\n")) *channels = 1;
  else { set_error("rcmvs_pfm_read_header: %s: Not a PFM file.", path); rc = RCMVS_ERR_BAD_ARG; }
  if (rc == RCMVS_OK) {
    read_line(f, line, sizeof line);
    if (sscanf(line, "%d %d", width, height) != 2 || *width <= 0 || *height <= 0) {
      set_error("rcmvs_pfm_read_header: %s: Malformed PFM header.", path);
      rc = RCMVS_ERR_BAD_ARG;
    }
  }
  if (rc == RCMVS_OK) {
    read_line(f, line, sizeof line);
    if (sscanf(line, "%f", scale) != 1) { set_error("rcmvs_pfm_read_header: %s: bad scale line", path); rc = RCMVS_ERR_BAD_ARG; }
  }
  fclose(f);
  return rc;
}

// out: HOST buffer of height*width*channels floats, rows top-down (the reference flips after reading)
extern "C" int rcmvs_pfm_read(const char* path, float* out, int height, int width, int channels) {
  using namespace rcmvs;
  RCMVS_REQUIRE(path && out && height > 0 && width > 0 && (channels == 1 || channels == 3), RCMVS_ERR_BAD_ARG,
                "rcmvs_pfm_read: bad arguments");
  int w = 0, h = 0, c = 0;
  float scale = 0.f;
  const int rc = rcmvs_pfm_read_header(path, &w, &h, &c, &scale);
  if (rc != RCMVS_OK) return rc;
  RCMVS_REQUIRE(w == width && h == height && c == channels, RCMVS_ERR_BAD_ARG,
                "rcmvs_pfm_read: %s is %dx%dx%d, caller expects %dx%dx%d", path, h, w, c, height, width, channels);
  FILE* f = fopen(path, "rb");
  RCMVS_REQUIRE(f != nullptr, RCMVS_ERR_BAD_ARG, "rcmvs_pfm_read: cannot open %s", path);
  char line[128];
  for (int i = 0; i < 3; ++i) read_line(f, line, sizeof line);
  const size_t row = (size_t)width * channels;
  const bool swap = (scale < 0.f) != host_little_endian();
  int ok = 1;
  for (int y = height - 1; y >= 0 && ok; --y) {                 // file rows run bottom-up
    float* dst = out + (size_t)y * row;
    ok = fread(dst, sizeof(float), row, f) == row;
    if (ok && swap)
      for (size_t i = 0; i < row; ++i) {
        uint32_t v;
        memcpy(&v, dst + i, 4);
        v = (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
        memcpy(dst + i, &v, 4);
      }
  }
  fclose(f);
  RCMVS_REQUIRE(ok, RCMVS_ERR_BAD_ARG, "rcmvs_pfm_read: %s is truncated", path);
  return RCMVS_OK;
}

// img: HOST buffer [height][width][channels] float32, rows top-down; written like save_pfm(filename, image, scale)
extern "C" int rcmvs_pfm_write(const char* path, const float* img, int height, int width, int channels, float scale) {
  using namespace rcmvs;
  RCMVS_REQUIRE(path && img && height > 0 && width > 0 && (channels == 1 || channels == 3), RCMVS_ERR_BAD_ARG,
                "rcmvs_pfm_write: Image must have H x W x 3, H x W x 1 or H x W dimensions.");
  FILE* f = fopen(path, "wb");
  RCMVS_REQUIRE(f != nullptr, RCMVS_ERR_BAD_ARG, "rcmvs_pfm_write: cannot open %s: %s", path, strerror(errno));
  fprintf(f, channels == 3 ? "PF\n" : "Pf\n");
  fprintf(f, "%d %d\n", width, height);
  fprintf(f, "%f\n", host_little_endian() ? -scale : scale);
  const size_t row = (size_t)width * channels;
  int ok = 1;
  for (int y = height - 1; y >= 0 && ok; --y) ok = fwrite(img + (size_t)y * row, sizeof(float), row, f) == row;
  ok = (fclose(f) == 0) && ok;
  RCMVS_REQUIRE(ok, RCMVS_ERR_BAD_ARG, "rcmvs_pfm_write: short write to %s", path);
  return RCMVS_OK;
}
