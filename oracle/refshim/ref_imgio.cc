// ref_imgio.cc — drives the REFERENCE's read_img / write_rgb / crop (lib/imgio.cc,
// lib/imgproc.cc, compiled unmodified by oracle/Makefile) through lossless PNM
// files, so the 8-bit boundary of the checker is pinned against the reference's
// own code.  TEST INFRASTRUCTURE ONLY; contains no algorithm.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "lib/mat.h"
#include "lib/imgproc.hh"
#include "../oracle_api.h"

using namespace pano;

namespace {

std::string temp_name(const char* suffix) {
  char buf[64];
  snprintf(buf, sizeof buf, "/tmp/pano_ref_XXXXXX%s", suffix);
  int fd = mkstemps(buf, (int)strlen(suffix));
  if (fd >= 0) close(fd);
  return buf;
}

// binary PNM reader for what CImg::save_pnm wrote (header tokens, '#' comments)
bool read_pnm(const char* fname, int* w, int* h, int* ch, std::vector<unsigned char>* data) {
  FILE* f = fopen(fname, "rb");
  if (!f) return false;
  char magic[3] = {0, 0, 0};
  if (fscanf(f, "%2s", magic) != 1) { fclose(f); return false; }
  *ch = (magic[1] == '6') ? 3 : 1;
  int vals[3], got = 0;
  while (got < 3) {
    int c = fgetc(f);
    if (c == EOF) { fclose(f); return false; }
    if (c == '#') { while (c != '\n' && c != EOF) c = fgetc(f); continue; }
    if (c == ' ' || c == '\n' || c == '\r' || c == '\t') continue;
    ungetc(c, f);
    if (fscanf(f, "%d", &vals[got]) != 1) { fclose(f); return false; }
    ++got;
  }
  fgetc(f);   // the single whitespace after maxval
  *w = vals[0]; *h = vals[1];
  data->resize((size_t)*w * *h * *ch);
  size_t rd = fread(data->data(), 1, data->size(), f);
  fclose(f);
  return rd == data->size() && vals[2] == 255;
}

}  // namespace

extern "C" {

int ref_read_img_rgb8(const unsigned char* pix, int w, int h, int channels, float* out) {
  if (channels != 1 && channels != 3) return -1;
  std::string fname = temp_name(channels == 3 ? ".ppm" : ".pgm");
  FILE* f = fopen(fname.c_str(), "wb");
  if (!f) return -1;
  fprintf(f, "P%d\n%d %d\n255\n", channels == 3 ? 6 : 5, w, h);
  fwrite(pix, 1, (size_t)w * h * channels, f);
  fclose(f);
  Mat32f m = read_img(fname.c_str());
  unlink(fname.c_str());
  if (m.width() != w || m.height() != h || m.channels() != 3) return -1;
  memcpy(out, m.ptr(), sizeof(float) * 3 * (size_t)w * h);
  return 0;
}

int ref_write_rgb8(const float* mat, int w, int h, unsigned char* out) {
  Mat32f m(h, w, 3);
  memcpy(m.ptr(), mat, sizeof(float) * 3 * (size_t)w * h);
  std::string fname = temp_name(".ppm");
  write_rgb(fname.c_str(), m);
  int rw, rh, rc;
  std::vector<unsigned char> data;
  bool ok = read_pnm(fname.c_str(), &rw, &rh, &rc, &data);
  unlink(fname.c_str());
  if (!ok || rw != w || rh != h || rc != 3) return -1;
  memcpy(out, data.data(), data.size());
  return 0;
}

int ref_crop(const float* mat, int w, int h, int* rect, float* out) {
  Mat32f m(h, w, 3);
  memcpy(m.ptr(), mat, sizeof(float) * 3 * (size_t)w * h);
  Mat32f r = crop(m);
  if (rect) { rect[0] = -1; rect[1] = -1; rect[2] = r.width(); rect[3] = r.height(); }
  if (out && r.height() > 0) memcpy(out, r.ptr(), sizeof(float) * 3 * (size_t)r.width() * r.height());
  return 0;
}

}  // extern "C"
