/*
 * orc_imgio.c — CHECKER (test infrastructure only): plain-C restatement of the
 * byte formats either side of the hot path.
 *
 *   orc_read_img_rgb8   read_img's pixel conversion     lib/imgio.cc:67-90
 *   orc_crop            crop                             lib/imgproc.cc:200-235
 *   orc_write_rgb8      write_rgb's pixel conversion    lib/imgio.cc:98-113
 *
 * Pinned against the reference's own translation units through lossless PNM
 * files (tests/test_oracle_vs_ref.py; refshim/ref_shim.cc writes a P6/P5 file,
 * calls read_img / write_rgb on it and hands the result back).
 */
#include <stdlib.h>
#include <string.h>
#include "oracle_api.h"

/* imgio.cc:75-88.  Colour: each sample is `(float)v / 255.0` — the division is
 * done in double (255.0 is a double literal) and the quotient is rounded to float
 * on the store.  Grey: the raw value is replicated to the three channels with no
 * division (imgio.cc:84-87). */
int orc_read_img_rgb8(const unsigned char* pix, int w, int h, int channels, float* out_hwc) {
  if (!pix || !out_hwc || w <= 1 || h <= 1) return -1;         /* m_assert(rows > 1 && cols > 1), imgio.cc:89 */
  if (channels == 3) {
    size_t n = (size_t)w * h * 3;
    for (size_t i = 0; i < n; ++i) out_hwc[i] = (float)((double)(float)pix[i] / 255.0);
  } else if (channels == 1) {
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; ++i) {
      float v = (float)pix[i];
      out_hwc[i * 3] = out_hwc[i * 3 + 1] = out_hwc[i * 3 + 2] = v;
    }
  } else {
    return -1;                                                    /* m_assert(spectrum == 3 || == 1), imgio.cc:74 */
  }
  return 0;
}

/* imgproc.cc:200-235.  Line by line: run[k] = number of consecutive pixels with
 * max(r,g,b) >= 0 ending at this line in column k; lo[k] / hi[k] = furthest columns
 * reachable from k through runs >= run[k] (:213-222); the first strictly larger
 * (hi-lo+1)*run over (line, column) order wins (:223-225). */
int orc_crop(const float* mat, int w, int h, int* rect, float* out_hwc) {
  if (!mat || w <= 0 || h <= 0) return -1;
  int* run = (int*)calloc((size_t)w, sizeof(int));      /* valid pixels ending at this line, per column */
  int* lo = (int*)malloc(sizeof(int) * (size_t)w);      /* leftmost column reachable over runs >= run[k] */
  int* hi = (int*)malloc(sizeof(int) * (size_t)w);      /* rightmost such column */
  int best = 0, best_lo = 0, best_hi = 0, best_run = 0, best_line = 0;
  for (int line = 0; line < h; ++line) {
    const float* px = mat + (size_t)line * w * 3;
    for (int k = 0; k < w; ++k, px += 3) {
      float m01 = (px[0] < px[1]) ? px[1] : px[0];                /* std::max keeps its first argument on ties/NaN */
      float m = (m01 < px[2]) ? px[2] : m01;
      run[k] = m < 0 ? 0 : run[k] + 1;                            /* Color::NO ends the run */
    }
    /* spans by pointer jumping over already-resolved neighbours, left then right (:213-222) */
    for (int k = 0; k < w; ++k) {
      int l = k;
      while (l > 0 && run[k] <= run[l - 1]) l = lo[l - 1];
      lo[k] = l;
    }
    for (int k = w - 1; k >= 0; --k) {
      int r = k;
      while (r < w - 1 && run[k] <= run[r + 1]) r = hi[r + 1];
      hi[k] = r;
    }
    for (int k = 0; k < w; ++k) {
      int area = (hi[k] - lo[k] + 1) * run[k];
      if (area > best) {                                          /* update_max: strictly larger only */
        best = area;
        best_lo = lo[k]; best_hi = hi[k]; best_run = run[k]; best_line = line;
      }
    }
  }
  free(run); free(lo); free(hi);
  const int cw = best_hi - best_lo + 1, ch = best_run;
  const int offx = best_lo, offy = best_line - best_run + 1;
  if (rect) { rect[0] = offx; rect[1] = offy; rect[2] = cw; rect[3] = ch; }
  if (out_hwc)
    for (int i = 0; i < ch; ++i)
      memcpy(out_hwc + (size_t)i * cw * 3, mat + ((size_t)(i + offy) * w + offx) * 3, sizeof(float) * 3 * (size_t)cw);
  return 0;
}

/* imgio.cc:104-111: `(v < 0 ? 1 : v) * 255` is a float product, truncated by the
 * conversion to unsigned char (Color::NO = -1 becomes white). */
int orc_write_rgb8(const float* mat, int w, int h, unsigned char* out) {
  if (!mat || !out || w <= 0 || h <= 0) return -1;
  size_t n = (size_t)w * h * 3;
  for (size_t i = 0; i < n; ++i) {
    float v = mat[i];
    out[i] = (unsigned char)((v < 0 ? 1 : v) * 255);
  }
  return 0;
}
