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

/* imgproc.cc:200-235.  Line by line: height[k] = run of pixels with
 * max(r,g,b) >= 0 ending at this line; left[k] / right[k] = furthest columns
 * reachable through heights >= height[k] (path-compressed walks, :213-222);
 * the first strictly larger (right-left+1)*height wins (:223-225). */
int orc_crop(const float* mat, int w, int h, int* rect, float* out_hwc) {
  if (!mat || w <= 0 || h <= 0) return -1;
  int* height = (int*)calloc((size_t)w, sizeof(int));
  int* left = (int*)malloc(sizeof(int) * (size_t)w);
  int* right = (int*)malloc(sizeof(int) * (size_t)w);
  int maxarea = 0, ll = 0, rr = 0, hh = 0, nl = 0;
  for (int line = 0; line < h; ++line) {
    for (int k = 0; k < w; ++k) {
      const float* p = mat + ((size_t)line * w + k) * 3;
      float m01 = (p[0] < p[1]) ? p[1] : p[0];                    /* std::max */
      float m = (m01 < p[2]) ? p[2] : m01;
      height[k] = m < 0 ? 0 : height[k] + 1;
    }
    for (int k = 0; k < w; ++k) {
      left[k] = k;
      while (left[k] > 0 && height[k] <= height[left[k] - 1]) left[k] = left[left[k] - 1];
    }
    for (int k = w - 1; k >= 0; --k) {
      right[k] = k;
      while (right[k] < w - 1 && height[k] <= height[right[k] + 1]) right[k] = right[right[k] + 1];
    }
    for (int k = 0; k < w; ++k) {
      int area = (right[k] - left[k] + 1) * height[k];
      if (area > maxarea) {                                       /* update_max, utils.hh */
        maxarea = area;
        ll = left[k]; rr = right[k]; hh = height[k]; nl = line;
      }
    }
  }
  free(height); free(left); free(right);
  int cw = rr - ll + 1, ch = hh;
  int offx = ll, offy = nl - hh + 1;
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
