// Column intervals of the descriptor window (feature/sift.cc:107-124).
//
// calc_descriptor scans xx outer / yy inner over the (2r+1)^2 window and keeps the
// positions that are (i) inside the image interior, (ii) inside the circle of radius r and
// (iii) inside the rotated 5x5-cell box (xbin, ybin in [-1, 3]).  All three sets are convex,
// so for a fixed column xx the accepted yy form ONE interval.  desc_col_interval() returns a
// SUPERSET [y0, y1] of that interval from float arithmetic with explicit slack; the kernel
// enumerates only those positions (in scan order) and applies the reference's exact tests to
// each, so the accepted set — and with it every float sum — is unchanged while ~2/3 of the
// window is never visited.
//
// Plain C/C++ so that tools/probes/desc_interval_check.cc can run the same code on the
// host against a brute-force scan (gcc -ffp-contract=off == nvcc --fmad=false).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define DI_FN __host__ __device__ __forceinline__
#else
#define DI_FN static inline
#endif

// lo <= c + k*fy <= hi  ->  [*a, *b] narrowed (integers, inclusive).  The caller's lo/hi
// already carry the margin that covers the reference's float rounding (0.02*hist_w + 1e-3,
// at least 0.06 — five hundred times the worst rounding error of c + k*fy, in the same
// units, whatever k is); the 0.05 here covers the division and the int conversion.
DI_FN void di_apply_linear(float k, float c, float lo, float hi, int* a, int* b) {
  if (fabsf(k) < 1e-3f) {
    // |k*fy| <= 1e-3 * 181 < 0.2: the constraint is (almost) independent of fy
    if (c < lo - 0.5f || c > hi + 0.5f) { *a = 1; *b = 0; }
    return;
  }
  const float inv = 1.0f / k;
  const float e0 = (lo - c) * inv, e1 = (hi - c) * inv;
  float mn = fminf(e0, e1) - 0.05f, mx = fmaxf(e0, e1) + 0.05f;
  mn = fmaxf(mn, -300.f);
  mx = fminf(mx, 300.f);
  const int ia = (int)ceilf(mn), ib = (int)floorf(mx);
  if (ia > *a) *a = ia;
  if (ib < *b) *b = ib;
}

// Superset of { yy : calc_descriptor accepts (xx, yy) }.  Empty when *y1 < *y0.
// px,py: keypoint (octave) coordinates; w,h: octave size; lo/hi: conservative bounds on the
// un-normalised rotated coordinates (see k_descriptor).
DI_FN void desc_col_interval(int xx, int radius, int px, int py, int w, int h, float sinort, float cosort,
                             float lo, float hi, int* y0, int* y1) {
  const int nowx = px + xx;
  if (nowx < 1 || nowx > w - 2) { *y0 = 1; *y1 = 0; return; }
  const float fx = (float)xx;
  const float t = (float)(radius * radius) - fx * fx;         // exact small integers
  const int ym = (int)sqrtf(t < 0.f ? 0.f : t);               // floor(sqrt(t)): correctly rounded sqrtf never reaches the next integer for t < 2^24
  int a = -ym, b = ym;
  if (1 - py > a) a = 1 - py;
  if (h - 2 - py < b) b = h - 2 - py;
  di_apply_linear(cosort, (float)(-xx) * sinort, lo, hi, &a, &b);   // y_rot numerator
  di_apply_linear(sinort, fx * cosort, lo, hi, &a, &b);             // x_rot numerator
  *y0 = a;
  *y1 = b;
}
