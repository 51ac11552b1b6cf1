// Host check of openpano_b200/csrc/desc_interval.h: for random keypoints the column
// intervals must contain every position feature/sift.cc:107-124 accepts (brute force over
// the whole window with the reference's float expressions), and should contain little else.
//   g++ -O2 -ffp-contract=off -o /tmp/dic tools/probes/desc_interval_check.cc -lm && /tmp/dic
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../../openpano_b200/csrc/desc_interval.h"

static bool ref_accepts(int xx, int yy, int radius, int px, int py, int w, int h, float sinort, float cosort, float hist_w) {
  int nowx = px + xx, nowy = py + yy;
  if (!(nowx >= 1 && nowx <= w - 2)) return false;
  if (!(nowy >= 1 && nowy <= h - 2)) return false;
  if (xx * xx + yy * yy > radius * radius) return false;
  float y_rot = (-xx * sinort + yy * cosort) / hist_w, x_rot = (xx * cosort + yy * sinort) / hist_w;
  float ybin = y_rot + 4 / 2 - 0.5, xbin = x_rot + 4 / 2 - 0.5;
  if (!(ybin >= -1 && ybin <= 3) || !(xbin >= -1 && xbin <= 3)) return false;
  return true;
}

int main(int argc, char** argv) {
  const long trials = argc > 1 ? atol(argv[1]) : 200000;
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  long long accepted = 0, listed = 0, window = 0, missing = 0;
  for (long it = 0; it < trials; ++it) {
    float sf = 1.0f + 11.0f * U(rng) * U(rng);
    float hist_w = sf * 3.0f;
    int radius = (int)round(M_SQRT1_2 * (double)hist_w * 5);
    if (radius > 127) { --it; continue; }
    int w = 8 + (int)(U(rng) * 400), h = 8 + (int)(U(rng) * 400);
    int px = (int)(U(rng) * w), py = (int)(U(rng) * h);
    float ort;
    switch (it % 8) {            // stress the axis-aligned cases
      case 0: ort = 0.f; break;
      case 1: ort = (float)M_PI_2; break;
      case 2: ort = (float)M_PI; break;
      case 3: ort = (float)(1.5 * M_PI); break;
      case 4: ort = (float)M_PI_2 + (U(rng) - 0.5f) * 4e-3f; break;
      case 5: ort = (float)M_PI + (U(rng) - 0.5f) * 4e-3f; break;
      default: ort = U(rng) * (float)(2 * M_PI);
    }
    float sinort = sinf(ort), cosort = cosf(ort);
    const float lo = -2.5f * hist_w - 0.02f * hist_w - 1e-3f, hi = 1.5f * hist_w + 0.02f * hist_w + 1e-3f;
    for (int xx = -radius; xx <= radius; ++xx) {
      int y0, y1;
      desc_col_interval(xx, radius, px, py, w, h, sinort, cosort, lo, hi, &y0, &y1);
      if (y1 >= y0) {
        if (y0 < -radius || y1 > radius) { printf("interval outside the window: %d %d r=%d\n", y0, y1, radius); return 1; }
        listed += y1 - y0 + 1;
      }
      for (int yy = -radius; yy <= radius; ++yy) {
        ++window;
        if (ref_accepts(xx, yy, radius, px, py, w, h, sinort, cosort, hist_w)) {
          ++accepted;
          if (!(y1 >= y0 && yy >= y0 && yy <= y1)) {
            if (missing < 10) printf("MISSING xx=%d yy=%d [%d,%d] r=%d ort=%.9g hw=%.9g p=(%d,%d) wh=(%d,%d)\n", xx, yy, y0, y1, radius, ort, hist_w, px, py, w, h);
            ++missing;
          }
        }
      }
    }
  }
  printf("trials %ld  window %lld  accepted %lld  listed %lld  (listed/accepted %.4f, window/accepted %.3f)  missing %lld\n",
         trials, window, accepted, listed, (double)listed / accepted, (double)window / accepted, missing);
  return missing ? 1 : 0;
}
