// pano_host.hh — the C++ host side of the drop-in: adaptor classes that a maintainer adds to
// the reference tree.  They derive from / mirror the reference's own seams
//   FeatureDetector      feature/feature.hh:42-52
//   PairWiseMatcher      feature/matcher.hh:40-67  (same constructor shape and match(i, j))
//   BlenderBase          stitch/blender.hh:14-59
//   CylinderWarper       stitch/warp.hh:41-66
// and forward to the C ABI of libpano_b200.so (include/pano_b200.h).  Header-only; compiles
// against the reference's headers (-I <reference>/src) with the reference's own flags.
// tests/adaptor/adaptor_test.cc builds these against the reference tree and checks them
// against the reference classes they replace, bit for bit.
#pragma once
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include "pano_b200.h"

#include "lib/config.hh"
#include "lib/mat.h"
#include "lib/geometry.hh"
#include "lib/debugutils.hh"
#include "feature/feature.hh"
#include "feature/matcher.hh"
#include "stitch/blender.hh"
#include "stitch/warp.hh"
#include "stitch/homography.hh"

namespace pano_b200 {

// lib/config.hh:24-68 -> the POD snapshot every engine call takes
inline pano_params snapshot_params() {
  using namespace config;
  pano_params p;
  pano_params_default(&p);
  p.sift_working_size = SIFT_WORKING_SIZE; p.num_octave = NUM_OCTAVE; p.num_scale = NUM_SCALE;
  p.scale_factor = SCALE_FACTOR; p.gauss_sigma = GAUSS_SIGMA; p.gauss_window_factor = GAUSS_WINDOW_FACTOR;
  p.judge_extrema_diff_thres = JUDGE_EXTREMA_DIFF_THRES; p.contrast_thres = CONTRAST_THRES;
  p.pre_color_thres = PRE_COLOR_THRES; p.edge_ratio = EDGE_RATIO; p.calc_offset_depth = CALC_OFFSET_DEPTH;
  p.offset_thres = OFFSET_THRES; p.ori_radius = ORI_RADIUS; p.ori_hist_smooth_count = ORI_HIST_SMOOTH_COUNT;
  p.desc_hist_scale_factor = DESC_HIST_SCALE_FACTOR; p.desc_int_factor = DESC_INT_FACTOR;
  p.match_reject_next_ratio = MATCH_REJECT_NEXT_RATIO; p.focal_length = FOCAL_LENGTH;
  p.ordered_input = ORDERED_INPUT; p.lazy_read = LAZY_READ; p.multiband = MULTIBAND;
  p.max_output_size = MAX_OUTPUT_SIZE;
  return p;
}

// One engine context per process (main.cc creates it after init_config()).
class Context {
 public:
  explicit Context(int device = 0) {
    if (pano_create(&ctx_, device, nullptr) != 0) error_exit(pano_last_error(nullptr));
  }
  ~Context() { pano_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  pano_ctx* get() const { return ctx_; }
  void check(int rc) const { if (rc != 0) error_exit(pano_last_error(ctx_)); }   // the reference's error convention
 private:
  pano_ctx* ctx_ = nullptr;
};

// ---- features: a FeatureDetector (feature.hh:42-52).  do_detect_feature returns real_coor in
// [0,1) exactly as SIFTDetector::do_detect_feature does; the non-virtual detect_feature of the
// base class then applies its own scaling (feature.cc:20-28).
class B200SIFTDetector : public pano::FeatureDetector {
 public:
  explicit B200SIFTDetector(const Context& c) : c_(c) {}
  std::vector<pano::Descriptor> do_detect_feature(const Mat32f& img) const override {
    pano_params p = snapshot_params();
    pano_featureset* fs = nullptr;
    c_.check(pano_sift_detect(c_.get(), img.ptr(), img.width(), img.height(), &p, &fs));
    auto out = unpack(fs, 0, /*real=*/true);
    pano_featureset_free(fs);
    return out;
  }
  // calc_feature()'s loop over images (stitcherbase.cc:14-17) as ONE batched call; descriptors
  // stay on the device in *keep for B200PairMatcher.  Coordinates come back already scaled
  // ((c - 0.5) * w), like detect_feature's.
  std::vector<std::vector<pano::Descriptor>> detect_batch(const std::vector<const Mat32f*>& imgs,
                                                          pano_featureset** keep = nullptr) const {
    const int n = (int)imgs.size();
    std::vector<const float*> ptr(n);
    std::vector<int> w(n), h(n);
    for (int k = 0; k < n; ++k) { ptr[k] = imgs[k]->ptr(); w[k] = imgs[k]->width(); h[k] = imgs[k]->height(); }
    pano_params p = snapshot_params();
    pano_featureset* fs = nullptr;
    c_.check(pano_sift_detect_batch(c_.get(), n, ptr.data(), w.data(), h.data(), &p, &fs));
    std::vector<std::vector<pano::Descriptor>> feats(n);
    for (int k = 0; k < n; ++k) {
      feats[k] = unpack(fs, k, /*real=*/false);
      if (feats[k].empty()) error_exit(ssprintf("Cannot find feature in image %d!\n", k));   // stitcherbase.cc:20-21
    }
    if (keep) *keep = fs; else pano_featureset_free(fs);
    return feats;
  }
 private:
  std::vector<pano::Descriptor> unpack(pano_featureset* fs, int k, bool real) const {
    const int m = pano_featureset_count(fs, k);
    if (m < 0) error_exit(pano_last_error(c_.get()));
    std::vector<double> xy(2 * (size_t)m + 2);
    std::vector<float> d(128 * (size_t)m + 1);
    if (m) {
      c_.check(pano_featureset_download(fs, k, real ? nullptr : xy.data(), d.data()));
      if (real) c_.check(pano_featureset_download_real(fs, k, xy.data()));
    }
    std::vector<pano::Descriptor> out(m);
    for (int i = 0; i < m; ++i) {
      out[i].coor = Vec2D(xy[2 * i], xy[2 * i + 1]);
      out[i].descriptor.assign(d.begin() + 128 * (size_t)i, d.begin() + 128 * (size_t)(i + 1));
    }
    return out;
  }
  const Context& c_;
};

// ---- matching: PairWiseMatcher's shape (matcher.hh:40-67) with the exact rule of
// FeatureMatcher::match (matcher.cc:15-71).
class B200PairMatcher {
 public:
  B200PairMatcher(const Context& c, pano_featureset* from_detect) : c_(c), fs_(from_detect), owned_(false) {}
  B200PairMatcher(const Context& c, const std::vector<std::vector<pano::Descriptor>>& feats) : c_(c), owned_(true) {
    const int n = (int)feats.size();
    std::vector<int> cnt(n);
    std::vector<std::vector<float>> buf(n);
    std::vector<const float*> ptr(n);
    for (int i = 0; i < n; ++i) {
      cnt[i] = (int)feats[i].size();
      buf[i].resize(128 * (size_t)cnt[i] + 1);
      for (int k = 0; k < cnt[i]; ++k) memcpy(&buf[i][128 * (size_t)k], feats[i][k].descriptor.data(), 512);
      ptr[i] = buf[i].data();
    }
    c_.check(pano_featureset_upload(c_.get(), n, cnt.data(), ptr.data(), nullptr, &fs_));
  }
  ~B200PairMatcher() { if (owned_) pano_featureset_free(fs_); }
  B200PairMatcher(const B200PairMatcher&) = delete;
  B200PairMatcher& operator=(const B200PairMatcher&) = delete;

  // every task of pairwise_match() / linear_pairwise_match() (stitcher.cc:96-136) in one call
  std::vector<pano::MatchData> match_all(const std::vector<std::pair<int, int>>& tasks) const {
    std::vector<int> ij(2 * tasks.size() + 2);
    for (size_t t = 0; t < tasks.size(); ++t) { ij[2 * t] = tasks[t].first; ij[2 * t + 1] = tasks[t].second; }
    pano_params p = snapshot_params();
    pano_matches m;
    c_.check(pano_match_pairs(c_.get(), fs_, (int)tasks.size(), ij.data(), &p, &m));
    std::vector<pano::MatchData> out(tasks.size());
    for (size_t t = 0; t < tasks.size(); ++t)
      for (int q = 0; q < m.count[t]; ++q)
        out[t].data.emplace_back(m.idx[2 * (m.offset[t] + q)], m.idx[2 * (m.offset[t] + q) + 1]);
    pano_matches_free(&m);
    return out;
  }
  pano::MatchData match(int i, int j) const { return match_all({{i, j}})[0]; }   // = pwmatcher.match(i, j)
 private:
  const Context& c_;
  pano_featureset* fs_ = nullptr;
  bool owned_;
};

// ---- blend: a BlenderBase (blender.hh:14-59).  The std::function the reference passes cannot
// cross a C ABI; it is always the closed form of stitcher_image.cc:142-151, so the caller hands
// over what that lambda closes over (homo_inv) through add_image(), and the projection through
// the constructor.  bands == 0: LinearBlender; bands > 0: MultiBandBlender{bands}.
class B200Blender : public pano::BlenderBase {
 public:
  B200Blender(const Context& c, int bands, int projection, Vec2D resolution, Vec2D proj_min)
      : c_(c), bands_(bands) {
    g_.projection = projection; g_.res_x = resolution.x; g_.res_y = resolution.y;
    g_.proj_min_x = proj_min.x; g_.proj_min_y = proj_min.y;
  }
  void add_image(const Coor& upper_left, const Coor& bottom_right, pano::ImageRef& img, const pano::Homography& homo_inv) {
    img.load();
    pano_blend_image b;
    b.rgb_hwc = img.img->ptr(); b.w = img.width(); b.h = img.height();
    b.x0 = upper_left.x; b.y0 = upper_left.y; b.x1 = bottom_right.x; b.y1 = bottom_right.y;
    memcpy(b.homo_inv, homo_inv.data, sizeof(double) * 9);
    imgs_.push_back(b);
  }
  // the reference signature: only usable when the closure's parameters were handed over first
  void add_image(const Coor&, const Coor&, pano::ImageRef&, std::function<Vec2D(Coor)>) override {
    error_exit("B200Blender: pass the homography (add_image(ul, br, img, homo_inv)), a closure cannot cross the C ABI");
  }
  Mat32f run() override {
    int ow = 0, oh = 0;
    c_.check(pano_blend_target_size((int)imgs_.size(), imgs_.data(), &ow, &oh));
    Mat32f out(oh, ow, 3);
    pano_params p = snapshot_params();
    c_.check(pano_blend(c_.get(), (int)imgs_.size(), imgs_.data(), &g_, bands_, &p, out.ptr(), ow, oh));
    return out;
  }
 private:
  const Context& c_;
  int bands_;
  pano_blend_geom g_;
  std::vector<pano_blend_image> imgs_;
};

// ---- cylinder warp: CylinderWarper(h_factor).warp(mat, kpts) (warp.hh:41-66)
class B200CylinderWarper {
 public:
  B200CylinderWarper(const Context& c, real_t h_factor) : c_(c), h_factor_(h_factor) {}
  void warp(Mat32f& mat, std::vector<Vec2D>& kpts) const {
    pano_params p = snapshot_params();
    int ow, oh;
    double ox, oy;
    c_.check(pano_cyl_warp_shape(mat.width(), mat.height(), h_factor_, &p, &ow, &oh, &ox, &oy));
    Mat32f out(oh, ow, 3);
    std::vector<double> xy(2 * kpts.size() + 2);
    for (size_t i = 0; i < kpts.size(); ++i) { xy[2 * i] = kpts[i].x; xy[2 * i + 1] = kpts[i].y; }
    c_.check(pano_cyl_warp(c_.get(), mat.ptr(), mat.width(), mat.height(), h_factor_, &p, out.ptr(), ow, oh, xy.data(),
                           (int)kpts.size()));
    for (size_t i = 0; i < kpts.size(); ++i) kpts[i] = Vec2D(xy[2 * i], xy[2 * i + 1]);
    mat = out;
  }
 private:
  const Context& c_;
  real_t h_factor_;
};

// ---- Stitcher::build()'s hot path (stitcher.cc:32-64) on the engine: calc_feature ->
// pairwise / linear match -> blend.  The geometry in between (RANSAC, camera estimation,
// bundle adjustment: host code that stays the reference's) is supplied by the caller as the
// per-image ranges and inverse homographies ConnectedImages would hold.
struct StitchGeometry {
  int projection = PANO_PROJ_FLAT;
  Vec2D resolution{1, 1}, proj_min{0, 0};
  std::vector<Coor> upper_left, bottom_right;
  std::vector<pano::Homography> homo_inv;
};

class B200Stitcher {
 public:
  explicit B200Stitcher(const Context& c) : c_(c), det_(c) {}
  // returns the mosaic; feats / matches are left for the host geometry
  Mat32f build(std::vector<pano::ImageRef>& imgs, const StitchGeometry& geo) {
    const int n = (int)imgs.size();
    std::vector<const Mat32f*> mats(n);
    for (int k = 0; k < n; ++k) { imgs[k].load(); mats[k] = imgs[k].img; }
    pano_featureset* fs = nullptr;
    feats = det_.detect_batch(mats, &fs);                                     // calc_feature()
    std::vector<std::pair<int, int>> tasks;
    if (config::ORDERED_INPUT) for (int i = 0; i < n; ++i) tasks.emplace_back(i, (i + 1) % n);       // stitcher.cc:121-122
    else for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) tasks.emplace_back(i, j);       // stitcher.cc:98-100
    {
      B200PairMatcher pm(c_, fs);
      matches = pm.match_all(tasks);
    }
    pano_featureset_free(fs);
    pairs = tasks;
    B200Blender bl(c_, config::MULTIBAND, geo.projection, geo.resolution, geo.proj_min);              // stitcher_image.cc:132-136
    for (int k = 0; k < n; ++k) bl.add_image(geo.upper_left[k], geo.bottom_right[k], imgs[k], geo.homo_inv[k]);
    return bl.run();
  }
  std::vector<std::vector<pano::Descriptor>> feats;
  std::vector<std::pair<int, int>> pairs;
  std::vector<pano::MatchData> matches;
 private:
  const Context& c_;
  B200SIFTDetector det_;
};

}  // namespace pano_b200
