// ref_ba.cc — the reference's own symbolic bundle-adjustment Jacobian behind the checker API.
// TEST INFRASTRUCTURE ONLY.  The reference TU is compiled WHERE IT LIES by including it (nothing
// is copied): that makes IncrementalBundleAdjuster::calcJacobianSymbolic callable and the
// file-local dRdvi / dKd* visible for the per-pair matrices, which are evaluated here with the
// reference's own Homography / Camera operations exactly as the loop at
// incremental_bundle_adjuster.cc:288-352 spells them.
#include <vector>
#include <set>
#include <map>
#include <array>
#include <memory>
#include <cmath>
#include <iostream>
#include <sstream>
#include <string>
#include <algorithm>
#include <limits>
#define private public            // calcJacobianSymbolic, J, JtJ, match_pairs ... are private members
#define protected public
#include "stitch/incremental_bundle_adjuster.cc"
#undef private
#undef protected
#include "../oracle_api.h"

using namespace pano;

static std::vector<Camera> make_cameras(int n_cam, const double* cams) {
  std::vector<Camera> cs(n_cam);
  for (int i = 0; i < n_cam; ++i) {
    const double* c = cams + 12 * i;
    cs[i].focal = c[0]; cs[i].ppx = c[1]; cs[i].ppy = c[2]; cs[i].aspect = 1;
    for (int k = 0; k < 9; ++k) cs[i].R.data[k] = c[3 + k];
  }
  return cs;
}

extern "C" int ref_ba_pair_mats(int n_cam, const double* cams, int n_pair, orc_ba_pair* pairs) {
  std::vector<Camera> cameras = make_cameras(n_cam, cams);
  std::vector<std::array<Homography, 3>> all_dRdvi(cameras.size());
  for (int i = 0; i < n_cam; ++i) all_dRdvi[i] = dRdvi(cameras[i].R);          // :284-286
  for (int p = 0; p < n_pair; ++p) {
    const Camera &c_from = cameras[pairs[p].from], &c_to = cameras[pairs[p].to];
    const auto fromK = c_from.K();                                              // :297-302
    const auto toKinv = c_to.Kinv();
    const auto toRinv = c_to.Rinv();
    const auto& dRfromdvi = all_dRdvi[pairs[p].from];
    auto dRtodviT = all_dRdvi[pairs[p].to];
    for (auto& m : dRtodviT) m = m.transpose();
    Homography out[13];
    out[0] = (fromK * c_from.R) * (toRinv * toKinv);                            // :304
    out[1] = c_from.R * toRinv * toKinv;                                        // :323
    out[2] = toRinv * toKinv;                                                   // :332
    for (int k = 0; k < 3; ++k) out[3 + k] = fromK * dRfromdvi[k];              // :333-335
    out[6] = toKinv;                                                            // :339, :349
    Homography m = fromK * c_from.R * toRinv * toKinv;                          // :338
    out[7] = m * dKdfocal; out[8] = m * dKdppx; out[9] = m * dKdppy;            // :341-345
    m = fromK * c_from.R;                                                       // :348
    for (int k = 0; k < 3; ++k) out[10 + k] = m * dRtodviT[k];                  // :350-352
    for (int q = 0; q < 13; ++q)
      for (int k = 0; k < 9; ++k) pairs[p].m[q][k] = out[q].data[k];
  }
  return 0;
}

extern "C" int ref_ba_jacobian(int n_cam, const double* cams, int n_pair, const orc_ba_pair* pairs, const double* pts,
                               double* j_rows, double* jtj) {
  std::vector<Camera> cameras = make_cameras(n_cam, cams);
  IncrementalBundleAdjuster ba(cameras);
  std::vector<MatchInfo> infos(n_pair);                 // MatchPair keeps a reference (incremental_bundle_adjuster.hh:55-60)
  for (int p = 0; p < n_pair; ++p) {
    for (int k = 0; k < pairs[p].n_match; ++k) {
      const double* q = pts + 4 * (size_t)(pairs[p].match_begin + k);
      infos[p].match.emplace_back(Vec2D(q[0], q[1]), Vec2D(q[2], q[3]));
    }
    ba.add_match(pairs[p].from, pairs[p].to, infos[p]);
  }
  // what optimize() does before the first get_param_update (:121-129)
  ba.update_index_map();
  const int nr_img = (int)ba.idx_added.size();
  if (nr_img != n_cam) return -1;                       // every camera must appear in a pair (slots == camera indices)
  ba.J = Eigen::MatrixXd{2 * ba.nr_pointwise_match, 6 * nr_img};
  ba.JtJ = Eigen::MatrixXd{6 * nr_img, 6 * nr_img};
  IncrementalBundleAdjuster::ParamState state;
  for (auto& idx : ba.idx_added) state.cameras.emplace_back(cameras[idx]);
  ba.calcJacobianSymbolic(state);
  const int N = 6 * nr_img;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) jtj[(size_t)i * N + j] = ba.JtJ(i, j);
  if (j_rows)
    for (int p = 0; p < n_pair; ++p) {
      const int pf = ba.index_map[pairs[p].from] * 6, pt = ba.index_map[pairs[p].to] * 6;
      for (int k = 0; k < pairs[p].n_match; ++k) {
        const int idx = 2 * (pairs[p].match_begin + k);
        double* r = j_rows + 24 * (size_t)(pairs[p].match_begin + k);
        for (int i = 0; i < 6; ++i) {
          r[i] = ba.J(idx, pf + i); r[6 + i] = ba.J(idx, pt + i);
          r[12 + i] = ba.J(idx + 1, pf + i); r[18 + i] = ba.J(idx + 1, pt + i);
        }
      }
    }
  return 0;
}
