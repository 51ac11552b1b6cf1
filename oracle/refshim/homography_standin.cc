// homography_standin.cc — stand-in for the three out-of-line functions of the reference's
// stitch/homography.cc, which needs Eigen (absent here).  TEST INFRASTRUCTURE ONLY.  It exists
// so that the reference's own stitch/transform_estimate.cc links: the RANSAC inlier test
// (TransformEstimation::get_inliers, transform_estimate.cc:132-148) is then the reference's
// code; it uses none of these three (they serve calc_transform / the overlap filter).
#include <cstdio>
#include <cstdlib>
#include "stitch/homography.hh"
#include "lib/matrix.hh"
#include "../small_linalg.h"

namespace pano {

Homography Homography::inverse(bool* succ) const {
  Homography ret;
  const bool ok = orc_lu3_inverse(data, ret.data) != 0;
  if (succ) *succ = ok;
  else if (!ok) { fprintf(stderr, "homography_standin: singular matrix\n"); abort(); }
  return ret;
}

Homography Homography::operator * (const Homography& r) const {
  Homography ret;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = data[i * 3] * r.data[j];
      acc += data[i * 3 + 1] * r.data[3 + j];
      acc += data[i * 3 + 2] * r.data[6 + j];
      ret.data[i * 3 + j] = acc;
    }
  return ret;
}

std::vector<Vec2D> overlap_region(const Shape2D&, const Shape2D&, const Matrix&, const Homography&) {
  fprintf(stderr, "homography_standin: overlap_region is off the checked path\n");
  abort();
}

}  // namespace pano
