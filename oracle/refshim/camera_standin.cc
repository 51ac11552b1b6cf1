// camera_standin.cc — stand-in for the out-of-line members of the reference's stitch/camera.cc
// that its incremental_bundle_adjuster.cc links against.  camera.cc needs Eigen (JacobiSVD in
// rotation_to_angle and straighten), which is absent here.  TEST INFRASTRUCTURE ONLY: it exists
// so that the reference's own calcJacobianSymbolic (incremental_bundle_adjuster.cc:276-385)
// runs in the checker.  Unpinned step, stated: Camera::rotation_to_angle first replaces R by the
// nearest rotation U V^T of its SVD (camera.cc:92-98); this stand-in takes R as it is, which
// differs by rounding for the orthonormal R the tests feed.  The angle only enters dRdvi, a
// PER-CAMERA INPUT of the per-point Jacobian code under test — checker, oracle and CUDA kernel
// all receive the same matrices.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "stitch/camera.hh"
#include "stitch/match_info.hh"

namespace pano {

Camera::Camera() : R(Homography::I()) {}

Homography Camera::K() const {               // camera.cc:58-65
  Homography ret{Homography::I()};
  ret[0] = focal;
  ret[2] = ppx;
  ret[4] = focal * aspect;
  ret[5] = ppy;
  return ret;
}

void Camera::rotation_to_angle(const Homography& r, double& rx, double& ry, double& rz) {   // camera.cc:91-117 without the SVD polish
  rx = r.data[7] - r.data[5];
  ry = r.data[2] - r.data[6];
  rz = r.data[3] - r.data[1];
  double s = sqrt(rx * rx + ry * ry + rz * rz);
  if (s < GEO_EPS) {
    rx = ry = rz = 0;
  } else {
    double c = (r.data[0] + r.data[4] + r.data[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    double mul = 1.0 / s * theta;
    rx *= mul; ry *= mul; rz *= mul;
  }
}

void Camera::angle_to_rotation(double, double, double, Homography&) {
  fprintf(stderr, "camera_standin: angle_to_rotation is off the checked path\n");
  abort();
}

double Camera::estimate_focal(const std::vector<std::vector<MatchInfo>>&) {
  fprintf(stderr, "camera_standin: estimate_focal is off the checked path\n");
  abort();
}

void Camera::straighten(std::vector<Camera>&) {
  fprintf(stderr, "camera_standin: straighten is off the checked path\n");
  abort();
}

}  // namespace pano
