// matrix_standin.cc — stand-in for the reference's lib/matrix.cc, whose only
// dependency (Eigen) is absent from this container.  TEST INFRASTRUCTURE ONLY.
// Implements the Matrix methods declared in lib/matrix.hh that the hot-path
// translation units link against (extrema.cc:134-148 uses inverse,
// pseudo_inverse and prod on 3x3 / 3x1 doubles).  See ../small_linalg.h for the
// algorithm and its "unpinned vs real Eigen" status.
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include "lib/matrix.hh"
#include "lib/geometry.hh"
#include "../small_linalg.h"

static void need3(const Matrix& m, const char* what) {
  if (m.rows() != 3 || m.cols() != 3) {
    fprintf(stderr, "matrix_standin: %s only implemented for 3x3 (got %dx%d)\n", what, m.rows(), m.cols());
    abort();
  }
}

std::ostream& operator << (std::ostream& os, const Matrix& m) {
  os << "[" << m.rows() << " " << m.cols() << "] :" << std::endl;
  for (int i = 0; i < m.rows(); ++i) for (int j = 0; j < m.cols(); ++j)
    os << m.at(i, j) << (j == m.cols() - 1 ? "\n" : ", ");
  return os;
}

Matrix Matrix::transpose() const {
  Matrix ret(m_cols, m_rows);
  for (int i = 0; i < m_rows; ++i) for (int j = 0; j < m_cols; ++j) ret.at(j, i) = at(i, j);
  return ret;
}

Matrix Matrix::prod(const Matrix& r) const {
  Matrix ret(m_rows, r.cols());
  for (int i = 0; i < m_rows; ++i)
    for (int j = 0; j < r.cols(); ++j) {
      double acc = at(i, 0) * r.at(0, j);
      for (int k = 1; k < m_cols; ++k) acc += at(i, k) * r.at(k, j);
      ret.at(i, j) = acc;
    }
  return ret;
}

bool Matrix::inverse(Matrix& ret) const {
  need3(*this, "inverse");
  ret = Matrix(3, 3);
  return orc_lu3_inverse(ptr(), ret.ptr()) != 0;
}

Matrix Matrix::pseudo_inverse() const {
  need3(*this, "pseudo_inverse");
  Matrix ret(3, 3);
  orc_sym3_pinv(ptr(), ret.ptr());
  return ret;
}

void Matrix::zero() { memset(ptr(), 0, sizeof(double) * pixels()); }

Matrix Matrix::I(int k) {
  Matrix ret(k, k); ret.zero();
  for (int i = 0; i < k; ++i) ret.at(i, i) = 1;
  return ret;
}
