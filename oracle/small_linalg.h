/*
 * small_linalg.h — 3x3 double solve used by the keypoint refinement.
 * TEST INFRASTRUCTURE (oracle side).
 *
 * The reference calls Matrix::inverse (lib/matrix.cc:76-86: Eigen FullPivLU,
 * isInvertible(), inverse()) and falls back to Matrix::pseudo_inverse
 * (lib/matrix.cc:89-106: JacobiSVD, singular values <= 1e-6 zeroed).  Eigen is a
 * system dependency of the reference (src/CMakeLists.txt:15, version unpinned)
 * and is absent from this container, so this is a restatement of Eigen 3.4's
 * published algorithm: right-looking LU with full pivoting, rank decided with
 * threshold eps*n relative to the largest pivot, inverse by solving against the
 * identity.  Parity of this one step against real Eigen is UNPINNED;
 * its results only feed round() and a <0.5 test plus sub-pixel offsets compared
 * at 1e-4.
 */
#ifndef ORACLE_SMALL_LINALG_H
#define ORACLE_SMALL_LINALG_H
#include <math.h>
#include <float.h>

/* a: row-major 3x3 (destroyed: becomes LU). Returns 1 and writes inv (row-major)
 * when invertible, 0 otherwise. */
static inline int orc_lu3_inverse(const double a_in[9], double inv[9]) {
  double a[9];
  int rowperm[3] = {0, 1, 2}; /* original row now at position i    */
  int colperm[3] = {0, 1, 2}; /* original column now at position j */
  int i, j, k, nonzero = 3, rank = 0, col;
  double maxpivot = 0.0, thr;
  for (i = 0; i < 9; ++i) a[i] = a_in[i];
  for (k = 0; k < 3; ++k) {
    int br = k, bc = k, t;
    double big = -1.0;
    /* biggest |coefficient| of the bottom-right corner; the scan order only
     * matters for exact ties */
    for (i = k; i < 3; ++i)
      for (j = k; j < 3; ++j) {
        double v = fabs(a[i * 3 + j]);
        if (v > big) { big = v; br = i; bc = j; }
      }
    if (big == 0.0) { nonzero = k; break; }
    if (big > maxpivot) maxpivot = big;
    if (br != k) {
      for (j = 0; j < 3; ++j) { double w = a[k * 3 + j]; a[k * 3 + j] = a[br * 3 + j]; a[br * 3 + j] = w; }
      t = rowperm[k]; rowperm[k] = rowperm[br]; rowperm[br] = t;
    }
    if (bc != k) {
      for (i = 0; i < 3; ++i) { double w = a[i * 3 + k]; a[i * 3 + k] = a[i * 3 + bc]; a[i * 3 + bc] = w; }
      t = colperm[k]; colperm[k] = colperm[bc]; colperm[bc] = t;
    }
    for (i = k + 1; i < 3; ++i) a[i * 3 + k] /= a[k * 3 + k];
    for (i = k + 1; i < 3; ++i)
      for (j = k + 1; j < 3; ++j)
        a[i * 3 + j] -= a[i * 3 + k] * a[k * 3 + j];
  }
  /* rank = pivots with |p| > eps * diagonalSize * |maxpivot| (FullPivLU::rank) */
  thr = DBL_EPSILON * 3.0 * maxpivot;
  for (k = 0; k < nonzero; ++k) if (fabs(a[k * 3 + k]) > thr) ++rank;
  if (rank < 3) return 0;
  /* inverse = solve(I).  P A Q = L U  =>  x = Q U^-1 L^-1 P b */
  for (col = 0; col < 3; ++col) {
    double c[3];
    for (i = 0; i < 3; ++i) c[i] = rowperm[i] == col ? 1.0 : 0.0;
    c[1] -= a[3] * c[0];
    c[2] -= a[6] * c[0];
    c[2] -= a[7] * c[1];
    c[2] /= a[8];
    c[1] -= a[5] * c[2];
    c[0] -= a[2] * c[2];
    c[1] /= a[4];
    c[0] -= a[1] * c[1];
    c[0] /= a[0];
    for (i = 0; i < 3; ++i) inv[colperm[i] * 3 + col] = c[i];
  }
  return 1;
}

/* Moore-Penrose pseudo-inverse of a SYMMETRIC 3x3 (the refinement Hessian is
 * symmetric) by cyclic Jacobi eigen-decomposition; eigenvalues with |l| <= 1e-6
 * are dropped (singular value = |eigenvalue|; lib/matrix.cc:97-101). */
static inline void orc_sym3_pinv(const double a_in[9], double out[9]) {
  double a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  int sweep, p, q, i, j, k;
  for (i = 0; i < 9; ++i) a[i] = a_in[i];
  for (sweep = 0; sweep < 32; ++sweep) {
    double off = fabs(a[1]) + fabs(a[2]) + fabs(a[5]);
    if (off < 1e-300) break;
    for (p = 0; p < 2; ++p)
      for (q = p + 1; q < 3; ++q) {
        double apq = a[p * 3 + q], theta, t, c, s;
        if (apq == 0.0) continue;
        theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
        t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        c = 1.0 / sqrt(t * t + 1.0);
        s = t * c;
        for (k = 0; k < 3; ++k) { /* A <- A J */
          double akp = a[k * 3 + p], akq = a[k * 3 + q];
          a[k * 3 + p] = c * akp - s * akq;
          a[k * 3 + q] = s * akp + c * akq;
        }
        for (k = 0; k < 3; ++k) { /* A <- J^T A */
          double apk = a[p * 3 + k], aqk = a[q * 3 + k];
          a[p * 3 + k] = c * apk - s * aqk;
          a[q * 3 + k] = s * apk + c * aqk;
        }
        for (k = 0; k < 3; ++k) { /* V <- V J */
          double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
          v[k * 3 + p] = c * vkp - s * vkq;
          v[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (k = 0; k < 3; ++k) {
        double l = a[k * 3 + k];
        if (fabs(l) > 1e-6) acc += v[i * 3 + k] * (1.0 / l) * v[j * 3 + k];
      }
      out[i * 3 + j] = acc;
    }
}

#endif
