// ba.cu — bundle-adjustment Jacobian assembly on the device.
//
// Replaces the per-point part of IncrementalBundleAdjuster::calcJacobianSymbolic
// (stitch/incremental_bundle_adjuster.cc:306-383): for every point match of every image pair
// the 2 x 12 block of d(residual)/d(camera parameters) (two rows of J, :355-361) and the
// running sums of J^T J (:363-382).  "J.rows() could reach 700000" (:280): the rows are
// independent, the J^T J entries are independent chains.  What stays on the host is the
// per-PAIR algebra in front of the loop (:288-304 and the loop-invariant 3x3 products inside
// it): those go through Homography::operator* / inverse and Camera::rotation_to_angle, which
// are Eigen calls in the reference — the caller evaluates them with its own Eigen and hands
// the 13 matrices per pair in (pano_ba_pair, include/pano_b200.h).
//
// Arithmetic is the reference's, operation for operation, in f64 without contraction:
// Homography::trans (homography.hh:52-57) sums its three products left to right;
// `sqr(homo.z)` is lib/utils.hh:25's FLOAT sqr (the double is narrowed first); drdv is the
// macro at :316-319.  Every J^T J entry is one sequential sum over pairs in list order and
// points in match order — exactly the order `JtJ(i1, i2) += val` runs in — so the result is
// bit-identical, not merely close.
#include "common.cuh"
#include <string.h>
#include <vector>

struct BaPairDev {
  int from, to, match_begin, n_match;
  double m[13][9];
};

struct BaVec { double x, y, z; };

__device__ __forceinline__ BaVec ba_trans(const double* __restrict__ d, BaVec v) {   // Homography::trans(const Vec&)
  BaVec r;
  r.x = d[0] * v.x + d[1] * v.y + d[2] * v.z;
  r.y = d[3] * v.x + d[4] * v.y + d[5] * v.z;
  r.z = d[6] * v.x + d[7] * v.y + d[8] * v.z;
  return r;
}

// the three constant matrices dK/dfocal, dK/dppx, dK/dppy (:84-95), applied with the same full products
__constant__ double c_dKd[3][9] = {{1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0},
                                   {0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
                                   {0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0}};

// pair matrices (order fixed by the header): 0 Hto_to_from; 1 R_from*toRinv*toKinv; 2 toRinv*toKinv;
// 3..5 fromK*dRfromdvi[k]; 6 toKinv; 7..9 m*dKd{focal,ppx,ppy}, m = fromK*R_from*toRinv*toKinv;
// 10..12 (fromK*R_from)*dRtodviT[k]
__global__ void __launch_bounds__(128)
k_ba_rows(const BaPairDev* __restrict__ pairs, const double2* __restrict__ pts_to, double* __restrict__ rows) {
  __shared__ double s_m[13][9];
  const BaPairDev& pr = pairs[blockIdx.y];
  for (int q = threadIdx.x; q < 13 * 9; q += blockDim.x) s_m[q / 9][q % 9] = pr.m[q / 9][q % 9];
  __syncthreads();
  const int n = pr.n_match, begin = pr.match_begin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double2 to = pts_to[begin + i];
    const BaVec tov{to.x, to.y, 1.0};                    // trans(Vec2D) = trans(Vec(x, y, 1))
    const BaVec homo = ba_trans(s_m[0], tov);
    const float hzf = (float)homo.z;                     // sqr(float), lib/utils.hh:25
    const double hz_sqr_inv = 1.0 / (double)(hzf * hzf);
    const double hz_inv = 1.0 / homo.z;
    double* out = rows + (size_t)(begin + i) * 24;       // row idx: dfrom.x[6], dto.x[6]; row idx+1: dfrom.y[6], dto.y[6]
    auto drdv = [&](BaVec dhdv, int col) {
      out[col] = -dhdv.x * hz_inv + dhdv.z * homo.x * hz_sqr_inv;
      out[12 + col] = -dhdv.y * hz_inv + dhdv.z * homo.y * hz_sqr_inv;
    };
    BaVec dot_u2 = ba_trans(s_m[1], tov);
#pragma unroll
    for (int k = 0; k < 3; ++k) drdv(ba_trans(c_dKd[k], dot_u2), k);            // dfrom: focal, ppx, ppy
    dot_u2 = ba_trans(s_m[2], tov);
#pragma unroll
    for (int k = 0; k < 3; ++k) drdv(ba_trans(s_m[3 + k], dot_u2), 3 + k);       // dfrom: rotation
    const BaVec ku = ba_trans(s_m[6], tov);
    dot_u2 = BaVec{ku.x * -1.0, ku.y * -1.0, ku.z * -1.0};                       // Vec::operator*(-1)
#pragma unroll
    for (int k = 0; k < 3; ++k) drdv(ba_trans(s_m[7 + k], dot_u2), 6 + k);       // dto: focal, ppx, ppy
#pragma unroll
    for (int k = 0; k < 3; ++k) drdv(ba_trans(s_m[10 + k], ku), 9 + k);          // dto: rotation
  }
}

// One CTA per 6x6 block (a, b) of J^T J, one thread per entry (i, j).  The pair list is walked in
// order; a pair contributes to this block as the cross term (from=a, to=b: dfrom[i].dto[j]), its
// mirror (from=b, to=a: JtJ(i2, i1) += dfrom[i'].dto[j'] with i'=j, j'=i), or a diagonal term
// (a == b == from: dfrom[i].dfrom[j]; a == b == to: dto[i].dto[j]).  Vec2D::dot = x*v.x + y*v.y.
__global__ void __launch_bounds__(64)
k_ba_jtj(const BaPairDev* __restrict__ pairs, int n_pair, int n_cam, const double* __restrict__ rows,
         double* __restrict__ jtj) {
  const int a = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
  if (t >= 36) return;
  const int i = t / 6, j = t % 6;
  double acc = 0.0;                                       // JtJ.setZero()
  for (int p = 0; p < n_pair; ++p) {
    const int from = pairs[p].from, to = pairs[p].to;
    int c0, c1;                                           // columns of the compact row whose 2-vectors are multiplied
    if (a == from && b == to) { c0 = i; c1 = 6 + j; }
    else if (a == to && b == from) { c0 = j; c1 = 6 + i; }
    else if (a == b && a == from) { c0 = i; c1 = j; }
    else if (a == b && a == to) { c0 = 6 + i; c1 = 6 + j; }
    else continue;
    const double* r = rows + (size_t)pairs[p].match_begin * 24;
    const int n = pairs[p].n_match;
    for (int k = 0; k < n; ++k, r += 24) {
      const double val = r[c0] * r[c1] + r[12 + c0] * r[12 + c1];
      acc += val;
    }
  }
  const size_t N = (size_t)n_cam * 6;
  jtj[(size_t)(a * 6 + i) * N + (size_t)b * 6 + j] = acc;
}

extern "C" int pano_ba_jacobian(pano_ctx* ctx, int n_cam, int n_pair, const pano_ba_pair* pairs, const double* pts_to,
                                double* j_rows, double* jtj) {
  ctx_enter(ctx);
  if (!ctx || n_cam <= 0 || n_pair < 0 || (n_pair && (!pairs || !pts_to)) || !jtj) return PANO_ERR_INVALID;
  static_assert(sizeof(BaPairDev) == sizeof(pano_ba_pair), "pano_ba_pair layout");
  long long nm = 0;
  int max_match = 0;
  for (int k = 0; k < n_pair; ++k) {
    const pano_ba_pair& p = pairs[k];
    if (p.from < 0 || p.from >= n_cam || p.to < 0 || p.to >= n_cam || p.from == p.to || p.n_match < 0 || p.match_begin != nm)
      return ctx_fail(ctx, PANO_ERR_INVALID, "ba: pair %d has a camera slot out of range or a match range that does not follow the previous pair's", k);
    nm += p.n_match;
    max_match = std::max(max_match, p.n_match);
  }
  const size_t N = (size_t)n_cam * 6;
  const size_t b_pairs = align_up((size_t)n_pair * sizeof(BaPairDev), 16), b_pts = (size_t)nm * 16;   // double2 loads behind the pair table
  const size_t b_rows = (size_t)nm * 24 * sizeof(double), b_jtj = N * N * sizeof(double);
  PANO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));            // the staging buffers may still feed earlier copies
  char* st = (char*)ctx_pinned(ctx, b_pairs + b_pts + 64);
  char* so = (char*)ctx_pinned2(ctx, (j_rows ? b_rows : 0) + b_jtj + 64);
  if (!st || !so) return ctx_fail(ctx, PANO_ERR_CUDA, "ba: pinned staging allocation failed");
  if (n_pair) memcpy(st, pairs, (size_t)n_pair * sizeof(BaPairDev));
  if (b_pts) memcpy(st + b_pairs, pts_to, b_pts);
  char* d_in = nullptr; char* d_out = nullptr;
  int rc = ctx_alloc(ctx, (void**)&d_in, b_pairs + b_pts + 64);
  if (!rc) rc = ctx_alloc(ctx, (void**)&d_out, b_rows + b_jtj + 64);
  if (rc) { ctx_free(ctx, d_in); ctx_free(ctx, d_out); return rc; }
  cudaError_t e = cudaMemcpyAsync(d_in, st, b_pairs + b_pts, cudaMemcpyHostToDevice, ctx->stream);
  const BaPairDev* d_pairs = (const BaPairDev*)d_in;
  const double2* d_pts = (const double2*)(d_in + b_pairs);
  double* d_rows = (double*)d_out;
  double* d_jtj = (double*)(d_out + b_rows);
  if (e == cudaSuccess) {
    if (n_pair && max_match) {
      ctx->launches++;
      if (ctx->profiling) ctx_prof_begin(ctx, "k_ba_rows");
      dim3 grid((unsigned)std::max(1, std::min((max_match + 127) / 128, 1024)), (unsigned)n_pair);
      k_ba_rows<<<grid, 128, 0, ctx->stream>>>(d_pairs, d_pts, d_rows);
      if (ctx->profiling) ctx_prof_end(ctx);
    }
    ctx->launches++;
    if (ctx->profiling) ctx_prof_begin(ctx, "k_ba_jtj");
    dim3 gj((unsigned)n_cam, (unsigned)n_cam);
    k_ba_jtj<<<gj, 64, 0, ctx->stream>>>(d_pairs, n_pair, n_cam, d_rows, d_jtj);
    if (ctx->profiling) ctx_prof_end(ctx);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && j_rows && b_rows) e = cudaMemcpyAsync(so, d_rows, b_rows, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(so + (j_rows ? b_rows : 0), d_jtj, b_jtj, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, d_in); ctx_free(ctx, d_out);
  if (e != cudaSuccess) return ctx_cuda(ctx, e, "ba jacobian");
  if (j_rows && b_rows) memcpy(j_rows, so, b_rows);
  memcpy(jtj, so + (j_rows ? b_rows : 0), b_jtj);
  return PANO_OK;
}
