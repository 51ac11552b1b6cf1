// warp.cu — cylindrical pre-warp of an image and its keypoints.
//
// Replaces CylinderWarper::warp (stitch/warp.hh:41-66) = CylinderProject::project
// (stitch/warp.cc:25-67), proj/proj_r (:13-23), get_projector (:70-75).
// All geometry is f64.  The transcendental part of the inverse map depends only
// on the destination COLUMN (x = r*tan(px)+cx, 1/cos(px)), so it is evaluated on
// the host with the same libm the reference calls and uploaded as two per-column
// tables; the kernel then performs only IEEE mul/div/add and the f32 bilinear
// gather, which makes the image bit-identical to the reference.
#include "common.cuh"
#include <math.h>
#include <float.h>
#include <vector>

struct CylProj { double cx, cy; int r; int sizefactor; };

static CylProj get_projector(int w, int h, double h_factor, const pano_params* p) {
  CylProj c;
  c.r = (int)(hypot((double)w, (double)h) * (p->focal_length / 43.266));
  c.cx = w / 2;
  c.cy = h / 2 * h_factor;
  c.sizefactor = c.r;
  return c;
}

static void proj(const CylProj& c, double px, double py, double* ox, double* oy) {
  *ox = atan((px - c.cx) / c.r);
  *oy = (py - c.cy) / hypot(px - c.cx, (double)c.r);
}

// Bounds of proj over the whole pixel grid (warp.cc:47-52 scans every pixel; min
// starts at +max, max starts at 0).  proj.x is monotone in the column and proj.y,
// for a fixed row, is extremal at the column nearest the centre, so the scan
// reduces to the border rows/columns — evaluated with the same expressions.
static void proj_bounds(const CylProj& c, int w, int h, double* minx, double* miny, double* maxx, double* maxy) {
  double mnx = DBL_MAX, mny = DBL_MAX, mxx = 0, mxy = 0;
  const int rows[2] = {0, h - 1};
  for (int ri = 0; ri < 2; ++ri)
    for (int j = 0; j < w; ++j) {
      double x, y;
      proj(c, j, rows[ri], &x, &y);
      if (x < mnx) mnx = x;
      if (y < mny) mny = y;
      if (mxx < x) mxx = x;
      if (mxy < y) mxy = y;
    }
  const int cols[2] = {0, w - 1};
  for (int ci = 0; ci < 2; ++ci)
    for (int i = 0; i < h; ++i) {
      double x, y;
      proj(c, cols[ci], i, &x, &y);
      if (x < mnx) mnx = x;
      if (y < mny) mny = y;
      if (mxx < x) mxx = x;
      if (mxy < y) mxy = y;
    }
  *minx = mnx; *miny = mny; *maxx = mxx; *maxy = mxy;
}

// warp.cc:46-67 project(Shape2D&, pts)
static void project_shape(const CylProj& c, int* w, int* h, double* kpts, int nk, double* offx, double* offy) {
  double minx, miny, maxx, maxy;
  proj_bounds(c, *w, *h, &minx, &miny, &maxx, &maxy);
  maxx = maxx * c.sizefactor; maxy = maxy * c.sizefactor;
  minx = minx * c.sizefactor; miny = miny * c.sizefactor;
  double rsx = maxx - minx, rsy = maxy - miny;
  *offx = minx * (-1); *offy = miny * (-1);
  int sx = (int)rsx, sy = (int)rsy;
  for (int i = 0; i < nk; ++i) {
    double x, y;
    proj(c, kpts[2 * i] + *w / 2, kpts[2 * i + 1] + *h / 2, &x, &y);
    x = x * c.sizefactor + *offx;
    y = y * c.sizefactor + *offy;
    x -= sx / 2;
    y -= sy / 2;
    kpts[2 * i] = x; kpts[2 * i + 1] = y;
  }
  *w = sx; *h = sy;
}

// one thread per destination pixel (warp.cc:33-41)
__global__ void k_cyl_warp(const float* __restrict__ src, int w, int h, float* __restrict__ dst, int ow, int oh,
                           const double* __restrict__ col_x, const double* __restrict__ col_cos, double r,
                           double cy, double offy, double sizefactor_inv) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= ow || i >= oh) return;
  double py = ((double)i - offy) * sizefactor_inv;
  double x = col_x[j];
  double y = py * r / col_cos[j] + cy;
  float o0 = -1.f, o1 = -1.f, o2 = -1.f;
  if (x >= 0 && x <= (double)(w - 1) && y >= 0 && y <= (double)(h - 1)) {
    float c0, c1, c2;
    if (interpolate_rgb(src, w, h, (float)y, (float)x, &c0, &c1, &c2)) { o0 = c0; o1 = c1; o2 = c2; }
  }
  float* p = dst + ((size_t)i * ow + j) * 3;
  p[0] = o0; p[1] = o1; p[2] = o2;
}

// The same for a batch of device-resident images: blockIdx.z = image; per-image parameters and the
// two per-column tables (concatenated) sit in device memory.
struct CylJobDev {
  const float* src; float* dst;
  int w, h, ow, oh;
  long long tab_off;           // first entry of this image's col_x[ow] followed by col_cos[ow]
  double r, cy, offy, sizefactor_inv;
};

__global__ void k_cyl_warp_batch(const CylJobDev* __restrict__ jobs, const double* __restrict__ tabs) {
  const CylJobDev jb = jobs[blockIdx.z];
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= jb.ow || i >= jb.oh) return;
  const double* col_x = tabs + jb.tab_off;
  const double* col_cos = col_x + jb.ow;
  double py = ((double)i - jb.offy) * jb.sizefactor_inv;
  double x = col_x[j];
  double y = py * jb.r / col_cos[j] + jb.cy;
  float o0 = -1.f, o1 = -1.f, o2 = -1.f;
  if (x >= 0 && x <= (double)(jb.w - 1) && y >= 0 && y <= (double)(jb.h - 1)) {
    float c0, c1, c2;
    if (interpolate_rgb(jb.src, jb.w, jb.h, (float)y, (float)x, &c0, &c1, &c2)) { o0 = c0; o1 = c1; o2 = c2; }
  }
  float* p = jb.dst + ((size_t)i * jb.ow + j) * 3;
  p[0] = o0; p[1] = o1; p[2] = o2;
}

extern "C" {

int pano_cyl_warp_batch_dev(pano_ctx* ctx, int n, const pano_cyl_job* jobs, double h_factor, const pano_params* p) {
  ctx_enter(ctx);
  if (!ctx || n < 0 || (n && !jobs) || !p) return PANO_ERR_INVALID;
  if (n == 0) return PANO_OK;
  std::vector<CylJobDev> dj(n);
  std::vector<double> tabs;
  int max_ow = 0, max_oh = 0;
  for (int k = 0; k < n; ++k) {
    const pano_cyl_job& jb = jobs[k];
    if (!jb.d_rgb_hwc || !jb.d_out_hwc || jb.w <= 1 || jb.h <= 1 || jb.n_kpts < 0 || (jb.n_kpts && !jb.kpts_xy))
      return ctx_fail(ctx, PANO_ERR_INVALID, "cyl_warp_batch: job %d has a null pointer or an empty image", k);
    CylProj c = get_projector(jb.w, jb.h, h_factor, p);
    if (c.r <= 0) return ctx_fail(ctx, PANO_ERR_INVALID, "cylinder radius <= 0");
    int sw = jb.w, sh = jb.h;
    double offx, offy;
    project_shape(c, &sw, &sh, jb.kpts_xy, jb.n_kpts, &offx, &offy);       // keypoints: host arithmetic, in place
    if (sw != jb.out_w || sh != jb.out_h || sw <= 0 || sh <= 0)
      return ctx_fail(ctx, PANO_ERR_INVALID, "cyl_warp_batch: job %d output buffer is %dx%d but the warp is %dx%d", k,
                      jb.out_w, jb.out_h, sw, sh);
    const double sizefactor_inv = 1.0 / c.sizefactor;
    const size_t t0 = tabs.size();
    tabs.resize(t0 + 2 * (size_t)sw);
    for (int j = 0; j < sw; ++j) {                                          // warp.cc:19-23 proj_r per column
      double px = ((double)j - offx) * sizefactor_inv;
      tabs[t0 + j] = c.r * tan(px) + c.cx;
      tabs[t0 + sw + j] = cos(px);
    }
    dj[k] = CylJobDev{jb.d_rgb_hwc, jb.d_out_hwc, jb.w, jb.h, sw, sh, (long long)t0, (double)c.r, c.cy, offy, sizefactor_inv};
    max_ow = std::max(max_ow, sw); max_oh = std::max(max_oh, sh);
  }
  CylJobDev* d_jobs = nullptr;
  double* d_tabs = nullptr;
  int rc = 0;
  if ((rc = ctx_alloc(ctx, (void**)&d_jobs, dj.size() * sizeof(CylJobDev))) ||
      (rc = ctx_alloc(ctx, (void**)&d_tabs, tabs.size() * sizeof(double)))) {
    ctx_free(ctx, d_jobs); ctx_free(ctx, d_tabs);
    return rc;
  }
  rc = ctx_put(ctx, d_jobs, dj.data(), dj.size() * sizeof(CylJobDev));
  if (!rc) rc = ctx_put(ctx, d_tabs, tabs.data(), tabs.size() * sizeof(double));
  if (!rc) {
    dim3 b(32, 8), g(ceil_div(max_ow, 32), ceil_div(max_oh, 8), n);
    ctx->launches++;
    if (ctx->profiling) ctx_prof_begin(ctx, "k_cyl_warp");
    k_cyl_warp_batch<<<g, b, 0, ctx->stream>>>(d_jobs, d_tabs);
    if (ctx->profiling) ctx_prof_end(ctx);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = ctx_cuda(ctx, e, "k_cyl_warp_batch");
  }
  ctx_free(ctx, d_jobs); ctx_free(ctx, d_tabs);      // stream-ordered: released after the kernel
  return rc;
}

int pano_cyl_warp_shape(int w, int h, double h_factor, const pano_params* p, int* ow, int* oh, double* offx,
                        double* offy) {
  if (w <= 0 || h <= 0 || !p || !ow || !oh || !offx || !offy) return PANO_ERR_INVALID;
  CylProj c = get_projector(w, h, h_factor, p);
  if (c.r <= 0) return PANO_ERR_INVALID;
  *ow = w; *oh = h;
  project_shape(c, ow, oh, nullptr, 0, offx, offy);
  return PANO_OK;
}

int pano_cyl_warp(pano_ctx* ctx, const float* rgb, int w, int h, double h_factor, const pano_params* p, float* out,
                  int ow, int oh, double* kpts, int nk) {
  ctx_enter(ctx);
  if (!ctx || !rgb || !out || !p || w <= 1 || h <= 1 || nk < 0 || (nk && !kpts)) return PANO_ERR_INVALID;
  CylProj c = get_projector(w, h, h_factor, p);
  if (c.r <= 0) return ctx_fail(ctx, PANO_ERR_INVALID, "cylinder radius <= 0");
  int sw = w, sh = h;
  double offx, offy;
  project_shape(c, &sw, &sh, kpts, nk, &offx, &offy);
  if (sw != ow || sh != oh || ow <= 0 || oh <= 0)
    return ctx_fail(ctx, PANO_ERR_INVALID, "cyl_warp: output buffer is %dx%d but the warp is %dx%d", ow, oh, sw, sh);
  const double sizefactor_inv = 1.0 / c.sizefactor;
  // per-column tables (warp.cc:19-23 proj_r with p.x = (j - offset.x) * sizefactor_inv)
  std::vector<double> tab(2 * (size_t)ow);
  for (int j = 0; j < ow; ++j) {
    double px = ((double)j - offx) * sizefactor_inv;
    tab[j] = c.r * tan(px) + c.cx;
    tab[ow + j] = cos(px);
  }
  float *d_src = nullptr, *d_dst = nullptr;
  double* d_tab = nullptr;
  size_t bs = (size_t)w * h * 3 * sizeof(float), bd = (size_t)ow * oh * 3 * sizeof(float);
  int rc = 0;
  if ((rc = ctx_alloc(ctx, (void**)&d_src, bs)) || (rc = ctx_alloc(ctx, (void**)&d_dst, bd)) ||
      (rc = ctx_alloc(ctx, (void**)&d_tab, tab.size() * sizeof(double)))) {
    ctx_free(ctx, d_src); ctx_free(ctx, d_dst); ctx_free(ctx, d_tab);
    return rc;
  }
  cudaError_t e = cudaMemcpyAsync(d_src, rgb, bs, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    dim3 b(32, 8), g(ceil_div(ow, 32), ceil_div(oh, 8));
    ctx->launches++;
    if (ctx->profiling) ctx_prof_begin(ctx, "k_cyl_warp");
    k_cyl_warp<<<g, b, 0, ctx->stream>>>(d_src, w, h, d_dst, ow, oh, d_tab, d_tab + ow, (double)c.r, c.cy, offy,
                                        sizefactor_inv);
    if (ctx->profiling) ctx_prof_end(ctx);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_dst, bd, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, d_src); ctx_free(ctx, d_dst); ctx_free(ctx, d_tab);
  if (e != cudaSuccess) return ctx_cuda(ctx, e, "pano_cyl_warp");
  return PANO_OK;
}

}  // extern "C"
