// blend.cu — final composite: LinearBlender and MultiBandBlender.
//
// Replaces BlenderBase::add_image/run (stitch/blender.hh:14-59) for
// LinearBlender::run (stitch/blender.cc:24-96) and MultiBandBlender::run
// (stitch/multiband.cc:19-151).  The per-image inverse map, an opaque
// std::function in the reference, is the closed form of
// stitch/stitcher_image.cc:142-151 + stitch/projection.hh:14-71: its only
// transcendental terms depend on the canvas column (sin/cos of c.x) or row
// (tan of c.y), so they are tabulated on the host with the reference's libm and
// the kernels do IEEE f64 arithmetic only -> bit-identical coordinates.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <vector>
#include <algorithm>

struct BlendImg {
  const float* rgb;       // device
  int w, h;
  int x0, y0, x1, y1;
  double hi[9];
  long long roi_off;      // multiband: first float4 of this image's ROI buffers
  int rw, rh;
};

struct BlendGeom {
  int projection;
  double res_x, res_y, min_x, min_y;
  const double* col_sin;  // [tw+1] cylindrical/spherical
  const double* col_cos;
  const double* row_tan;  // [th+1] spherical
};

// stitcher_image.cc:142-151
__device__ __forceinline__ void coor_func(const BlendImg& im, const BlendGeom& g, int tx, int ty, double* ox, double* oy) {
  double cx = (double)tx * g.res_x + g.min_x;
  double cy = (double)ty * g.res_y + g.min_y;
  double hx, hy, hz;
  if (g.projection == PANO_PROJ_FLAT) { hx = cx; hy = cy; hz = 1.0; }
  else if (g.projection == PANO_PROJ_CYLINDRICAL) { hx = g.col_sin[tx]; hy = cy; hz = g.col_cos[tx]; }
  else { hx = g.col_sin[tx]; hy = g.row_tan[ty]; hz = g.col_cos[tx]; }
  double rx = im.hi[0] * hx + im.hi[1] * hy + im.hi[2] * hz;
  double ry = im.hi[3] * hx + im.hi[4] * hy + im.hi[5] * hz;
  double rz = im.hi[6] * hx + im.hi[7] * hy + im.hi[8] * hz;
  if (rz < 0) { *ox = -10; *oy = -10; return; }
  double denom = 1.0 / rz;
  *ox = rx * denom + im.w * 0.5;
  *oy = ry * denom + im.h * 0.5;
}

// ------------------------------------------------------------ per-tile image list
// Canvas kernels run one thread per output pixel over a 32x8 tile.  A pixel is covered
// by a few images, a mosaic has dozens: the first warp tests every image's range
// against the tile once and compacts the hits IN ORDER (the per-pixel loops must keep
// the reference's image order), so that the per-pixel loop runs over ~3 entries
// instead of n.  More than TILE_LIST_CAP hits: the pixel loop falls back to all n.
#define TILE_LIST_CAP 64
struct TileList { int n; unsigned short idx[TILE_LIST_CAP]; };

__device__ __forceinline__ void build_tile_list(const BlendImg* __restrict__ imgs, int n, int j0, int i0, int j1, int i1,
                                                TileList* tl) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 32) {
    int cnt = 0;
    for (int base = 0; base < n; base += 32) {
      const int k = base + tid;
      bool hit = false;
      if (k < n) {
        const BlendImg& im = imgs[k];
        hit = im.y0 <= i1 && im.y1 >= i0 && im.x0 <= j1 && im.x1 >= j0;   // inclusive: superset of both range rules
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int slot = cnt + __popc(bal & ((1u << tid) - 1));
        if (slot < TILE_LIST_CAP) tl->idx[slot] = (unsigned short)k;
      }
      cnt += __popc(bal);
    }
    if (tid == 0) tl->n = (cnt <= TILE_LIST_CAP && n <= 65535) ? cnt : -1;
  }
  __syncthreads();
}

// ============================================================ linear blend
// blender.cc:24-96.  lazy != 0 selects the LAZY_READ branch (exclusive max
// bounds, accumulate then divide); otherwise the per-pixel branch.
__global__ void k_linear_blend(const BlendImg* __restrict__ imgs, int n, BlendGeom g, int lazy, int ordered,
                               float* __restrict__ out, int tw, int row0, int row1) {
  // rows [row0, row1) of the canvas; `out` starts at row0 (a strip of a row-sharded mosaic, or the whole)
  __shared__ TileList tl;
  {
    const int tj0 = blockIdx.x * blockDim.x, ti0 = row0 + blockIdx.y * blockDim.y;
    build_tile_list(imgs, n, tj0, ti0, tj0 + blockDim.x - 1, ti0 + blockDim.y - 1, &tl);
  }
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = row0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= tw || i >= row1) return;
  const int nl = tl.n < 0 ? n : tl.n;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, wsum = 0.f;
  for (int q = 0; q < nl; ++q) {
    const BlendImg& im = imgs[tl.n < 0 ? q : (int)tl.idx[q]];
    bool in = lazy ? (i >= im.y0 && i < im.y1 && j >= im.x0 && j < im.x1)
                   : (i >= im.y0 && i <= im.y1 && j >= im.x0 && j <= im.x1);
    if (!in) continue;
    double x, y;
    coor_func(im, g, j, i, &x, &y);
    if (x < 0 || x >= im.w || y < 0 || y >= im.h) continue;       // map_coor -> NaN
    float r = (float)y, c = (float)x;
    float c0, c1, c2;
    if (!interpolate_rgb(im.rgb, im.w, im.h, r, c, &c0, &c1, &c2)) continue;
    if (c0 < 0) continue;
    float w = (float)(0.5 - fabs((double)(c / (float)im.w) - 0.5));
    if (!ordered) w = (float)((double)w * (0.5 - fabs((double)(r / (float)im.h) - 0.5)));
    s0 += c0 * w; s1 += c1 * w; s2 += c2 * w;
    wsum += w;
  }
  float* p = out + ((size_t)(i - row0) * tw + j) * 3;
  if (lazy) {
    if (wsum != 0.f) { p[0] = s0 / wsum; p[1] = s1 / wsum; p[2] = s2 / wsum; }
    else { p[0] = -1.f; p[1] = -1.f; p[2] = -1.f; }
  } else {
    if (wsum > 0) {
      float inv = (float)(1.0 / (double)wsum);
      p[0] = s0 * inv; p[1] = s1 * inv; p[2] = s2 * inv;
    } else { p[0] = -1.f; p[1] = -1.f; p[2] = -1.f; }
  }
}

// ============================================================ multiband
// multiband.cc:19-57 create_first_level
__global__ void k_mb_first_level(const BlendImg* __restrict__ imgs, BlendGeom g, float4* __restrict__ cur,
                                 unsigned char* __restrict__ mask) {
  const BlendImg& im = imgs[blockIdx.z];
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= im.rw || i >= im.rh) return;
  double x, y;
  coor_func(im, g, j + im.x0, i + im.y0, &x, &y);
  float c0, c1, c2;
  bool ok = interpolate_rgb(im.rgb, im.w, im.h, (float)y, (float)x, &c0, &c1, &c2);
  if (ok && fminf(c0, fminf(c1, c2)) < 0) ok = false;
  size_t o = (size_t)im.roi_off + (size_t)i * im.rw + j;
  if (!ok) {
    cur[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    mask[o] = 1;
  } else {
    double ox = x / im.w - 0.5, oy = y / im.h - 0.5;
    double ww = (0.5 - fabs(ox)) * (0.5 - fabs(oy));
    if (ww < 0.0) ww = 0.0;
    cur[o] = make_float4(c0, c1, c2, (float)(ww + 1e-6));
    mask[o] = 0;
  }
}

// multiband.cc:125-143 update_weight_map (first image with the largest weight wins)
__global__ void k_mb_weight_argmax(const BlendImg* __restrict__ imgs, int n, float4* __restrict__ cur, int tw,
                                   int row0, int row1) {
  __shared__ TileList tl;
  {
    const int tj0 = blockIdx.x * blockDim.x, ti0 = row0 + blockIdx.y * blockDim.y;
    build_tile_list(imgs, n, tj0, ti0, tj0 + blockDim.x - 1, ti0 + blockDim.y - 1, &tl);
  }
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = row0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= tw || i >= row1) return;
  const int nl = tl.n < 0 ? n : tl.n;
  float mx = 0.f;
  long long best = -1;
  for (int q = 0; q < nl; ++q) {
    const BlendImg& im = imgs[tl.n < 0 ? q : (int)tl.idx[q]];
    if (i >= im.y0 && i <= im.y1 && j >= im.x0 && j <= im.x1) {
      size_t o = (size_t)im.roi_off + (size_t)(i - im.y0) * im.rw + (j - im.x0);
      float w = cur[o].w;
      if (w > mx) { mx = w; best = (long long)o; }
      cur[o].w = 0.f;
    }
  }
  if (best >= 0) cur[best].w = 1.f;
}

// gaussian.hh:29-90 on WeightedPixel (4 floats): column pass ...
struct BlurTaps { int center; float taps[64]; };

#define MB_TW 64
#define MB_TH 32

// Both passes of GaussianBlur::blur on one ROI tile (gaussian.hh:29-90 on WeightedPixel):
// the tile plus a halo of `center` pixels (replicated at the ROI border, exactly the
// clamp of the reference's column/row buffers) is staged in shared memory, the column
// pass writes an intermediate strip to shared memory, the row pass reads it — one read
// and one write of the level per ROI pixel instead of two of each through HBM.
__global__ void __launch_bounds__(256)
k_mb_blur(const BlendImg* __restrict__ imgs, const float4* __restrict__ src, float4* __restrict__ dst,
          const __grid_constant__ BlurTaps bt) {
  extern __shared__ float4 mb_smem[];
  const BlendImg& im = imgs[blockIdx.z];
  const int tx0 = blockIdx.x * MB_TW, ty0 = blockIdx.y * MB_TH;
  if (tx0 >= im.rw || ty0 >= im.rh) return;
  const int c = bt.center, kw = 2 * c + 1;
  const int SW = MB_TW + 2 * c, SH = MB_TH + 2 * c;
  float4* tile = mb_smem;                 // [SH][SW]
  float4* colres = mb_smem + SH * SW;     // [MB_TH][SW]
  const float4* base = src + im.roi_off;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < SH * SW; idx += 256) {
    const int r = idx / SW, q = idx - r * SW;
    const int y = min(max(ty0 - c + r, 0), im.rh - 1), x = min(max(tx0 - c + q, 0), im.rw - 1);
    tile[idx] = __ldg(base + (size_t)y * im.rw + x);
  }
  __syncthreads();
  // column pass: output row i (tile-local), every staged column
  for (int idx = tid; idx < MB_TH * SW; idx += 256) {
    const int i = idx / SW, q = idx - i * SW;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* col = tile + i * SW + q;
    for (int k = 0; k < kw; ++k) {
      const float4 v = col[k * SW];
      const float t = bt.taps[k];
      acc.x += v.x * t; acc.y += v.y * t; acc.z += v.z * t; acc.w += v.w * t;
    }
    colres[idx] = acc;
  }
  __syncthreads();
  // row pass
  for (int idx = tid; idx < MB_TH * MB_TW; idx += 256) {
    const int i = idx / MB_TW, j = idx - i * MB_TW;
    if (ty0 + i >= im.rh || tx0 + j >= im.rw) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* row = colres + i * SW + j;
    for (int k = 0; k < kw; ++k) {
      const float4 v = row[k];
      const float t = bt.taps[k];
      acc.x += v.x * t; acc.y += v.y * t; acc.z += v.z * t; acc.w += v.w * t;
    }
    dst[(size_t)im.roi_off + (size_t)(ty0 + i) * im.rw + tx0 + j] = acc;
  }
}

// multiband.cc:75-108 per-level accumulate (+ :113-121 clamp on the last level)
__global__ void k_mb_accumulate(const BlendImg* __restrict__ imgs, int n, const float4* __restrict__ cur,
                                const float4* __restrict__ next, const unsigned char* __restrict__ mask,
                                int is_last, float* __restrict__ out, unsigned char* __restrict__ tmask, int tw,
                                int row0, int row1) {
  // rows [row0, row1) of the canvas; out / tmask start at row0
  __shared__ TileList tl;
  {
    const int tj0 = blockIdx.x * blockDim.x, ti0 = row0 + blockIdx.y * blockDim.y;
    build_tile_list(imgs, n, tj0, ti0, tj0 + blockDim.x - 1, ti0 + blockDim.y - 1, &tl);
  }
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = row0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= tw || i >= row1) return;
  const int nl = tl.n < 0 ? n : tl.n;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, wsum = 0.f;
  for (int q = 0; q < nl; ++q) {
    const BlendImg& im = imgs[tl.n < 0 ? q : (int)tl.idx[q]];
    if (!(i >= im.y0 && i <= im.y1 && j >= im.x0 && j <= im.x1)) continue;
    size_t o = (size_t)im.roi_off + (size_t)(i - im.y0) * im.rw + (j - im.x0);
    if (mask[o]) continue;
    float4 cc = cur[o];
    float w = cc.w;
    if (w <= 0) continue;
    if (!is_last) {
      float4 cn = next[o];
      s0 += (cc.x - cn.x) * w; s1 += (cc.y - cn.y) * w; s2 += (cc.z - cn.z) * w;
    } else {
      s0 += cc.x * w; s1 += cc.y * w; s2 += cc.z * w;
    }
    wsum += w;
  }
  size_t t = (size_t)(i - row0) * tw + j;
  float* p = out + t * 3;
  bool touched = tmask[t] != 0;
  if (!((double)wsum < 1e-6)) {
    s0 /= wsum; s1 /= wsum; s2 /= wsum;
    if (!touched) { p[0] = s0; p[1] = s1; p[2] = s2; tmask[t] = 1; touched = true; }
    else { p[0] += s0; p[1] += s1; p[2] += s2; }
  }
  if (is_last && touched) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { float v = p[c] < 1.0f ? p[c] : 1.0f; p[c] = v > 0.f ? v : 0.f; }
  }
}

__global__ void k_fill(float* __restrict__ p, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ host driver

struct BlendJob {
  std::vector<BlendImg> imgs;
  BlendGeom g;
  int tw = 0, th = 0;
  long long roi_total = 0;
  int max_rw = 0, max_rh = 0;
};

#define BL_LAUNCH(ctx, name, kernel, grid, block, ...)                              \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    if ((ctx)->profiling) ctx_prof_begin((ctx), (name));                            \
    kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);                     \
    if ((ctx)->profiling) ctx_prof_end((ctx));                                      \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) { rc = ctx_cuda((ctx), _e, name); goto done; }           \
  } while (0)

static int blend_device(pano_ctx* ctx, int n, const pano_blend_image* imgs, const pano_blend_geom* g, int bands,
                        const pano_params* p, float* d_out, int ow, int oh, int row0, int row1) {
  if (!ctx || n <= 0 || !imgs || !g || !p || !d_out || bands < 0) return PANO_ERR_INVALID;
  if (row0 < 0 || row1 > oh || row0 > row1)
    return ctx_fail(ctx, PANO_ERR_INVALID, "blend: rows [%d, %d) outside the %d-row canvas", row0, row1, oh);
  if (row0 == row1) return PANO_OK;
  // Multiband on a row strip: a band at level l of pixel p depends on level 0 inside a
  // radius of the summed half-widths of the blurs up to l, so the strip is computed from
  // each image's ROI clipped to [row0 - H, row1 + H) with H = that sum over all blurred
  // levels.  The replicate rule at a clipped edge differs from the true neighbourhood only
  // within H rows of it, i.e. outside the strip; true ROI edges are kept as they are.
  std::vector<BlurTaps> level_taps;
  int halo = 0;
  for (int level = 0; level + 1 < bands; ++level) {   // multiband.cc:145-151
    float sigma = (float)(sqrt(level * 2 + 1.0) * 4);
    BlurTaps bt;
    memset(&bt, 0, sizeof(bt));
    int kw = host_gauss_kernel(sigma, p->gauss_window_factor, bt.taps, 63);
    if (kw < 0) return ctx_fail(ctx, PANO_ERR_INVALID, "blend: gaussian window %d too wide", -kw);
    bt.center = kw / 2;
    halo += bt.center;
    level_taps.push_back(bt);
  }
  const bool strip = bands > 0 && (row0 != 0 || row1 != oh);
  const int clip0 = (strip && row0 > 0) ? std::max(0, row0 - halo) : INT_MIN;          // first ROI row kept
  const int clip1 = (strip && row1 < oh) ? row1 + halo - 1 : INT_MAX;                  // last ROI row kept
  BlendJob job;
  job.imgs.reserve(n);
  for (int k = 0; k < n; ++k) {
    const pano_blend_image& s = imgs[k];
    if (!s.rgb_hwc || s.w < 2 || s.h < 2 || s.x1 < s.x0 || s.y1 < s.y0 || s.x0 < 0 || s.y0 < 0)
      return ctx_fail(ctx, PANO_ERR_INVALID, "blend: image %d has an invalid shape or range", k);
    job.tw = std::max(job.tw, s.x1); job.th = std::max(job.th, s.y1);
    BlendImg d;
    d.rgb = s.rgb_hwc; d.w = s.w; d.h = s.h;
    d.x0 = s.x0; d.x1 = s.x1;
    d.y0 = std::max(s.y0, clip0); d.y1 = std::min(s.y1, clip1);
    if (d.y0 > d.y1) continue;                         // no row of this image reaches the strip
    memcpy(d.hi, s.homo_inv, sizeof(d.hi));
    d.rw = d.x1 - d.x0 + 1; d.rh = d.y1 - d.y0 + 1;
    d.roi_off = job.roi_total;
    job.roi_total += (long long)align_up((size_t)d.rw * d.rh, 32);
    job.max_rw = std::max(job.max_rw, d.rw); job.max_rh = std::max(job.max_rh, d.rh);
    job.imgs.push_back(d);
  }
  if (job.tw != ow || job.th != oh || ow <= 0 || oh <= 0)
    return ctx_fail(ctx, PANO_ERR_INVALID, "blend: output is %dx%d but target_size is %dx%d", ow, oh, job.tw, job.th);
  const int tw = job.tw, th = job.th;
  n = (int)job.imgs.size();                            // images that reach the strip (all of them for a full canvas)
  if (n == 0) {
    size_t nfl = (size_t)tw * (row1 - row0) * 3;
    PANO_LAUNCH(ctx, "k_fill", k_fill, (unsigned)((nfl + 255) / 256), 256, 0, d_out, nfl, -1.f);
    return PANO_OK;
  }
  // ROIs reach one pixel past the canvas (inclusive max): tables cover [0, tw] / [0, th]
  std::vector<double> tab;
  size_t ncol = (size_t)tw + 2, nrow = (size_t)th + 2;
  if (g->projection != PANO_PROJ_FLAT) {
    tab.resize(2 * ncol + nrow);
    for (size_t j = 0; j < ncol; ++j) {
      double cx = (double)j * g->res_x + g->proj_min_x;
      tab[j] = sin(cx); tab[ncol + j] = cos(cx);
    }
    for (size_t i = 0; i < nrow; ++i) {
      double cy = (double)i * g->res_y + g->proj_min_y;
      tab[2 * ncol + i] = tan(cy);
    }
  }
  int rc = 0;
  BlendImg* d_imgs = nullptr;
  double* d_tab = nullptr;
  float4 *d_cur = nullptr, *d_next = nullptr, *d_tmp = nullptr;
  unsigned char *d_mask = nullptr, *d_tmask = nullptr;
  cudaError_t e = cudaSuccess;
  if ((rc = ctx_alloc(ctx, (void**)&d_imgs, n * sizeof(BlendImg)))) goto done;
  if ((rc = ctx_alloc(ctx, (void**)&d_tab, std::max<size_t>(tab.size(), 1) * sizeof(double)))) goto done;
  {
    void* dsts[2] = {d_imgs, d_tab};
    const void* srcs[2] = {job.imgs.data(), tab.data()};
    size_t sizes[2] = {n * sizeof(BlendImg), tab.size() * sizeof(double)};
    if ((rc = ctx_put_many(ctx, 2, dsts, srcs, sizes))) goto done;
  }
  job.g.projection = g->projection; job.g.res_x = g->res_x; job.g.res_y = g->res_y;
  job.g.min_x = g->proj_min_x; job.g.min_y = g->proj_min_y;
  job.g.col_sin = d_tab; job.g.col_cos = d_tab + ncol; job.g.row_tan = d_tab + 2 * ncol;
  {
    dim3 b(32, 8), gt(ceil_div(tw, 32), ceil_div(th, 8));
    if (bands == 0) {
      dim3 gs(ceil_div(tw, 32), ceil_div(row1 - row0, 8));
      BL_LAUNCH(ctx, "k_linear_blend", k_linear_blend, gs, b, d_imgs, n, job.g, p->lazy_read, p->ordered_input, d_out, tw,
                row0, row1);
    } else {
      size_t roi = (size_t)job.roi_total;
      if ((rc = ctx_alloc(ctx, (void**)&d_cur, roi * sizeof(float4)))) goto done;
      if ((rc = ctx_alloc(ctx, (void**)&d_next, roi * sizeof(float4)))) goto done;
      if ((rc = ctx_alloc(ctx, (void**)&d_mask, roi))) goto done;
      const size_t strip_px = (size_t)tw * (row1 - row0);
      // the weight map is needed wherever a clipped ROI has pixels on the canvas
      const int wrow0 = std::max(0, std::max(row0 - halo, clip0)), wrow1 = std::min(th, strip ? row1 + halo : th);
      if ((rc = ctx_alloc(ctx, (void**)&d_tmask, strip_px))) goto done;
      if ((rc = ctx_zero(ctx, d_tmask, strip_px))) goto done;
      dim3 gr(ceil_div(job.max_rw, 32), ceil_div(job.max_rh, 8), n);
      dim3 gw(ceil_div(tw, 32), ceil_div(wrow1 - wrow0, 8)), gs(ceil_div(tw, 32), ceil_div(row1 - row0, 8));
      BL_LAUNCH(ctx, "k_mb_first_level", k_mb_first_level, gr, b, d_imgs, job.g, d_cur, d_mask);
      BL_LAUNCH(ctx, "k_mb_weight_argmax", k_mb_weight_argmax, gw, b, d_imgs, n, d_cur, tw, wrow0, wrow1);
      {
        size_t nfl = strip_px * 3;
        BL_LAUNCH(ctx, "k_fill", k_fill, (unsigned)((nfl + 255) / 256), 256, d_out, nfl, -1.f);
      }
      for (int level = 0; level < bands; ++level) {
        int is_last = level == bands - 1;
        if (!is_last) {
          const BlurTaps& bt = level_taps[level];
          const int kw = 2 * bt.center + 1;
          {
            const int c = bt.center;
            const size_t smem = sizeof(float4) * ((size_t)(MB_TH + 2 * c) * (MB_TW + 2 * c) + (size_t)MB_TH * (MB_TW + 2 * c));
            if (smem > 200 * 1024) { rc = ctx_fail(ctx, PANO_ERR_INVALID, "blend: gaussian window %d too wide", kw); goto done; }
            e = cudaFuncSetAttribute(k_mb_blur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { rc = ctx_cuda(ctx, e, "k_mb_blur attribute"); goto done; }
            dim3 gb(ceil_div(job.max_rw, MB_TW), ceil_div(job.max_rh, MB_TH), n);
            ctx->launches++;
            if (ctx->profiling) ctx_prof_begin(ctx, "k_mb_blur");
            k_mb_blur<<<gb, 256, smem, ctx->stream>>>(d_imgs, d_cur, d_next, bt);
            if (ctx->profiling) ctx_prof_end(ctx);
            e = cudaGetLastError();
            if (e != cudaSuccess) { rc = ctx_cuda(ctx, e, "k_mb_blur"); goto done; }
          }
        }
        BL_LAUNCH(ctx, "k_mb_accumulate", k_mb_accumulate, gs, b, d_imgs, n, d_cur, d_next, d_mask, is_last, d_out,
                  d_tmask, tw, row0, row1);
        if (!is_last) std::swap(d_cur, d_next);
      }
    }
  }
done:
  ctx_free(ctx, d_imgs); ctx_free(ctx, d_tab); ctx_free(ctx, d_cur); ctx_free(ctx, d_next); ctx_free(ctx, d_tmp);
  ctx_free(ctx, d_mask); ctx_free(ctx, d_tmask);
  return rc;
}

extern "C" {

int pano_blend_target_size(int n, const pano_blend_image* imgs, int* ow, int* oh) {
  if (n <= 0 || !imgs || !ow || !oh) return PANO_ERR_INVALID;
  int tw = 0, th = 0;
  for (int k = 0; k < n; ++k) { tw = std::max(tw, imgs[k].x1); th = std::max(th, imgs[k].y1); }
  *ow = tw; *oh = th;
  return PANO_OK;
}

int pano_blend_dev(pano_ctx* ctx, int n, const pano_blend_image* imgs, const pano_blend_geom* g, int bands,
                   const pano_params* p, float* d_out, int ow, int oh) {
  ctx_enter(ctx);
  return blend_device(ctx, n, imgs, g, bands, p, d_out, ow, oh, 0, oh);
}

int pano_blend_rows_dev(pano_ctx* ctx, int n, const pano_blend_image* imgs, const pano_blend_geom* g, int bands,
                        const pano_params* p, float* d_out_rows, int ow, int oh, int row0, int row1) {
  ctx_enter(ctx);
  return blend_device(ctx, n, imgs, g, bands, p, d_out_rows, ow, oh, row0, row1);
}

int pano_blend(pano_ctx* ctx, int n, const pano_blend_image* imgs, const pano_blend_geom* g, int bands,
               const pano_params* p, float* out, int ow, int oh) {
  ctx_enter(ctx);
  if (!ctx || n <= 0 || !imgs || !out) return PANO_ERR_INVALID;
  std::vector<pano_blend_image> dimgs(imgs, imgs + n);
  std::vector<float*> bufs(n, nullptr);
  float* d_out = nullptr;
  int rc = 0;
  cudaError_t e = cudaSuccess;
  for (int k = 0; k < n && !rc; ++k) {
    if (!imgs[k].rgb_hwc || imgs[k].w <= 0 || imgs[k].h <= 0) { rc = ctx_fail(ctx, PANO_ERR_INVALID, "blend: image %d empty", k); break; }
    size_t bytes = (size_t)imgs[k].w * imgs[k].h * 3 * sizeof(float);
    rc = ctx_alloc(ctx, (void**)&bufs[k], bytes);
    if (rc) break;
    e = cudaMemcpyAsync(bufs[k], imgs[k].rgb_hwc, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { rc = ctx_cuda(ctx, e, "blend image upload"); break; }
    dimgs[k].rgb_hwc = bufs[k];
  }
  size_t ob = (size_t)std::max(ow, 0) * std::max(oh, 0) * 3 * sizeof(float);
  if (!rc) rc = ctx_alloc(ctx, (void**)&d_out, ob);
  if (!rc) rc = blend_device(ctx, n, dimgs.data(), g, bands, p, d_out, ow, oh, 0, oh);
  if (!rc) {
    e = cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = ctx_cuda(ctx, e, "blend download");
  }
  for (auto b : bufs) ctx_free(ctx, b);
  ctx_free(ctx, d_out);
  return rc;
}

}  // extern "C"
