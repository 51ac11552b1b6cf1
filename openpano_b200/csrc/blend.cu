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
#include "blur_tile.cuh"
#include <math.h>
#include <string.h>
#include <vector>
#include <algorithm>

struct BlendImg {
  const float* rgb;       // device
  int w, h;
  int x0, y0, x1, y1;
  double hi[9];
  // multiband: the ROI is kept as FOUR PLANES (r, g, b, weight) of rh x pitch floats — the
  // reference's WeightedPixel (multiband.hh:13-23) de-interleaved, so that the weight map, the
  // blur and the accumulation each touch only the bytes they use and rows start on 128-byte lines
  long long roi_off;      // first float of plane 0 in the level buffers
  long long plane;        // floats per plane (pitch * rh)
  long long mask_off;     // first byte of the validity mask (same pitch)
  int rw, rh, pitch;
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
__global__ void k_mb_first_level(const BlendImg* __restrict__ imgs, BlendGeom g, float* __restrict__ cur,
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
  const size_t o = (size_t)i * im.pitch + j;
  float* p = cur + im.roi_off + o;
  float ww = 0.f;
  if (!ok) { c0 = 0.f; c1 = 0.f; c2 = 0.f; }
  else {
    double ox = x / im.w - 0.5, oy = y / im.h - 0.5;
    double wd = (0.5 - fabs(ox)) * (0.5 - fabs(oy));
    if (wd < 0.0) wd = 0.0;
    ww = (float)(wd + 1e-6);
  }
  p[0] = c0; p[im.plane] = c1; p[2 * im.plane] = c2; p[3 * im.plane] = ww;
  mask[im.mask_off + o] = ok ? 0 : 1;
}

// multiband.cc:125-143 update_weight_map (first image with the largest weight wins)
__global__ void k_mb_weight_argmax(const BlendImg* __restrict__ imgs, int n, float* __restrict__ cur, int tw,
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
  // Two passes so that the weight loads of all covering images are in flight together: a
  // store into `cur` between two loads from it would order them (the compiler cannot prove the
  // planes disjoint), and the kernel then crawls through one DRAM round trip per image.
  // Every location is read and written by this thread only, so the read-only path is safe.
  float mx = 0.f;
  int best = -1;
  for (int q = 0; q < nl; ++q) {
    const BlendImg& im = imgs[tl.n < 0 ? q : (int)tl.idx[q]];
    if (i >= im.y0 && i <= im.y1 && j >= im.x0 && j <= im.x1) {
      const float w = __ldg(cur + im.roi_off + 3 * im.plane + (size_t)(i - im.y0) * im.pitch + (j - im.x0));
      if (w > mx) { mx = w; best = q; }
    }
  }
  for (int q = 0; q < nl; ++q) {
    const BlendImg& im = imgs[tl.n < 0 ? q : (int)tl.idx[q]];
    if (i >= im.y0 && i <= im.y1 && j >= im.x0 && j <= im.x1)
      cur[im.roi_off + 3 * im.plane + (size_t)(i - im.y0) * im.pitch + (j - im.x0)] = q == best ? 1.f : 0.f;
  }
}

// gaussian.hh:29-90 on WeightedPixel: every channel of the pixel goes through the same
// column-then-row passes, i.e. four independent scalar planes.
struct BlurTaps { int center; float taps[64]; };
struct MbPlane { long long off; int w, h, pitch, pad; };   // one (image, channel) plane of a level buffer

// Both passes of one blur level on 64x32 tiles of every plane (blur_tile.cuh), PERSISTENT CTAs:
// the tile + halo of the next work item is fetched by TMA into the other half of a double
// buffer while this one is computed.  TMA zero-fills outside the ROI where the reference's
// line buffers replicate the ROI edge (gaussian.hh:52-58,74-81): border tiles patch those cells
// from the staged in-range ones.
template <int C>
__global__ void __launch_bounds__(BT_THREADS, 4)
k_mb_blur_tma(const MbPlane* __restrict__ planes, const int2* __restrict__ span, int n_planes, int n_tiles,
              const TmaDesc* __restrict__ maps, float* __restrict__ dst, const __grid_constant__ BlurTaps bt) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ BlurTile s_tile[2];                   // looked up once per tile by thread 0 (binary search)
  constexpr int RX = (C + 3) & ~3;                 // TMA box origin: 16-byte aligned columns
  constexpr int GW = BT_W + 2 * RX, GH = BT_H + 2 * C;
  constexpr int GSZ = (GH * GW + 31) & ~31;
  float* grey0 = smem;
  float* colbuf = smem + 2 * GSZ;
  float* outT = colbuf + BLUR_COLBUF_FLOATS(C);
  const int tid = threadIdx.x;
  constexpr uint32_t tile_bytes = (uint32_t)(GH * GW * sizeof(float));
  int t = blockIdx.x;
  if (tid == 0) {
    sbar_init(sm_u32(&s_bar[0]), 1);
    sbar_init(sm_u32(&s_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (t < n_tiles) {
      const BlurTile tl = find_blur_tile(span, n_planes, t);
      s_tile[0] = tl;
      sbar_expect_tx(sm_u32(&s_bar[0]), tile_bytes);
      tma_load_2d(sm_u32(grey0), maps + tl.om, tl.tx * BT_W - RX, tl.ty * BT_H - C, sm_u32(&s_bar[0]));
    }
  }
  __syncthreads();
  for (int it = 0; t < n_tiles; t += gridDim.x, ++it) {
    const int b = it & 1;
    float* grey = grey0 + b * GSZ;
    const BlurTile tl = s_tile[b];
    const MbPlane pl = planes[tl.om];
    const int x0 = tl.tx * BT_W, y0 = tl.ty * BT_H;
    if (tid == 0 && t + (int)gridDim.x < n_tiles) {
      const BlurTile nx = find_blur_tile(span, n_planes, t + gridDim.x);
      s_tile[b ^ 1] = nx;      // read by everyone after the barrier that ends this iteration
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      sbar_expect_tx(sm_u32(&s_bar[b ^ 1]), tile_bytes);
      tma_load_2d(sm_u32(grey0 + (b ^ 1) * GSZ), maps + nx.om, nx.tx * BT_W - RX, nx.ty * BT_H - C, sm_u32(&s_bar[b ^ 1]));
    }
    sbar_wait(sm_u32(&s_bar[b]), (uint32_t)(it >> 1) & 1u);
    if (x0 - RX < 0 || y0 - C < 0 || x0 + BT_W + RX > pl.w || y0 + BT_H + C > pl.h) {   // uniform per CTA
      for (int i = tid; i < GH * GW; i += BT_THREADS) {
        const int yy = i / GW, xx = i - yy * GW;
        const int gy = y0 + yy - C, gx = x0 + xx - RX;
        const int cy = min(max(gy, 0), pl.h - 1), cx = min(max(gx, 0), pl.w - 1);
        if (cy != gy || cx != gx) grey[i] = grey[(cy - y0 + C) * GW + (cx - x0 + RX)];
      }
      __syncthreads();
    }
    blur_level<C>(grey, colbuf, outT, bt.taps, C, RX, GW, tid);
    const int tx = tid & (BT_W - 1), ty = tid / BT_W;   // 64 x 4
    const int gx = x0 + tx;
    float* out = dst + pl.off;
#pragma unroll
    for (int i = 0; i < BT_H / 4; ++i) {
      const int y = ty + 4 * i, gy = y0 + y;
      if (gx < pl.w && gy < pl.h) out[(size_t)gy * pl.pitch + gx] = outT[y * (BT_W + 1) + tx];
    }
    __syncthreads();   // colbuf / outT / this staging buffer are free for the next tiles
  }
}

// Any other window width (GAUSS_WINDOW_FACTOR != 6, more than 5 bands): one tile per CTA,
// clamped loads, taps looped from the table.
#define MB_TW 64
#define MB_TH 32
__global__ void __launch_bounds__(256)
k_mb_blur_generic(const MbPlane* __restrict__ planes, const int2* __restrict__ span, int n_planes,
                  const float* __restrict__ src, float* __restrict__ dst, const __grid_constant__ BlurTaps bt) {
  extern __shared__ float mb_smem[];
  const BlurTile tl = find_blur_tile(span, n_planes, blockIdx.x);
  const MbPlane pl = planes[tl.om];
  const int tx0 = tl.tx * MB_TW, ty0 = tl.ty * MB_TH;
  const int c = bt.center, kw = 2 * c + 1;
  const int SW = MB_TW + 2 * c, SH = MB_TH + 2 * c;
  float* tile = mb_smem;                 // [SH][SW]
  float* colres = mb_smem + SH * SW;     // [MB_TH][SW]
  const float* base = src + pl.off;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < SH * SW; idx += 256) {
    const int r = idx / SW, q = idx - r * SW;
    const int y = min(max(ty0 - c + r, 0), pl.h - 1), x = min(max(tx0 - c + q, 0), pl.w - 1);
    tile[idx] = __ldg(base + (size_t)y * pl.pitch + x);
  }
  __syncthreads();
  for (int idx = tid; idx < MB_TH * SW; idx += 256) {
    const int i = idx / SW, q = idx - i * SW;
    float acc = 0.f;
    const float* col = tile + i * SW + q;
    for (int k = 0; k < kw; ++k) acc += col[k * SW] * bt.taps[k];
    colres[idx] = acc;
  }
  __syncthreads();
  for (int idx = tid; idx < MB_TH * MB_TW; idx += 256) {
    const int i = idx / MB_TW, j = idx - i * MB_TW;
    if (ty0 + i >= pl.h || tx0 + j >= pl.w) continue;
    float acc = 0.f;
    const float* row = colres + i * SW + j;
    for (int k = 0; k < kw; ++k) acc += row[k] * bt.taps[k];
    dst[pl.off + (size_t)(ty0 + i) * pl.pitch + tx0 + j] = acc;
  }
}

// multiband.cc:75-108 per-level accumulate (+ :113-121 clamp on the last level).  `first`: no
// level has touched the strip yet, so every pixel is written (value, or -1 = Color::NO) and
// no separate fill pass is needed.
__global__ void k_mb_accumulate(const BlendImg* __restrict__ imgs, int n, const float* __restrict__ cur,
                                const float* __restrict__ next, const unsigned char* __restrict__ mask,
                                int first, int is_last, float* __restrict__ out, unsigned char* __restrict__ tmask, int tw,
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
    const size_t o = (size_t)(i - im.y0) * im.pitch + (j - im.x0);
    // the weight plane decides: colours are fetched only where this image contributes
    const float w = __ldg(cur + im.roi_off + 3 * im.plane + o);
    if (w <= 0) continue;
    if (mask[im.mask_off + o]) continue;
    const float* pc = cur + im.roi_off + o;
    const float c0 = __ldg(pc), c1 = __ldg(pc + im.plane), c2 = __ldg(pc + 2 * im.plane);
    if (!is_last) {
      const float* pn = next + im.roi_off + o;
      const float n0 = __ldg(pn), n1 = __ldg(pn + im.plane), n2 = __ldg(pn + 2 * im.plane);
      s0 += (c0 - n0) * w; s1 += (c1 - n1) * w; s2 += (c2 - n2) * w;
    } else {
      s0 += c0 * w; s1 += c1 * w; s2 += c2 * w;
    }
    wsum += w;
  }
  size_t t = (size_t)(i - row0) * tw + j;
  float* p = out + t * 3;
  bool touched = first ? false : tmask[t] != 0;
  float v0 = -1.f, v1 = -1.f, v2 = -1.f;
  if (touched) { v0 = p[0]; v1 = p[1]; v2 = p[2]; }
  bool dirty = first != 0;
  if (!((double)wsum < 1e-6)) {
    s0 /= wsum; s1 /= wsum; s2 /= wsum;
    if (!touched) { v0 = s0; v1 = s1; v2 = s2; touched = true; }
    else { v0 += s0; v1 += s1; v2 += s2; }
    dirty = true;
  }
  if (is_last && touched) {
    v0 = v0 < 1.0f ? v0 : 1.0f; v0 = v0 > 0.f ? v0 : 0.f;
    v1 = v1 < 1.0f ? v1 : 1.0f; v1 = v1 > 0.f ? v1 : 0.f;
    v2 = v2 < 1.0f ? v2 : 1.0f; v2 = v2 > 0.f ? v2 : 0.f;
    dirty = true;
  }
  if (dirty) { p[0] = v0; p[1] = v1; p[2] = v2; }
  if (first || touched) tmask[t] = touched ? 1 : 0;
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
  long long roi_floats = 0;   // floats of one level buffer (4 planes per image)
  long long mask_bytes = 0;
  int max_rw = 0, max_rh = 0;
};

#define BL_LAUNCH(ctx, name, kernel, grid, block, smem, ...)                        \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    if ((ctx)->profiling) ctx_prof_begin((ctx), (name));                            \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                \
    if ((ctx)->profiling) ctx_prof_end((ctx));                                      \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) { rc = ctx_cuda((ctx), _e, name); goto done; }           \
  } while (0)

template <int C>
static cudaError_t launch_mb_blur_tma(pano_ctx* ctx, int grid, const MbPlane* planes, const int2* span, int n_planes,
                                      int n_tiles, const TmaDesc* maps, float* dst, const BlurTaps& bt) {
  constexpr int RX = (C + 3) & ~3, GW = BT_W + 2 * RX, GH = BT_H + 2 * C, GSZ = (GH * GW + 31) & ~31;
  const size_t smem = (size_t)(2 * GSZ + BLUR_COLBUF_FLOATS(C) + BT_H * (BT_W + 1)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(k_mb_blur_tma<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_mb_blur_tma<C><<<grid, BT_THREADS, smem, ctx->stream>>>(planes, span, n_planes, n_tiles, maps, dst, bt);
  return cudaGetLastError();
}

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
    d.pitch = (int)align_up((size_t)d.rw, 32);
    d.plane = (long long)d.pitch * d.rh;
    d.roi_off = job.roi_floats;
    d.mask_off = job.mask_bytes;
    job.roi_floats += 4 * d.plane;
    job.mask_bytes += d.plane;
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
  float *d_cur = nullptr, *d_next = nullptr;
  unsigned char *d_mask = nullptr, *d_tmask = nullptr;
  MbPlane* d_planes = nullptr;
  int2* d_span = nullptr;
  TmaDesc* d_maps = nullptr;
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
    dim3 b(32, 8);
    if (bands == 0) {
      dim3 gs(ceil_div(tw, 32), ceil_div(row1 - row0, 8));
      BL_LAUNCH(ctx, "k_linear_blend", k_linear_blend, gs, b, 0, d_imgs, n, job.g, p->lazy_read, p->ordered_input, d_out, tw,
                row0, row1);
    } else {
      const size_t roi = (size_t)job.roi_floats;
      if ((rc = ctx_alloc(ctx, (void**)&d_cur, roi * sizeof(float)))) goto done;
      if ((rc = ctx_alloc(ctx, (void**)&d_next, roi * sizeof(float)))) goto done;
      if ((rc = ctx_alloc(ctx, (void**)&d_mask, (size_t)job.mask_bytes))) goto done;
      const size_t strip_px = (size_t)tw * (row1 - row0);
      // the weight map is needed wherever a clipped ROI has pixels on the canvas
      const int wrow0 = std::max(0, std::max(row0 - halo, clip0)), wrow1 = std::min(th, strip ? row1 + halo : th);
      if ((rc = ctx_alloc(ctx, (void**)&d_tmask, strip_px))) goto done;
      // plane table of the blur launches: (image, channel) -> offset, size; tiles of 64x32
      const int n_planes = 4 * n;
      std::vector<MbPlane> planes(n_planes);
      std::vector<int2> span(n_planes);
      int n_tiles = 0;
      for (int k = 0; k < n; ++k)
        for (int ch = 0; ch < 4; ++ch) {
          const BlendImg& im = job.imgs[k];
          planes[4 * k + ch] = MbPlane{im.roi_off + ch * im.plane, im.rw, im.rh, im.pitch, 0};
          span[4 * k + ch] = make_int2(n_tiles, ceil_div(im.rw, BT_W));
          n_tiles += ceil_div(im.rw, BT_W) * ceil_div(im.rh, BT_H);
        }
      if (bands > 1) {
        if ((rc = ctx_alloc(ctx, (void**)&d_planes, n_planes * sizeof(MbPlane)))) goto done;
        if ((rc = ctx_alloc(ctx, (void**)&d_span, n_planes * sizeof(int2)))) goto done;
        if ((rc = ctx_put(ctx, d_planes, planes.data(), n_planes * sizeof(MbPlane)))) goto done;
        if ((rc = ctx_put(ctx, d_span, span.data(), n_planes * sizeof(int2)))) goto done;
      }
      // TMA descriptors: one per (plane, level buffer, distinct window half-width)
      std::vector<int> centers;
      for (auto& bt : level_taps)
        if ((bt.center == 6 || bt.center == 9) && std::find(centers.begin(), centers.end(), bt.center) == centers.end())
          centers.push_back(bt.center);
      if (!centers.empty()) {
        std::vector<TmaDesc> maps((size_t)centers.size() * 2 * n_planes);
        for (size_t ci = 0; ci < centers.size(); ++ci)
          for (int buf = 0; buf < 2; ++buf)
            for (int q = 0; q < n_planes; ++q) {
              const MbPlane& pl = planes[q];
              const int C = centers[ci];
              unsigned long long dims[2] = {(unsigned long long)pl.w, (unsigned long long)pl.h};
              unsigned long long strides[1] = {(unsigned long long)pl.pitch * sizeof(float)};
              unsigned box[2] = {(unsigned)(BT_W + 2 * ((C + 3) & ~3)), (unsigned)(BT_H + 2 * C)};
              if ((rc = ctx_tma_encode(ctx, &maps[(ci * 2 + buf) * n_planes + q], (buf ? d_next : d_cur) + pl.off, 2, dims,
                                       strides, box)))
                goto done;
            }
        if ((rc = ctx_alloc(ctx, (void**)&d_maps, maps.size() * sizeof(TmaDesc)))) goto done;
        if ((rc = ctx_put(ctx, d_maps, maps.data(), maps.size() * sizeof(TmaDesc)))) goto done;
      }
      dim3 gr(ceil_div(job.max_rw, 32), ceil_div(job.max_rh, 8), n);
      dim3 gw(ceil_div(tw, 32), ceil_div(wrow1 - wrow0, 8)), gs(ceil_div(tw, 32), ceil_div(row1 - row0, 8));
      BL_LAUNCH(ctx, "k_mb_first_level", k_mb_first_level, gr, b, 0, d_imgs, job.g, d_cur, d_mask);
      BL_LAUNCH(ctx, "k_mb_weight_argmax", k_mb_weight_argmax, gw, b, 0, d_imgs, n, d_cur, tw, wrow0, wrow1);
      int buf = 0;     // which level buffer `d_cur` currently is (0: the first allocation)
      for (int level = 0; level < bands; ++level) {
        int is_last = level == bands - 1;
        if (!is_last) {
          const BlurTaps& bt = level_taps[level];
          const int c = bt.center;
          ctx->launches++;
          if (ctx->profiling) ctx_prof_begin(ctx, "k_mb_blur");
          auto ci = std::find(centers.begin(), centers.end(), c);
          if (ci != centers.end()) {
            const TmaDesc* maps = d_maps + ((size_t)(ci - centers.begin()) * 2 + buf) * n_planes;
            const int grid = std::min(n_tiles, ctx->num_sms * 4);
            e = c == 6 ? launch_mb_blur_tma<6>(ctx, grid, d_planes, d_span, n_planes, n_tiles, maps, d_next, bt)
                       : launch_mb_blur_tma<9>(ctx, grid, d_planes, d_span, n_planes, n_tiles, maps, d_next, bt);
          } else {
            const size_t smem = sizeof(float) * ((size_t)(MB_TH + 2 * c) * (MB_TW + 2 * c) + (size_t)MB_TH * (MB_TW + 2 * c));
            if (smem > 200 * 1024) { rc = ctx_fail(ctx, PANO_ERR_INVALID, "blend: gaussian window %d too wide", 2 * c + 1); goto done; }
            e = cudaFuncSetAttribute(k_mb_blur_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) {
              k_mb_blur_generic<<<n_tiles, 256, smem, ctx->stream>>>(d_planes, d_span, n_planes, d_cur, d_next, bt);
              e = cudaGetLastError();
            }
          }
          if (ctx->profiling) ctx_prof_end(ctx);
          if (e != cudaSuccess) { rc = ctx_cuda(ctx, e, "k_mb_blur"); goto done; }
        }
        BL_LAUNCH(ctx, "k_mb_accumulate", k_mb_accumulate, gs, b, 0, d_imgs, n, d_cur, d_next, d_mask, level == 0 ? 1 : 0, is_last,
                  d_out, d_tmask, tw, row0, row1);
        if (!is_last) { std::swap(d_cur, d_next); buf ^= 1; }
      }
    }
  }
done:
  ctx_free(ctx, d_imgs); ctx_free(ctx, d_tab); ctx_free(ctx, d_cur); ctx_free(ctx, d_next);
  ctx_free(ctx, d_mask); ctx_free(ctx, d_tmask); ctx_free(ctx, d_planes); ctx_free(ctx, d_span); ctx_free(ctx, d_maps);
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
