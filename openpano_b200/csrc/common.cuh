// common.cuh — context, launch/profiling helpers and the bit-exact device math
// shared by every kernel file of libpano_b200.so.
//
// Numerical contract: every float/double operation on the SIFT,
// match-recheck, warp and blend paths is written as the reference's compiled
// code executes it (x86-64, -ffp-contract=off): no FMA contraction (the library
// is compiled with --fmad=false), IEEE division and square root, the same
// float<->double promotions, and libm calls replaced by the exact algorithms
// glibc 2.39 runs (ARM optimized-routines expf / sinf / cosf; hypotf as
// sqrt of the double sum) so results are bit-identical, not just close.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include <unordered_map>

#include "../../include/pano_b200.h"

// ------------------------------------------------------------------ context

struct ProfEvent {
  const char* name;
  cudaEvent_t start, stop;
};

struct pano_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  cudaMemPool_t pool = nullptr;   // this context's own stream-ordered pool (see ctx_alloc)
  // Freed blocks kept for the next request of (about) the same size: see ctx_alloc
  struct CachedBlock { void* p; unsigned long long stamp; };
  std::multimap<size_t, CachedBlock> cache;        // size -> free block
  std::unordered_map<void*, size_t> live;           // blocks handed out -> their true size
  size_t cached_bytes = 0, cache_limit = (size_t)32 << 30;   // one 64-view multiband job parks ~8 GB; the GPU has 180
  unsigned long long cache_stamp = 0;
  std::string err;
  bool profiling = false;
  std::vector<ProfEvent> prof_pending;
  std::vector<cudaEvent_t> event_pool;
  std::map<std::string, std::pair<int, double>> prof_acc;  // name -> (launches, ms)
  long long launches = 0;
  int last_match_exact_rows = 0;   // rows the last match call had to decide exactly (gathered pass)
  int last_match_nominated_rows = 0;   // columns on demand: rows of the larger sets nominated on request
  int last_match_full_rescans = 0; // of those, rows that needed a scan of every target
  int num_sms = 148;
  // cudaFuncSetAttribute is per device: remembered per context, never per process
  bool attr_tc = false, attr_match = false;
  int sift_cap = 0;                // per-image list capacity SIFT batches start with (grows on overflow, sticky)
  void* tma_encode = nullptr;      // cuTensorMapEncodeTiled, resolved through the runtime (no -lcuda)
  // pinned host staging (grown on demand)
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  void* pinned2 = nullptr;
  size_t pinned2_bytes = 0;
  // pinned + device-mapped ring for small host<->device moves done by SM kernels
  char* ring = nullptr;
  size_t ring_cap = 0, ring_off = 0;
  // recycled pinned blocks for per-featureset count read-backs (cudaHostAlloc / cudaFreeHost
  // are slow and cudaFreeHost synchronises the whole device)
  std::vector<std::pair<void*, size_t>> small_pinned;
  std::vector<cudaEvent_t> sync_events;
  // completion markers: a word in pinned host memory the stream writes sequence numbers to
  volatile unsigned* flag = nullptr;
  unsigned flag_seq = 0;
};

// Every entry point makes its context's device current first: host threads other than the
// creating one start on device 0 (a stitch lane's thread on rank 1 would otherwise issue
// its copies against the wrong device).
static inline void ctx_enter(const pano_ctx* ctx) { if (ctx) cudaSetDevice(ctx->device); }

int  ctx_fail(pano_ctx* ctx, int code, const char* fmt, ...);
int  ctx_cuda(pano_ctx* ctx, cudaError_t e, const char* what);
// Stream-ordered device memory from the CONTEXT'S OWN pool.  Contexts sharing the device's
// default pool hand each other freed blocks, and the allocator then makes the taking
// stream wait for the giving stream ("internal dependencies"): concurrent stitch lanes
// drifted from 2.3 to 5+ ms per job as their arenas started to cross over.
// On top of the pool sits a per-context cache of freed blocks, matched by size (best fit within
// 25 %): a stitch job asks for the same ~40 sizes every time, up to a 0.9 GB pyramid arena, and
// cudaMallocFromPoolAsync was seen to block the host for 0.1 - 1.5 s a few times per second when
// it had to re-arrange the pool's mappings for such a request — with the allocator's lock held,
// so every other lane of the process stalled with it (profiles/r02s_e2e_pause_probe.txt).
// Everything a context allocates is used on its one stream, so handing a block freed after its
// last enqueued use to the next request is ordered by the stream itself.  PANO_CACHE_MB bounds
// the cache (default 8192, 0 = off); pano_trim() gives the cached blocks back to the pool.
int  ctx_alloc(pano_ctx* ctx, void** p, size_t bytes);
void ctx_cache_release(pano_ctx* ctx, size_t keep_bytes);
void ctx_free(pano_ctx* ctx, void* p);
void* ctx_pinned(pano_ctx* ctx, size_t bytes);   // staging buffer A (inputs)
void* ctx_pinned2(pano_ctx* ctx, size_t bytes);  // staging buffer B (results)
// Small host<->device moves that stay OFF the copy engines: a big image upload or
// mosaic download queued on another stream of the same device would otherwise
// delay every tiny metadata copy queued behind it on the same engine, and with it
// the kernels that depend on it.  ctx_ring reserves space in a pinned, device-
// mapped ring (valid until the ring wraps, which synchronises the stream);
// ctx_fetch / ctx_store move words with a small kernel that addresses the host
// memory directly (UVA); ctx_put = ring + memcpy + fetch; ctx_zero fills zeros.
// Host waits that SPIN on cudaEventQuery instead of sleeping in the driver: the hot
// path has two short waits per step (feature counts, match decisions) and a
// descheduled host thread on a busy machine costs milliseconds.
cudaError_t ctx_spin_event(cudaEvent_t ev);
cudaError_t ctx_spin_stream(pano_ctx* ctx);
// Short waits poll a marker word in pinned host memory instead of the driver: a host
// thread spinning in cudaEventQuery holds the context lock most of the time and starves
// another thread's kernel launches (two stitch lanes on one GPU went bimodal, 2.5 vs 5 ms).
// ctx_signal queues "write the next sequence number" on the ctx stream and returns it;
// ctx_wait_signal spins until the word has reached it (stream sync after a long timeout
// so that a device fault still surfaces).
cudaError_t ctx_signal(pano_ctx* ctx, unsigned* token);
cudaError_t ctx_wait_signal(pano_ctx* ctx, unsigned token);
void* ctx_ring(pano_ctx* ctx, size_t bytes);
void* ctx_small_pinned_get(pano_ctx* ctx, size_t bytes, size_t* cap);
void ctx_small_pinned_put(pano_ctx* ctx, void* p, size_t cap);
cudaEvent_t ctx_sync_event_get(pano_ctx* ctx);
void ctx_sync_event_put(pano_ctx* ctx, cudaEvent_t e);
int  ctx_fetch(pano_ctx* ctx, void* d_dst, const void* h_pinned_src, size_t bytes);
int  ctx_store(pano_ctx* ctx, void* h_pinned_dst, const void* d_src, size_t bytes);
int  ctx_put(pano_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int  ctx_zero(pano_ctx* ctx, void* d_dst, size_t bytes);
// up to CTX_MAX_SEGS moves in one launch; h_src[i] == nullptr zero-fills d_dst[i]
#define CTX_MAX_SEGS 8
int  ctx_put_many(pano_ctx* ctx, int n, void* const* d_dst, const void* const* h_src, const size_t* bytes);
int  ctx_store_many(pano_ctx* ctx, int n, void* const* h_pinned_dst, const void* const* d_src, const size_t* bytes);
int  ctx_copy_blocks(pano_ctx* ctx, int n, void* const* dst, const void* const* src, const size_t* bytes);
void ctx_prof_begin(pano_ctx* ctx, const char* name);
void ctx_prof_end(pano_ctx* ctx);

// Diagnostics (PANO_TRACE_SLOW_MS=<ms>): report every wrapped driver-facing call that blocks the host
// longer than the threshold — which call a stall sits in, not how long the GPU takes.
double pano_now_ms();
extern double g_trace_slow_ms;       // 0 = off
void pano_trace_slow(const char* what, double ms);
struct SlowCall {
  const char* what; double t0;
  explicit SlowCall(const char* w) : what(w), t0(g_trace_slow_ms > 0 ? pano_now_ms() : 0.0) {}
  ~SlowCall() { if (g_trace_slow_ms > 0) { const double d = pano_now_ms() - t0; if (d > g_trace_slow_ms) pano_trace_slow(what, d); } }
};

#define PANO_CUDA(ctx, call)                                          \
  do {                                                                \
    cudaError_t _e;                                                   \
    { SlowCall _sc(#call); _e = (call); }                             \
    if (_e != cudaSuccess) return ctx_cuda((ctx), _e, #call);         \
  } while (0)

// Launch a kernel on the ctx stream, counted and (optionally) event-timed.
#define PANO_LAUNCH(ctx, name, kernel, grid, block, smem, ...)                      \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    if ((ctx)->profiling) ctx_prof_begin((ctx), (name));                            \
    { SlowCall _sc(name); kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); } \
    if ((ctx)->profiling) ctx_prof_end((ctx));                                      \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) return ctx_cuda((ctx), _e, name);                        \
  } while (0)

// A TMA descriptor (CUtensorMap: 128 bytes, 64-byte aligned) for cp.async.bulk.tensor tile
// loads: an f32 tensor of `rank` dimensions (innermost first), strides_bytes[rank-1] for
// dimensions 1.., box = tile extent per dimension.  Encoded on the host, copied to device
// memory and handed to the kernels by pointer.
struct __align__(64) TmaDesc { unsigned long long opaque[16]; };
int ctx_tma_encode(pano_ctx* ctx, TmaDesc* out, void* base, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned* box);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Gaussian kernel exactly as GaussCache builds it (feature/gaussian.cc:17-40);
// host side, taps[0] is the tap at -center.  Returns kw.
int host_gauss_kernel(float sigma, int window_factor, float* taps, int cap);

// -------------------------------------------------------------- device math
#ifdef __CUDACC__

#define PANO_PI 3.14159265358979323846
#define PANO_PI_2 1.57079632679489661923
#define PANO_SQRT1_2 0.70710678118654752440

// lib/utils.hh:27 between(a,b,c)
#define DBETWEEN(a, b, c) (((a) >= (b)) && ((a) <= (c) - 1))

// glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c, ARM optimized-routines):
// x*N/ln2 = k + r, exp(x) = 2^(k/N) * p(r), N = 32, evaluated in double.
__constant__ uint64_t c_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

// `tab` is a 32-entry copy of c_exp2f_tab in shared memory (see load_exp2f_tab):
// lanes index it with different k, which constant memory would serialise.
__device__ __forceinline__ void load_exp2f_tab(uint64_t* s_tab, int tid) {
  if (tid < 32) s_tab[tid] = c_exp2f_tab[tid];
}

__device__ __forceinline__ float glibc_expf(float x, const uint64_t* __restrict__ tab) {
  // special ranges of the libm routine (|x| >= 88): only the underflow side can
  // occur here (arguments are -(d^2)/denom <= 0).
  if (x < -0x1.9fe368p6f) return 0.0f;
  const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32,
               C2 = 0x1.62e42ff0c52d6p-1 / 32;
  double xd = (double)x;
  double z = InvLn2N * xd;
  double kd = z + SHIFT;
  uint64_t ki = (uint64_t)__double_as_longlong(kd);
  kd -= SHIFT;
  double r = z - kd;
  uint64_t t = tab[ki & 31];
  t += ki << 47;
  double s = __longlong_as_double((long long)t);
  z = C0 * r + C1;
  double r2 = r * r;
  double y = C2 * r + 1;
  y = z * r2 + y;
  y = y * s;
  return (float)y;
}

// glibc 2.39 sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h): valid for |y| < 120.
__device__ __forceinline__ float glibc_sincos_poly(double x, double x2, bool neg_table, int n) {
  // table[1] negates the cosine coefficients only
  const double sgn = neg_table ? -1.0 : 1.0;
  if ((n & 1) == 0) {
    const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
    double x3 = x * x2;
    double s1 = s2c + x2 * s3c;
    double x7 = x3 * x2;
    double s = x + x3 * s1c;
    return (float)(s + x7 * s1);
  } else {
    const double c0 = sgn * 0x1p0, c1c = sgn * -0x1.ffffffd0c621cp-2, c2c = sgn * 0x1.55553e1068f19p-5,
                 c3c = sgn * -0x1.6c087e89a359dp-10, c4c = sgn * 0x1.99343027bf8c3p-16;
    double x4 = x2 * x2;
    double c2 = c3c + x2 * c4c;
    double c1 = c1c + x2 * c2c;
    double x6 = x4 * x2;
    double c = c0 + x2 * c1;
    return (float)(c + x6 * c2);
  }
}

__device__ __forceinline__ uint32_t glibc_abstop12(float x) { return (__float_as_uint(x) >> 20) & 0x7ff; }

__device__ __forceinline__ void glibc_sincosf(float y, float* sn, float* cs) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  double x = (double)y;
  if (glibc_abstop12(y) < glibc_abstop12(0x1.921FB6p-1f)) {
    double x2 = x * x;
    if (glibc_abstop12(y) < glibc_abstop12(0x1p-12f)) { *sn = y; *cs = 1.0f; return; }
    *sn = glibc_sincos_poly(x, x2, false, 0);
    *cs = glibc_sincos_poly(x, x2, false, 1);
    return;
  }
  double r = x * hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  x = x - n * hpi;
  const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // {1,-1,-1,1}
  bool neg = (n & 2) != 0;
  *sn = glibc_sincos_poly(x * sign, x * x, neg, n);
  *cs = glibc_sincos_poly(x * sign, x * x, neg, n ^ 1);
}

// glibc 2.39 hypotf == (float)sqrt((double)x*x + (double)y*y) (SURVEY.md §7
// hard part 2; re-verified here on 2e7 random pairs).
__device__ __forceinline__ float glibc_hypotf(float x, float y) {
  double dx = (double)x, dy = (double)y;
  return (float)sqrt(dx * dx + dy * dy);
}

// feature/dog.cc:22-37 fast_atan
__device__ __forceinline__ float fast_atan(float y, float x) {
  float absx = fabsf(x), absy = fabsf(y);
  float m = absx > absy ? absx : absy;
  if ((double)m < 1e-6) return (float)(-PANO_PI);
  float a = (absy < absx ? absy : absx) / m;
  float s = a * a;
  double sd = (double)s, ad = (double)a;
  float r = (float)(((-0.0464964749 * sd + 0.15931422) * sd - 0.327622764) * sd * ad + ad);
  if (absy > absx) r = (float)(PANO_PI_2 - (double)r);
  if (x < 0) r = (float)(PANO_PI - (double)r);
  if (y < 0) r = -r;
  return r;
}

// feature/dog.cc:60-94 cal_mag_ort for one INTERIOR pixel (1<=x<=w-2, 1<=y<=h-2);
// border pixels have mag=0, ort=pi and are never visited by the callers.
__device__ __forceinline__ void mag_ort_at(const float* __restrict__ img, int w, int x, int y,
                                           float* mag, float* ort) {
  const float* row = img + (size_t)y * w;
  float dy = row[x + w] - row[x - w];
  float dx = row[x + 1] - row[x - 1];
  *mag = glibc_hypotf(dx, dy);
  *ort = (float)((double)fast_atan(dy, dx) + PANO_PI);
}

// lib/imgproc.cc:135-156 interpolate; returns false for Color::NO
__device__ __forceinline__ bool interpolate_rgb(const float* __restrict__ img, int w, int h, float r,
                                                float c, float* o0, float* o1, float* o2) {
  int fr = (int)floorf(r), fc = (int)floorf(c);
  if (fr < 0 || fc < 0 || fc + 1 >= w || fr + 1 >= h) return false;
  r -= (float)fr;
  c -= (float)fc;
  const float* p00 = img + ((size_t)fr * w + fc) * 3;
  const float* p10 = p00 + (size_t)w * 3;
  // all twelve samples first (the four positions are in range), tests afterwards: the
  // loads overlap instead of each Color::NO test waiting on its own load
  const float q00 = __ldg(p00), q01 = __ldg(p00 + 1), q02 = __ldg(p00 + 2);
  const float q03 = __ldg(p00 + 3), q04 = __ldg(p00 + 4), q05 = __ldg(p00 + 5);
  const float q10 = __ldg(p10), q11 = __ldg(p10 + 1), q12 = __ldg(p10 + 2);
  const float q13 = __ldg(p10 + 3), q14 = __ldg(p10 + 4), q15 = __ldg(p10 + 5);
  if (q00 < 0 || q10 < 0 || q13 < 0 || q03 < 0) return false;
  float w00 = (1 - r) * (1 - c), w10 = r * (1 - c), w11 = r * c, w01 = (1 - r) * c;
  float a0 = 0.f + q00 * w00, a1 = 0.f + q01 * w00, a2 = 0.f + q02 * w00;
  a0 += q10 * w10; a1 += q11 * w10; a2 += q12 * w10;
  a0 += q13 * w11; a1 += q14 * w11; a2 += q15 * w11;
  a0 += q03 * w01; a1 += q04 * w01; a2 += q05 * w01;
  *o0 = a0; *o1 = a1; *o2 = a2;
  return true;
}

#endif  // __CUDACC__
