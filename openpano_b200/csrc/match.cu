// match.cu — exact pairwise descriptor matching.
//
// Replaces PairWiseMatcher::match (feature/matcher.cc:90-135) with the exact rule
// of FeatureMatcher::match (feature/matcher.cc:15-71), the parity contract
// (SURVEY.md §8c).  Distances are feature/dist.cc:22-57's SSE branch bit for bit:
// four lane accumulators over the 32 4-float steps, summed (l0+l1)+(l2+l3), no
// FMA.  The reference's rule
//     loop k over the smaller set: exact top-2 of row k (lowest index wins);
//     reject if min > R*next; next = min(next, min_{kk!=k} d(best, kk)); reject again
// is evaluated as two symmetric top-2 reductions (rows of A over B, rows of B over
// A) followed by a per-row decision: min_{kk!=k} d(j, kk) is column j's second
// minimum when its argmin is k, and its minimum otherwise.
#include "sift.cuh"
#include <float.h>
#include <string.h>
#include <algorithm>

#define MT 64            // rows of the query tile and of the target tile
#define MT_STRIDE 132    // padded row stride in floats (conflict-free float4 reads)
#define MT_THREADS 256

struct MatchTask {       // one top-2 reduction: rows [q_row0, q_row0+MT) of Q against all of T
  long long q_base, t_base;  // first descriptor row of the sets inside the featureset
  int q_n, t_n;
  int q_row0;
  long long res_off;     // where this query set's results start
};

struct Top2 { float mn; float second; int idx; int pad; };

__device__ __forceinline__ void top2_merge(float& mn, int& idx, float& sec, float m2, int i2, float s2) {
  if (m2 < mn || (m2 == mn && i2 < idx)) { sec = fminf(mn, s2); mn = m2; idx = i2; }
  else sec = fminf(sec, m2);
}

__global__ void __launch_bounds__(MT_THREADS)
k_match_top2(const float* __restrict__ desc, const MatchTask* __restrict__ tasks, Top2* __restrict__ res) {
  extern __shared__ float sm[];
  float* sq = sm;                       // [MT][MT_STRIDE]
  float* st = sm + MT * MT_STRIDE;      // [MT][MT_STRIDE]
  const MatchTask tk = tasks[blockIdx.x];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* Q = desc + tk.q_base * 128;
  const float* T = desc + tk.t_base * 128;

  for (int i = tid; i < MT * 32; i += MT_THREADS) {   // 32 float4 per row
    int r = i >> 5, c4 = i & 31;
    int gr = tk.q_row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < tk.q_n) v = __ldg((const float4*)(Q + (size_t)gr * 128) + c4);
    *(float4*)(sq + r * MT_STRIDE + c4 * 4) = v;
  }

  float bmn[4], bsec[4];
  int bidx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { bmn[i] = FLT_MAX; bsec[i] = FLT_MAX; bidx[i] = 0x7fffffff; }

  for (int t0 = 0; t0 < tk.t_n; t0 += MT) {
    __syncthreads();
    for (int i = tid; i < MT * 32; i += MT_THREADS) {
      int r = i >> 5, c4 = i & 31;
      int gr = t0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < tk.t_n) v = __ldg((const float4*)(T + (size_t)gr * 128) + c4);
      *(float4*)(st + r * MT_STRIDE + c4 * 4) = v;
    }
    __syncthreads();
    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int l = 0; l < 4; ++l) acc[i][j][l] = 0.f;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *(const float4*)(sq + (ty + 16 * i) * MT_STRIDE + k * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *(const float4*)(st + (tx + 16 * j) * MT_STRIDE + k * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d0 = a[i].x - b[j].x, d1 = a[i].y - b[j].y, d2 = a[i].z - b[j].z, d3 = a[i].w - b[j].w;
          acc[i][j][0] += d0 * d0; acc[i][j][1] += d1 * d1;
          acc[i][j][2] += d2 * d2; acc[i][j][3] += d3 * d3;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int col = t0 + tx + 16 * j;
        if (col < tk.t_n) {
          float d = (acc[i][j][0] + acc[i][j][1]) + (acc[i][j][2] + acc[i][j][3]);
          // ascending col within a thread: strict < keeps the lowest index
          if (d < bmn[i]) { bsec[i] = bmn[i]; bmn[i] = d; bidx[i] = col; }
          else if (d < bsec[i]) bsec[i] = d;
        }
      }
  }
  // merge the 16 threads (tx) that share a query row: they sit in one half-warp
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float mn = bmn[i], sec = bsec[i];
    int idx = bidx[i];
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, mn, off);
      float s2 = __shfl_xor_sync(0xffffffffu, sec, off);
      int i2 = __shfl_xor_sync(0xffffffffu, idx, off);
      top2_merge(mn, idx, sec, m2, i2, s2);
    }
    int row = tk.q_row0 + ty + 16 * i;
    if (tx == 0 && row < tk.q_n) {
      Top2 o; o.mn = mn; o.second = sec; o.idx = idx; o.pad = 0;
      res[tk.res_off + row] = o;
    }
  }
}

struct PairMeta {
  long long resA_off, resB_off;  // Top2 of the smaller set's rows / of the larger set's rows
  int n_small, n_large;
  long long out_off;             // per-row decision of the smaller set
};

// Decision per row k of the smaller set (matcher.cc:49-66).
__global__ void k_match_decide(const PairMeta* __restrict__ pairs, const Top2* __restrict__ res,
                               float ratio_sqr, int* __restrict__ out, int* __restrict__ total) {
  const PairMeta pm = pairs[blockIdx.y];
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= pm.n_small) return;
  Top2 r = res[pm.resA_off + k];
  int result = -1;
  if (pm.n_large > 0 && !(r.mn > ratio_sqr * r.second)) {
    Top2 c = res[pm.resB_off + r.idx];
    float colmin = (c.idx == k) ? c.second : c.mn;
    float next_min = fminf(r.second, colmin);
    if (!(r.mn > ratio_sqr * next_min)) result = r.idx;
  }
  out[pm.out_off + k] = result;
  if (result >= 0) atomicAdd(total, 1);
}

// ------------------------------------------------------------------ host driver

struct MatchPlan {
  std::vector<MatchTask> tasks;
  std::vector<PairMeta> pairs;
  std::vector<char> rev;   // pair was swapped (first image is the larger set)
  long long res_total = 0, out_total = 0;
};

static void plan_pair(MatchPlan& pl, long long baseA, int nA, long long baseB, int nB) {
  // matcher.cc:21-29: loop over the smaller one; rev = l1 > l2
  bool rev = nA > nB;
  long long bs = rev ? baseB : baseA, bl = rev ? baseA : baseB;
  int ns = rev ? nB : nA, nl = rev ? nA : nB;
  PairMeta pm;
  pm.n_small = ns; pm.n_large = nl;
  pm.resA_off = pl.res_total; pl.res_total += ns;
  pm.resB_off = pl.res_total; pl.res_total += nl;
  pm.out_off = pl.out_total; pl.out_total += ns;
  for (int r0 = 0; r0 < ns; r0 += MT) pl.tasks.push_back(MatchTask{bs, bl, ns, nl, r0, pm.resA_off});
  for (int r0 = 0; r0 < nl; r0 += MT) pl.tasks.push_back(MatchTask{bl, bs, nl, ns, r0, pm.resB_off});
  pl.pairs.push_back(pm);
  pl.rev.push_back(rev ? 1 : 0);
}

// Runs the plan; leaves per-row decisions in *d_out (caller frees) and the total in *d_total.
static int run_plan(pano_ctx* ctx, const float* d_desc, const MatchPlan& pl, float ratio, int** d_out, int** d_total) {
  *d_out = nullptr; *d_total = nullptr;
  MatchTask* d_tasks = nullptr; PairMeta* d_pairs = nullptr; Top2* d_res = nullptr;
  int rc = 0;
  size_t bt = pl.tasks.size() * sizeof(MatchTask), bp = pl.pairs.size() * sizeof(PairMeta);
  if ((rc = ctx_alloc(ctx, (void**)&d_tasks, bt)) || (rc = ctx_alloc(ctx, (void**)&d_pairs, bp)) ||
      (rc = ctx_alloc(ctx, (void**)&d_res, std::max<long long>(pl.res_total, 1) * sizeof(Top2))) ||
      (rc = ctx_alloc(ctx, (void**)d_out, std::max<long long>(pl.out_total, 1) * sizeof(int))) ||
      (rc = ctx_alloc(ctx, (void**)d_total, sizeof(int)))) {
    ctx_free(ctx, d_tasks); ctx_free(ctx, d_pairs); ctx_free(ctx, d_res);
    return rc;
  }
  char* stg = (char*)ctx_pinned2(ctx, bt + bp + 16);
  if (!stg) return ctx_fail(ctx, PANO_ERR_CUDA, "pinned alloc failed");
  if (bt) memcpy(stg, pl.tasks.data(), bt);
  if (bp) memcpy(stg + bt, pl.pairs.data(), bp);
  cudaError_t e = cudaSuccess;
  if (bt) e = cudaMemcpyAsync(d_tasks, stg, bt, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && bp) e = cudaMemcpyAsync(d_pairs, stg + bt, bp, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(*d_total, 0, sizeof(int), ctx->stream);
  if (e != cudaSuccess) { ctx_free(ctx, d_tasks); ctx_free(ctx, d_pairs); ctx_free(ctx, d_res); return ctx_cuda(ctx, e, "match upload"); }
  const size_t smem = (size_t)2 * MT * MT_STRIDE * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(k_match_top2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ctx_cuda(ctx, e, "cudaFuncSetAttribute(k_match_top2)");
    attr_set = true;
  }
  if (!pl.tasks.empty())
    PANO_LAUNCH(ctx, "k_match_top2", k_match_top2, (unsigned)pl.tasks.size(), MT_THREADS, smem, d_desc, d_tasks, d_res);
  int max_small = 0;
  for (auto& pm : pl.pairs) max_small = std::max(max_small, pm.n_small);
  if (max_small > 0) {
    dim3 g(ceil_div(max_small, 256), (unsigned)pl.pairs.size());
    PANO_LAUNCH(ctx, "k_match_decide", k_match_decide, g, 256, 0, d_pairs, d_res, ratio * ratio, *d_out, *d_total);
  }
  ctx_free(ctx, d_tasks); ctx_free(ctx, d_pairs); ctx_free(ctx, d_res);
  return PANO_OK;
}

static int build_plan(pano_ctx* ctx, pano_featureset* fs, int n_pairs, const int* ij, MatchPlan& pl) {
  int rc = featureset_sync_counts(fs);
  if (rc) return rc;
  for (int k = 0; k < n_pairs; ++k) {
    int i = ij[2 * k], j = ij[2 * k + 1];
    if (i < 0 || j < 0 || i >= fs->n_images || j >= fs->n_images)
      return ctx_fail(ctx, PANO_ERR_INVALID, "pair %d: image index out of range", k);
    plan_pair(pl, fs->base[i], fs->h_count[i], fs->base[j], fs->h_count[j]);
  }
  return PANO_OK;
}

extern "C" {

int pano_match_pairs(pano_ctx* ctx, pano_featureset* fs, int n_pairs, const int* ij, const pano_params* p,
                     pano_matches* out) {
  if (!ctx || !fs || !out || n_pairs < 0 || (n_pairs && !ij) || !p) return PANO_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  MatchPlan pl;
  int rc = build_plan(ctx, fs, n_pairs, ij, pl);
  if (rc) return rc;
  int *d_out = nullptr, *d_total = nullptr;
  rc = run_plan(ctx, fs->d_desc, pl, p->match_reject_next_ratio, &d_out, &d_total);
  if (rc) { ctx_free(ctx, d_out); ctx_free(ctx, d_total); return rc; }
  std::vector<int> h_out(std::max<long long>(pl.out_total, 1));
  cudaError_t e = cudaMemcpyAsync(h_out.data(), d_out, pl.out_total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, d_out); ctx_free(ctx, d_total);
  if (e != cudaSuccess) return ctx_cuda(ctx, e, "match download");
  out->n_pairs = n_pairs;
  out->count = (int*)calloc(std::max(n_pairs, 1), sizeof(int));
  out->offset = (int*)calloc(n_pairs + 1, sizeof(int));
  int total = 0;
  for (int k = 0; k < n_pairs; ++k) {
    const PairMeta& pm = pl.pairs[k];
    int c = 0;
    for (int r = 0; r < pm.n_small; ++r) c += h_out[pm.out_off + r] >= 0;
    out->count[k] = c; out->offset[k] = total; total += c;
  }
  out->offset[n_pairs] = total;
  out->idx = (int*)calloc(std::max(total, 1) * 2, sizeof(int));
  for (int k = 0; k < n_pairs; ++k) {
    const PairMeta& pm = pl.pairs[k];
    int* dst = out->idx + 2 * out->offset[k];
    for (int r = 0; r < pm.n_small; ++r) {
      int j = h_out[pm.out_off + r];
      if (j < 0) continue;
      if (pl.rev[k]) { dst[0] = j; dst[1] = r; } else { dst[0] = r; dst[1] = j; }  // MatchData::reverse
      dst += 2;
    }
  }
  return PANO_OK;
}

void pano_matches_free(pano_matches* m) {
  if (!m) return;
  free(m->count); free(m->offset); free(m->idx);
  memset(m, 0, sizeof(*m));
}

int pano_match_pairs_dev(pano_ctx* ctx, pano_featureset* fs, int n_pairs, const int* ij, const pano_params* p,
                         int* total_matches) {
  if (!ctx || !fs || n_pairs < 0 || (n_pairs && !ij) || !p || !total_matches) return PANO_ERR_INVALID;
  MatchPlan pl;
  int rc = build_plan(ctx, fs, n_pairs, ij, pl);
  if (rc) return rc;
  int *d_out = nullptr, *d_total = nullptr;
  rc = run_plan(ctx, fs->d_desc, pl, p->match_reject_next_ratio, &d_out, &d_total);
  if (rc) { ctx_free(ctx, d_out); ctx_free(ctx, d_total); return rc; }
  cudaError_t e = cudaMemcpyAsync(total_matches, d_total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  ctx_free(ctx, d_out); ctx_free(ctx, d_total);
  if (e != cudaSuccess) return ctx_cuda(ctx, e, "match total download");
  return PANO_OK;
}

int pano_match_bruteforce(pano_ctx* ctx, const float* a, int n, const float* b, int m, const pano_params* p,
                          int* pairs_out, int* n_pairs_out) {
  if (!ctx || n < 0 || m < 0 || !p || !pairs_out || !n_pairs_out) return PANO_ERR_INVALID;
  *n_pairs_out = 0;
  if (n == 0 || m == 0) return PANO_OK;
  int cnt[2] = {n, m};
  const float* ds[2] = {a, b};
  pano_featureset* fs = nullptr;
  int rc = pano_featureset_upload(ctx, 2, cnt, ds, nullptr, &fs);
  if (rc) return rc;
  int ij[2] = {0, 1};
  pano_matches mt;
  rc = pano_match_pairs(ctx, fs, 1, ij, p, &mt);
  pano_featureset_free(fs);
  if (rc) return rc;
  *n_pairs_out = mt.count[0];
  memcpy(pairs_out, mt.idx, sizeof(int) * 2 * mt.count[0]);
  pano_matches_free(&mt);
  return PANO_OK;
}

}  // extern "C"
