// match_tc.cu — descriptor distance matrix on the 5th-gen tensor cores.
//
// The N x 128 . 128 x M contraction of the matcher (feature/matcher.cc:34-47 and
// :57-61: every query against every target, both directions) is the one GEMM of
// the hot path.  Here it runs as tcgen05.mma (kind::f16, fp32 accumulate in
// TMEM) and only NOMINATES: the exact fp32 rule of the reference is then decided
// by match.cu from certified bounds, with an exact re-scan of the few ambiguous
// rows, so the match pairs stay bit-identical.
//
// Operands.  k_tc_prep converts every descriptor x (f32[128]) into two fp16 rows
// of K = 144 (128 + one extra 16-wide k-step):
//     query  form: [ s*x        | 1, 1, n_hi, n_lo, 1, 0... ]
//     target form: [ -2*s*x     | n_hi, n_lo, 1, 1, 1, 0... ]     n = s^2*|x|^2
// so that  q . t = s^2 * |xq - xt|^2 + 1   (>= ~1: positive floats order like
// ints).  s is a power of two making n <= 1024 (s = 1/16 for RootSIFT, |x| = 512).
// Rows are stored PRE-BLOCKED in the UMMA canonical K-major no-swizzle layout (8x16-byte core
// matrices, SBO = 128 B): query rows in 128-row blocks (LBO = 2 KiB), target rows in 256-row
// tiles (LBO = 4 KiB), so one contiguous cp.async.bulk brings a block / tile into shared memory
// ready for the MMA.
//
// Kernel k_tc_pass, persistent, one CTA per SM walking tasks = (query block of 128 rows) x (a range
// of target tiles):
//   warp 0   producer : cp.async.bulk of target tiles (256 rows, 72 KiB) into a 2-stage ring,
//                       query blocks into two buffers
//   warp 1   MMA      : 9 tcgen05.mma (M128 N256 K16) per tile into one of two 256-column TMEM
//                       accumulator stages, tcgen05.commit -> mbarriers (N = 256: the query block
//                       is read from shared memory once per tile and k-step; two N = 128
//                       instructions per k-step measured 13 % slower at 100 k x 100 k)
//   warp 2   TMEM alloc/dealloc (512 columns)
//   warps 4-11 epilogue: two groups of four warps; group g owns the g-th 128-column half of
//                       every accumulator stage (its own full/empty barriers), so each SM
//                       sub-partition holds two epilogue warps and one computes while the
//                       other waits for its tcgen05.ld.  32 columns at a time; thread = query
//                       row keeps a running (min, argmin, second-min) with the column packed
//                       into the low mantissa bits (3.5 ALU instructions per element: one LOP3,
//                       2.5 VIMNMX / VIMNMX3); on long target sets a chunk whose 32-way minimum
//                       is not below the row's running second best is skipped (0.5 per element);
//                       the groups' results are merged through shared memory per task
#include "sift.cuh"
#include "match_tc.cuh"
#include <cuda_fp16.h>
#include <float.h>
#include <algorithm>

#define TC_KC 18                       // 16-byte k-chunks per row: 144 fp16
#define TC_BLOCK_BYTES (TC_KC * 2048)  // 128 rows x 144 fp16 = 36864 B
#define TC_LBO 2048u                   // byte stride between k-chunks (128-row query blocks)
#define TC_TLBO 4096u                  // the same for the 256-row target tiles
#define TC_SBO 128u                    // byte stride between 8-row groups
#define TC_TILE_BLOCKS 2               // target tile = 2 blocks = 256 rows
#define TC_STAGES 2
#define TC_THREADS 384
#define TC_EPI_GROUPS 2

// ------------------------------------------------------------------ prep

__global__ void k_tc_maxnorm(const float* __restrict__ desc, const TcImage* __restrict__ imgs, int n_img,
                             float* __restrict__ norms, unsigned* __restrict__ maxnorm_bits) {
  const int img = blockIdx.y;
  const TcImage im = imgs[img];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float n = 0.f;
  if (r < im.n) {
    const float4* p = (const float4*)(desc + (size_t)(im.row0 + r) * 128);
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) { float4 v = __ldg(p + k); acc += v.x * v.x; acc += v.y * v.y; acc += v.z * v.z; acc += v.w * v.w; }
    n = acc;
    norms[im.row0 + r] = n;
  }
  // block max -> global max (floats >= 0 order like unsigned ints)
  for (int off = 16; off; off >>= 1) n = fmaxf(n, __shfl_xor_sync(0xffffffffu, n, off));
  if ((threadIdx.x & 31) == 0 && n > 0.f) atomicMax(maxnorm_bits, __float_as_uint(n));
}

__device__ __forceinline__ float tc_scale_from_maxnorm(float maxn2) {
  // s = 2^-e with s^2 * maxn2 <= 1024
  if (!(maxn2 > 0.f)) return 1.f;
  int e = (int)ceilf(0.5f * log2f(maxn2 / 1024.f));
  return exp2f((float)-e);
}

// one thread per (row, form): writes 18 16-byte chunks into the blocked layout
__global__ void k_tc_prep(const float* __restrict__ desc, const float* __restrict__ norms,
                          const unsigned* __restrict__ maxnorm_bits, const TcImage* __restrict__ imgs,
                          unsigned char* __restrict__ qbuf, unsigned char* __restrict__ tbuf) {
  const TcImage im = imgs[blockIdx.y];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;   // padded row index
  if (r >= im.n_pad) return;
  const float s = tc_scale_from_maxnorm(__uint_as_float(*maxnorm_bits));
  const size_t blk = (size_t)(im.blk0 + r / 128);
  const int rr = r % 128;
  unsigned char* qd = qbuf + blk * TC_BLOCK_BYTES + (rr / 8) * TC_SBO + (rr % 8) * 16;
  // targets are laid out in 256-ROW tiles (one N = 256 MMA reads a whole tile: 32 row groups at SBO
  // per k-chunk, k-chunks TC_TLBO apart); n_pad is a multiple of 256, so tile t of an image starts
  // at block blk0 + 2 t either way
  const int tr = r % 256;
  unsigned char* td = tbuf + (size_t)(im.blk0 + (r / 256) * 2) * TC_BLOCK_BYTES + (tr / 8) * TC_SBO + (tr % 8) * 16;
  const bool real = r < im.n;
  const float4* p = (const float4*)(desc + (size_t)(im.row0 + (real ? r : 0)) * 128);
#pragma unroll 4
  for (int kc = 0; kc < 16; ++kc) {
    float4 a = real ? __ldg(p + 2 * kc) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 b = real ? __ldg(p + 2 * kc + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    __half2 q0 = __floats2half2_rn(a.x * s, a.y * s), q1 = __floats2half2_rn(a.z * s, a.w * s);
    __half2 q2 = __floats2half2_rn(b.x * s, b.y * s), q3 = __floats2half2_rn(b.z * s, b.w * s);
    const float m = -2.f * s;
    __half2 t0 = __floats2half2_rn(a.x * m, a.y * m), t1 = __floats2half2_rn(a.z * m, a.w * m);
    __half2 t2 = __floats2half2_rn(b.x * m, b.y * m), t3 = __floats2half2_rn(b.z * m, b.w * m);
    uint4 qv, tv;
    qv.x = *(unsigned*)&q0; qv.y = *(unsigned*)&q1; qv.z = *(unsigned*)&q2; qv.w = *(unsigned*)&q3;
    tv.x = *(unsigned*)&t0; tv.y = *(unsigned*)&t1; tv.z = *(unsigned*)&t2; tv.w = *(unsigned*)&t3;
    *(uint4*)(qd + (size_t)kc * TC_LBO) = qv;
    *(uint4*)(td + (size_t)kc * TC_TLBO) = tv;
  }
  // extra k-step: chunks 16 and 17
  float n = real ? norms[im.row0 + r] * s * s : 0.f;
  __half nh = __float2half_rn(n);
  __half nl = __float2half_rn(n - __half2float(nh));
  const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
  __half qx[8] = {one, one, nh, nl, one, zero, zero, zero};
  __half tx[8] = {nh, nl, one, one, one, zero, zero, zero};
  if (!real) {  // padded target rows must never win: q.t = 30000
    tx[0] = __float2half_rn(30000.f); tx[1] = zero; tx[2] = zero; tx[3] = zero; tx[4] = zero;
    qx[2] = zero; qx[3] = zero;
  }
  *(uint4*)(qd + (size_t)16 * TC_LBO) = *(uint4*)qx;
  *(uint4*)(td + (size_t)16 * TC_TLBO) = *(uint4*)tx;
  *(uint4*)(qd + (size_t)17 * TC_LBO) = make_uint4(0, 0, 0, 0);
  *(uint4*)(td + (size_t)17 * TC_TLBO) = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------ PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo) {
  // UMMA::SmemDescriptor (K-major, SWIZZLE_NONE): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ the GEMM + top-2 kernel

struct __align__(8) TcBarriers {
  uint64_t full[TC_STAGES], empty[TC_STAGES], acc_full[2][TC_EPI_GROUPS], acc_empty[2][TC_EPI_GROUPS], a_full[2], a_empty[2];
  uint32_t tmem_base;
  int merge[2][128][3];   // group 1 -> group 0 hand-over of (best, second, argmin), by task parity
};

// FILTER = false: running top-2 per query row (the nomination pass).
// FILTER = true : second pass over GATHERED rows only; every column whose score is
//                 within the row's threshold key is appended to that row's candidate
//                 slots (the exact kernel then decides among a handful of columns).
//
// PERSISTENT: a CTA walks tasks blockIdx.x, blockIdx.x + gridDim.x, ...; the three roles run the
// same task sequence on their own, coupled only through mbarriers whose phases follow two running
// counters (target tiles and tasks).  Nothing is re-initialised between tasks, so the producer is
// already fetching the next task's query block (two A buffers) and target tiles while the MMA and
// the epilogue finish the current one: an image pair of ~2 k descriptors is only ~10 tiles per
// task, and the per-task prologue used to cost as much as the tiles themselves.
template <bool FILTER>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_pass(const unsigned char* __restrict__ qbuf, const unsigned char* __restrict__ tbuf,
          const TcTask* __restrict__ tasks, const int* __restrict__ n_tasks_dev, int n_tasks_host,
          const unsigned* __restrict__ maxnorm_bits, TcTop2* __restrict__ res,
          const int* __restrict__ g_thr, int* __restrict__ cand_cnt, int* __restrict__ cand) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int task_end = n_tasks_dev ? *n_tasks_dev : n_tasks_host;
  if ((int)blockIdx.x >= task_end) return;   // uniform per CTA, before any barrier / TMEM use
  unsigned char* sA = tc_smem;                                        // 2 query blocks
  unsigned char* sB = tc_smem + 2 * TC_BLOCK_BYTES;                   // TC_STAGES x 2 blocks
  TcBarriers* bars = (TcBarriers*)(sB + (size_t)TC_STAGES * TC_TILE_BLOCKS * TC_BLOCK_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(smem_u32(&bars->full[s]), 1); mbar_init(smem_u32(&bars->empty[s]), 1); }
    for (int s = 0; s < 2; ++s) {
      for (int g = 0; g < TC_EPI_GROUPS; ++g) { mbar_init(smem_u32(&bars->acc_full[s][g]), 1); mbar_init(smem_u32(&bars->acc_empty[s][g]), 128); }
      mbar_init(smem_u32(&bars->a_full[s]), 1); mbar_init(smem_u32(&bars->a_empty[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ===== producer
    if (lane == 0) {
      uint32_t gt = 0;                       // target tiles issued so far (all tasks)
      uint32_t ti = 0;                       // tasks started so far
      for (int task = blockIdx.x; task < task_end; task += gridDim.x, ++ti) {
        const TcTask tk = tasks[task];
        const int ntile = tk.t_blocks / TC_TILE_BLOCKS;
        const uint32_t as = ti & 1u;
        mbar_wait(smem_u32(&bars->a_empty[as]), ((ti >> 1) & 1u) ^ 1u);
        mbar_expect_tx(smem_u32(&bars->a_full[as]), TC_BLOCK_BYTES);
        bulk_g2s(smem_u32(sA + (size_t)as * TC_BLOCK_BYTES), qbuf + (size_t)tk.q_blk * TC_BLOCK_BYTES, TC_BLOCK_BYTES,
                 smem_u32(&bars->a_full[as]));
        for (int t = 0; t < ntile; ++t, ++gt) {
          const uint32_t s = gt % TC_STAGES, ph = (gt / TC_STAGES) & 1u;
          mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1u);
          const uint32_t bytes = TC_TILE_BLOCKS * TC_BLOCK_BYTES;
          mbar_expect_tx(smem_u32(&bars->full[s]), bytes);
          bulk_g2s(smem_u32(sB + (size_t)s * bytes), tbuf + (size_t)(tk.t_blk0 + t * TC_TILE_BLOCKS) * TC_BLOCK_BYTES, bytes,
                   smem_u32(&bars->full[s]));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread)
    if (lane == 0) {
      // UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b F16 = 0, K-major both, N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);   // M 128, N 256
      uint32_t gt = 0, ti = 0;
      for (int task = blockIdx.x; task < task_end; task += gridDim.x, ++ti) {
        const TcTask tk = tasks[task];
        const int ntile = tk.t_blocks / TC_TILE_BLOCKS;
        const uint32_t asl = ti & 1u;
        mbar_wait(smem_u32(&bars->a_full[asl]), (ti >> 1) & 1u);
        const uint32_t a0 = smem_u32(sA + (size_t)asl * TC_BLOCK_BYTES);
        for (int t = 0; t < ntile; ++t, ++gt) {
          const uint32_t s = gt % TC_STAGES, as = gt & 1u;
          mbar_wait(smem_u32(&bars->full[s]), (gt / TC_STAGES) & 1u);
          const uint32_t b0 = smem_u32(sB + (size_t)s * TC_TILE_BLOCKS * TC_BLOCK_BYTES);
#pragma unroll
          for (int half = 0; half < TC_EPI_GROUPS; ++half) mbar_wait(smem_u32(&bars->acc_empty[as][half]), ((gt >> 1) & 1u) ^ 1u);
          tc_fence_after();
          // one M128 N256 K16 instruction per k-step covers the whole 256-row target tile: the query
          // block is read from shared memory once per tile and k-step instead of once per half
          const uint32_t d = tmem + (uint32_t)(as * 256);
#pragma unroll
          for (int k = 0; k < TC_KC / 2; ++k) {
            const uint64_t ad = make_smem_desc(a0 + k * 2 * TC_LBO, TC_LBO);
            const uint64_t bd = make_smem_desc(b0 + k * 2 * TC_TLBO, TC_TLBO);
            umma_f16(d, ad, bd, idesc, k > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int half = 0; half < TC_EPI_GROUPS; ++half) umma_commit(smem_u32(&bars->acc_full[as][half]));   // both column halves are ready
          umma_commit(smem_u32(&bars->empty[s]));      // smem slot reusable once these MMAs retire
        }
        umma_commit(smem_u32(&bars->a_empty[asl]));    // every MMA that reads this query block has retired
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread <-> query row (TMEM lane)
    const int row = (warp & 3) * 32 + lane;
    const int grp = (warp - 4) >> 2;           // which 128-column half of every tile
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(grp * 128);
    uint32_t gt = 0, ti = 0;
    for (int task = blockIdx.x; task < task_end; task += gridDim.x, ++ti) {
      const TcTask tk = tasks[task];
      const int ntile = tk.t_blocks / TC_TILE_BLOCKS;
      int g1 = 0x7f7fff00, g2 = 0x7f7fff00;   // running best / second (value bits, low 8 cleared)
      int gi = 0x7fffffff;
      const int grow = tk.q_row0 + row;        // FILTER: index of this gathered row
      const int col0 = (FILTER || n_tasks_dev) ? 0 : tk.t_pad;   // first pass: first column of this task's range
      uint32_t keymask;
      asm volatile("mov.b32 %0, 0xffffff00;" : "=r"(keymask));
      const int thr = FILTER ? g_thr[grow] : 0;
      // long target sets only: on short ones nearly every chunk still improves some row of the warp
      const bool prefilter = !FILTER && tk.t_blocks >= 128;
      for (int t = 0; t < ntile; ++t, ++gt) {
        const uint32_t as = gt & 1u;
        mbar_wait(smem_u32(&bars->acc_full[as][grp]), (gt >> 1) & 1u);
        tc_fence_after();
        int k1 = 0x7fffffff, k2 = 0x7fffffff;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_addr + (uint32_t)(as * 256 + c0), v);
          tmem_ld_wait();
          if (FILTER) {
            // branch-free hit mask first: a conditional body inside the 256-way unrolled compare
            // blows the loop up past the instruction cache (measured 3x slower per tile)
            unsigned hit = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) hit |= ((int)(v[j] & 0xffffff00u) <= thr) ? (1u << j) : 0u;
            while (hit) {
              const int j = __ffs(hit) - 1;
              hit &= hit - 1;
              const int col = t * 256 + grp * 128 + c0 + j;
              if (col < tk.t_n) {
                const int slot = atomicAdd(&cand_cnt[grow], 1);
                if (slot < TC_CAND_CAP) cand[(size_t)grow * TC_CAND_CAP + slot] = col;
              }
            }
          } else {
            if (prefilter) {
              // A chunk none of whose 32 scores is below the row's running second best cannot change
              // the row's top-2 (masking is monotone, ties keep the earlier column): one 3-input min
              // per two elements decides that, against 3.5 ALU instructions per element of the full
              // update.  With c chunks seen, a row improves in a chunk with probability ~2/c, a warp
              // of 32 rows with ~1 - exp(-64/c): after 16 k columns most chunks are skipped.
              int m = (int)v[0];
#pragma unroll
              for (int j = 1; j < 31; j += 2) m = min(m, min((int)v[j], (int)v[j + 1]));
              m = min(m, (int)v[31]);
              if (!__any_sync(0xffffffffu, m < g2)) continue;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              // (v & mask) | column in ONE LOP3: the mask has to sit in a register, because the
              // instruction takes a single immediate (the compiler's and-imm / or-imm pair costs two
              // ALU slots per element of a loop that is ALU-issue bound)
              uint32_t ukey;
              asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(ukey) : "r"(v[j]), "r"(keymask), "r"((uint32_t)(c0 + j)));
              const int key = (int)ukey;
              k2 = min(k2, max(k1, key));
              k1 = min(k1, key);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->acc_empty[as][grp]));
        if (!FILTER) {
          // merge the tile's top-2 into the running top-2
          const int v1 = k1 & (int)0xffffff00, v2 = k2 & (int)0xffffff00;
          if (v1 < g1) { g2 = min(g1, v2); g1 = v1; gi = col0 + t * 256 + grp * 128 + (k1 & 0xff); }
          else g2 = min(g2, v1);
        }
      }
      if (!FILTER) {
        // the two column halves meet here: group 1 hands its running top-2 to group 0.  One named
        // barrier per task is enough with two slots: group 1 cannot reach task ti+2 before group 0
        // has arrived at the barrier of task ti+1, i.e. after it read slot ti.
        int* mg = bars->merge[ti & 1u][row];
        if (grp == 1) { mg[0] = g1; mg[1] = g2; mg[2] = gi; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (grp == 1) continue;
        const int o1 = mg[0], o2 = mg[1], oi = mg[2];
        // equal scores keep the lower column, as the single-chain scan did
        if (o1 < g1 || (o1 == g1 && oi < gi)) { g2 = min(g1, o2); g1 = o1; gi = oi; }
        else g2 = min(g2, o1);
      }
      // device-planned tasks (n_tasks_dev) are GATHERED blocks: q_row0 is the block's first gathered
      // row, q_n its real rows, and the nomination goes to the gathered row's slot
      const int qrow = tk.q_row0 + row;
      if (!FILTER && (n_tasks_dev ? row < tk.q_n : qrow < tk.q_n)) {
        const float s = tc_scale_from_maxnorm(__uint_as_float(*maxnorm_bits));
        const float inv = 1.f / (s * s);
        TcTop2 o;
        o.m1 = (__int_as_float(g1) - 1.f) * inv;
        o.m2 = g2 == 0x7f7fff00 ? FLT_MAX : (__int_as_float(g2) - 1.f) * inv;
        o.idx = gi;
        o.pad = 0;
        res[(n_tasks_dev ? 0 : tk.res_off) + qrow] = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();                 // every role is done with the barriers, smem and TMEM
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ host side

size_t tc_block_bytes() { return TC_BLOCK_BYTES; }

size_t tc_smem_bytes() {
  return (size_t)TC_BLOCK_BYTES * (2 + TC_STAGES * TC_TILE_BLOCKS) + sizeof(TcBarriers) + 1024;
}

int tc_prepare(pano_ctx* ctx, const float* d_desc, const std::vector<TcImage>& imgs, TcOperands* ops) {
  const int n = (int)imgs.size();
  long long rows = 0, blocks = 0;
  int max_pad = 0;
  for (auto& im : imgs) { rows = std::max<long long>(rows, im.row0 + im.n); blocks = std::max<long long>(blocks, im.blk0 + im.n_pad / 128); max_pad = std::max(max_pad, im.n_pad); }
  int rc = 0;
  if ((rc = ctx_alloc(ctx, (void**)&ops->d_imgs, n * sizeof(TcImage))) ||
      (rc = ctx_alloc(ctx, (void**)&ops->d_norms, std::max<long long>(rows, 1) * sizeof(float))) ||
      (rc = ctx_alloc(ctx, (void**)&ops->d_maxnorm, sizeof(unsigned))) ||
      (rc = ctx_alloc(ctx, (void**)&ops->qbuf, (size_t)std::max<long long>(blocks, 1) * TC_BLOCK_BYTES)) ||
      (rc = ctx_alloc(ctx, (void**)&ops->tbuf, (size_t)std::max<long long>(blocks, 1) * TC_BLOCK_BYTES)))
    return rc;
  if ((rc = ctx_put(ctx, ops->d_imgs, imgs.data(), n * sizeof(TcImage)))) return rc;
  if ((rc = ctx_zero(ctx, ops->d_maxnorm, sizeof(unsigned)))) return rc;
  if (max_pad == 0) return PANO_OK;
  dim3 g1(ceil_div(max_pad, 128), n);
  PANO_LAUNCH(ctx, "k_tc_maxnorm", k_tc_maxnorm, g1, 128, 0, d_desc, ops->d_imgs, n, ops->d_norms, ops->d_maxnorm);
  PANO_LAUNCH(ctx, "k_tc_prep", k_tc_prep, g1, 128, 0, d_desc, ops->d_norms, ops->d_maxnorm, ops->d_imgs, ops->qbuf, ops->tbuf);
  return PANO_OK;
}

void tc_release(pano_ctx* ctx, TcOperands* ops) {
  ctx_free(ctx, ops->d_imgs); ctx_free(ctx, ops->d_norms); ctx_free(ctx, ops->d_maxnorm);
  ctx_free(ctx, ops->qbuf); ctx_free(ctx, ops->tbuf);
  *ops = TcOperands();
}

// function attributes are per DEVICE: one process may hold contexts on several GPUs
static int tc_set_attrs(pano_ctx* ctx) {
  if (ctx->attr_tc) return PANO_OK;
  const size_t smem = tc_smem_bytes();
  PANO_CUDA(ctx, cudaFuncSetAttribute(k_tc_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PANO_CUDA(ctx, cudaFuncSetAttribute(k_tc_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctx->attr_tc = true;
  return PANO_OK;
}

int tc_run_top2(pano_ctx* ctx, const TcOperands* ops, const TcTask* d_tasks, int n_tasks, TcTop2* d_res) {
  if (n_tasks == 0) return PANO_OK;
  const size_t smem = tc_smem_bytes();
  if (int rc = tc_set_attrs(ctx)) return rc;
  PANO_LAUNCH(ctx, "k_tc_top2", k_tc_pass<false>, std::min(n_tasks, ctx->num_sms), TC_THREADS, smem, ops->qbuf, ops->tbuf, d_tasks,
              (const int*)nullptr, n_tasks, ops->d_maxnorm, d_res, (const int*)nullptr, (int*)nullptr, (int*)nullptr);
  return PANO_OK;
}

// ------------------------------------------------------------------ gathered second pass

// Copies the query-form fp16 rows of the listed rows into gather blocks and sets
// each gathered row's threshold key / bookkeeping (approx == nullptr: a nomination pass over
// rows that have no first-pass result yet — no thresholds).  One CTA per gather block.
__global__ void __launch_bounds__(128)
k_tc_gather_rows(const unsigned char* __restrict__ qbuf, unsigned char* __restrict__ gq,
                 const TcTask* __restrict__ tasks, const int* __restrict__ n_tasks_dev,
                 const TcGatherSide* __restrict__ gsides, const int* __restrict__ list_rows,
                 const TcTop2* __restrict__ approx, const float* __restrict__ norms,
                 const unsigned* __restrict__ maxnorm_bits, int2* __restrict__ g_meta, int* __restrict__ g_thr,
                 int* __restrict__ cand_cnt) {
  if ((int)blockIdx.x >= *n_tasks_dev) return;
  const TcTask tk = tasks[blockIdx.x];
  const int side = (int)tk.res_off;                 // filter tasks carry the side index here
  const TcGatherSide gs = gsides[side];
  const int r = threadIdx.x, g = tk.q_row0 + r;
  unsigned char* dst = gq + (size_t)tk.q_blk * TC_BLOCK_BYTES + (r / 8) * TC_SBO + (r % 8) * 16;
  if (r < tk.q_n) {
    const int row = list_rows[gs.list_off + tk.t_pad + r];      // t_pad = first list slot of this block
    const unsigned char* src = qbuf + (size_t)(gs.q_blk0 + row / 128) * TC_BLOCK_BYTES + ((row % 128) / 8) * TC_SBO + (row % 8) * 16;
#pragma unroll
    for (int kc = 0; kc < TC_KC; ++kc) *(uint4*)(dst + (size_t)kc * TC_LBO) = *(const uint4*)(src + (size_t)kc * TC_LBO);
    if (approx) {
      const float nmax = __uint_as_float(*maxnorm_bits);
      const float s = tc_scale_from_maxnorm(nmax);
      const float nq = norms[gs.q_base + row];
      const float eps = 0.00215f * sqrtf(nq * nmax) + 0.0005f * nmax + 1.0f;    // == tc_eps in match.cu
      const TcTop2 ap = approx[gs.res_off + row];
      // every column whose exact distance can be the best or the second best scores <= m2~ + 2 eps
      float thr_v = ap.m2 == FLT_MAX ? FLT_MAX : (ap.m2 + 2.5f * eps) * (s * s) + 1.f;
      g_thr[g] = (int)(__float_as_uint(thr_v) | 0xffu);
      cand_cnt[g] = 0;
    }
    g_meta[g] = make_int2(side, row);
  } else {
#pragma unroll
    for (int kc = 0; kc < TC_KC; ++kc) *(uint4*)(dst + (size_t)kc * TC_LBO) = make_uint4(0, 0, 0, 0);
    if (approx) { g_thr[g] = -1; cand_cnt[g] = 0; }
    g_meta[g] = make_int2(-1, -1);
  }
}

int tc_run_filter(pano_ctx* ctx, const TcOperands* ops, const TcFilter* f, int max_blocks) {
  if (max_blocks <= 0) return PANO_OK;
  const size_t smem = tc_smem_bytes();
  if (int rc = tc_set_attrs(ctx)) return rc;
  PANO_LAUNCH(ctx, "k_tc_gather_rows", k_tc_gather_rows, max_blocks, 128, 0, ops->qbuf, f->gq, f->tasks, f->n_tasks,
              f->gsides, f->list_rows, f->approx, ops->d_norms, ops->d_maxnorm, f->g_meta, f->g_thr, f->cand_cnt);
  PANO_LAUNCH(ctx, "k_tc_filter", k_tc_pass<true>, std::min(max_blocks, ctx->num_sms), TC_THREADS, smem, f->gq, ops->tbuf, f->tasks, f->n_tasks,
              0, ops->d_maxnorm, (TcTop2*)nullptr, f->g_thr, f->cand_cnt, f->cand);
  return PANO_OK;
}

// Nomination (running top-2) over gathered rows: the columns-on-demand pass of match.cu.
// d_res[g] receives gathered row g's result.
int tc_run_nominate(pano_ctx* ctx, const TcOperands* ops, const TcFilter* f, int max_blocks, TcTop2* d_res) {
  if (max_blocks <= 0) return PANO_OK;
  const size_t smem = tc_smem_bytes();
  if (int rc = tc_set_attrs(ctx)) return rc;
  PANO_LAUNCH(ctx, "k_tc_gather_rows", k_tc_gather_rows, max_blocks, 128, 0, ops->qbuf, f->gq, f->tasks, f->n_tasks,
              f->gsides, f->list_rows, (const TcTop2*)nullptr, ops->d_norms, ops->d_maxnorm, f->g_meta, (int*)nullptr,
              (int*)nullptr);
  PANO_LAUNCH(ctx, "k_tc_nominate", k_tc_pass<false>, std::min(max_blocks, ctx->num_sms), TC_THREADS, smem, f->gq, ops->tbuf, f->tasks,
              f->n_tasks, 0, ops->d_maxnorm, d_res, (const int*)nullptr, (int*)nullptr, (int*)nullptr);
  return PANO_OK;
}
