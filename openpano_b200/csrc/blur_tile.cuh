// blur_tile.cuh — the separable Gaussian of the reference (feature/gaussian.hh:29-90: column
// pass first, then the row pass over its result; tmp = 0; tmp += v[k] * tap[k], k ascending,
// separate multiply and add) on one 64x32 tile staged in shared memory, plus the TMA /
// mbarrier helpers that stage the tile.  Shared by the SIFT pyramid (sift.cu) and the
// multi-band blender (blend.cu).
#pragma once
#include "common.cuh"

#define BT_W 64
#define BT_H 32
#define BT_THREADS 256

struct BlurTile { int om; int tx, ty; };

// CTA-level tile index -> (plane entry, tile x, tile y): binary search over the ascending first-tile
// index of the n_om plane entries (the table is a few KB and L1-resident).
// span[k] = (first tile of plane entry k, tiles per tile row)
__device__ __forceinline__ BlurTile find_blur_tile(const int2* __restrict__ span, int n_om, int cta) {
  int lo = 0, hi = n_om - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (span[mid].x <= cta) lo = mid; else hi = mid - 1;
  }
  const int local = cta - span[lo].x, tx_n = span[lo].y;
  BlurTile t;
  t.om = lo; t.ty = local / tx_n; t.tx = local - t.ty * tx_n;
  return t;
}

// Fast path for the window half-widths the reference's defaults produce (kw = 7
// and 13): both passes keep a sliding window in registers so one
// shared-memory load feeds up to 2C+1 taps, 8 outputs per thread per pass.
//   column pass: lane <-> column, thread = 8 consecutive rows
//   row pass   : lane <-> row (odd row stride => conflict-free), thread = 8
//                consecutive columns, results staged transposed-safe in `outT`
//   store      : lane <-> column, coalesced level + |DoG| writes
// The arithmetic per output is unchanged: tmp = 0; tmp += v[k] * tap[k], k ascending.
template <int C>
__device__ __forceinline__ void blur_level(const float* __restrict__ grey, float* __restrict__ colbuf,
                                           float* __restrict__ outT, const float* __restrict__ taps_g,
                                           int R, int RX, int GW, int CS, int tid) {
  constexpr int KW = 2 * C + 1;
  float tap[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) tap[k] = taps_g[k];
  const int cw = BT_W + 2 * C;
  // column pass: items = (BT_H/8 row strips) x cw columns
  for (int item = tid; item < (BT_H / 8) * cw; item += BT_THREADS) {
    const int strip = item / cw, xx = item - strip * cw;
    const float* col = grey + (strip * 8 + R - C) * GW + (xx + RX - C);
    float win[8 + 2 * C];
#pragma unroll
    for (int j = 0; j < 8 + 2 * C; ++j) win[j] = col[j * GW];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float tmp = 0.f;
#pragma unroll
      for (int k = 0; k < KW; ++k) tmp += win[r + k] * tap[k];
      colbuf[(strip * 8 + r) * CS + xx] = tmp;
    }
  }
  __syncthreads();
  // row pass: warp <-> 8-column strip, lane <-> row
  {
    const int lane = tid & 31, xs = (tid >> 5) * 8;
    const float* row = colbuf + lane * CS + xs;
    float win[8 + 2 * C];
#pragma unroll
    for (int j = 0; j < 8 + 2 * C; ++j) win[j] = row[j];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float tmp = 0.f;
#pragma unroll
      for (int k = 0; k < KW; ++k) tmp += win[r + k] * tap[k];
      outT[lane * (BT_W + 1) + xs + r] = tmp;
    }
  }
  __syncthreads();
}

// ---- TMA / mbarrier helpers (tile loads of the blur kernel)
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbar_wait(uint32_t bar, uint32_t parity) {   // bounded: a protocol bug traps
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
// 2-D tiled TMA load (cp.async.bulk.tensor): box origin (cx, cy) may lie outside the tensor,
// out-of-range elements arrive as zeros.  The descriptor lives in global memory (written by
// an earlier kernel of this stream), hence the tensormap-proxy acquire fence.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const TmaDesc* map, int cx, int cy, uint32_t bar) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(map) : "memory");
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(cx), "r"(cy), "r"(bar) : "memory");
}

