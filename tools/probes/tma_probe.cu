// Probe: which forms of a 2-D TMA tile load work on this box?  One variant per process
// (faults are sticky).  nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>   // 0: grid constant, 1: global no fence, 2: global + fence
__global__ void k(const __grid_constant__ CUtensorMap pm, const CUtensorMap* gm, float* out, int bw, int bh, int cx, int cy) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const void* m = MODE == 0 ? (const void*)&pm : (const void*)gm;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(bw * bh * 4) : "memory");
    if (MODE == 2) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(m) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(su32(sm)), "l"(m), "r"(cx), "r"(cy), "r"(su32(&bar)) : "memory");
  }
  unsigned done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; ++spin)
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(su32(&bar)) : "memory");
  if (!done) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0, bw = argc > 2 ? atoi(argv[2]) : 76, bh = argc > 3 ? atoi(argv[3]) : 44;
  int cx = argc > 4 ? atoi(argv[4]) : 64, cy = argc > 5 ? atoi(argv[5]) : 32;
  const int W = 918, H = 681, P = 928;
  float* d; cudaMalloc(&d, P * H * 4);
  std::vector<float> h(P * H);
  for (int i = 0; i < P * H; ++i) h[i] = (float)(i % P) + 1000.f * (i / P);
  cudaMemcpy(d, h.data(), P * H * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry %d %d %p\n", (int)e, (int)q, fn);
  alignas(64) CUtensorMap tm;
  cuuint64_t gd[2] = {W, H}, gs[1] = {P * 4};
  cuuint32_t bx[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
  CUresult r = ((enc_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  CUtensorMap* gm; cudaMalloc(&gm, 128); cudaMemcpy(gm, &tm, 128, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, bw * bh * 4);
  size_t smem = bw * bh * 4;
  if (mode == 0) k<0><<<1, 128, smem>>>(tm, gm, out, bw, bh, cx, cy);
  if (mode == 1) k<1><<<1, 128, smem>>>(tm, gm, out, bw, bh, cx, cy);
  if (mode == 2) k<2><<<1, 128, smem>>>(tm, gm, out, bw, bh, cx, cy);
  e = cudaDeviceSynchronize();
  printf("mode %d box %dx%d at (%d,%d): sync -> %s\n", mode, bw, bh, cx, cy, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> o(bw * bh);
    cudaMemcpy(o.data(), out, bw * bh * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
      int gx = cx + x, gy = cy + y;
      float want = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0.f : (float)gx + 1000.f * gy;
      if (o[y * bw + x] != want) { if (bad < 3) printf("  mismatch (%d,%d): %g vs %g\n", x, y, o[y * bw + x], want); ++bad; }
    }
    printf("  %d mismatches, o[0]=%g\n", bad, o[0]);
  }
  return 0;
}
