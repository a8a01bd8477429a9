// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, fp16, no-swizzle K-major operands) for several N, issued
// back-to-back by one thread of one CTA.   ./umma_timing
#include <cstdio>
#include <vector>
#include "umma.cuh"
using namespace nais::umma;

__global__ void __launch_bounds__(128, 1) timing_kernel(int N, int reps, int alias, int kadv, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                 // 9 k-chunks x 128 rows x 16 B
  uint8_t* sB = sA + 9 * 128 * 16;    // 9 k-chunks x 256 rows x 16 B
  uint8_t* sZ = sB + 9 * 256 * 16;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sZ + 4096);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (9 * 128 * 16 + 9 * 256 * 16 + 4096) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(tslot, 512);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tslot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB), z0 = smem_u32(sZ);
    const uint32_t id = idesc_f16(128, N);
    const uint64_t da = alias ? smem_desc(a0, z0 - a0, 128) : smem_desc(a0, 128 * 16, 128);
    const uint64_t db = alias ? smem_desc(b0, z0 - b0, 128) : smem_desc(b0, N * 16, 128);
    const uint64_t as = kadv ? (2 * 128 * 16) >> 4 : 0, bs = kadv ? (2 * N * 16) >> 4 : 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) mma_f16(tmem, da + (r & 3) * as, db + (r & 3) * bs, id, r > 0);
    long long t1 = clock64();
    mma_commit(bar);
    mbar_wait(bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  const int smem = 9 * 128 * 16 + 9 * 256 * 16 + 4096 + 64;
  cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int Ns[] = {16, 64, 128, 144, 256};
  for (int alias = 0; alias < 2; ++alias) for (int kadv = 0; kadv < 2; ++kadv) for (int N : Ns) for (int reps : {15, 240}) {
    timing_kernel<<<1, 128, smem>>>(N, reps, alias, kadv, d);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    printf("alias=%d kadv=%d N=%3d reps=%3d issue=%6lld total=%7lld  per-mma=%.1f clk\n", alias, kadv, N, reps, h[0], h[1], (double)h[1] / reps);
  }
  return 0;
}
