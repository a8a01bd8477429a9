// Micro-benchmark: TMEM read bandwidth (tcgen05.ld 32x32b.x32 / .x16) as a function of the number of reading warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I poi_recommendation_models_b200/csrc -o tests/tmem_ld_timing.bin tests/tmem_ld_timing.cu
#include <cstdio>
#include "umma.cuh"
using namespace nais::umma;

template <int X>
__global__ void __launch_bounds__(512, 1) k(int reps, int inflight, long long* out, float* sink) {
  __shared__ uint32_t tslot;
  if (threadIdx.x < 32) tmem_alloc(&tslot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t a[32], b[32];
    if (X == 32) {
      tmem_ld32(base + ((r * 64) & 255), a);
      if (inflight == 2) tmem_ld32(base + ((r * 64 + 32) & 255), b);
    } else {
      tmem_ld16(base + ((r * 32) & 255), *reinterpret_cast<uint32_t(*)[16]>(&a[0]));
      if (inflight == 2) tmem_ld16(base + ((r * 32 + 16) & 255), *reinterpret_cast<uint32_t(*)[16]>(&b[0]));
    }
    tmem_wait_ld();
    for (int i = 0; i < X; ++i) acc += __uint_as_float(a[i]);
    if (inflight == 2) for (int i = 0; i < X; ++i) acc += __uint_as_float(b[i]);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
int main() {
  long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
  const int reps = 2000;
  for (int x : {32, 16}) for (int inflight : {1, 2}) for (int warps : {1, 4, 8, 16}) {
    if (x == 32) k<32><<<1, warps * 32>>>(reps, inflight, d, s); else k<16><<<1, warps * 32>>>(reps, inflight, d, s);
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    const double bytes = (double)reps * inflight * warps * 32 * x * 4;
    printf("x%d inflight=%d warps=%2d: %.1f clk/iter, %.1f B/clk per SM, %.1f clk per LDTM per warp\n", x, inflight, warps, (double)h / reps, bytes / h, (double)h / reps / inflight);
  }
  return 0;
}
