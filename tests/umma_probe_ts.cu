// Stand-alone probe of the "A operand in tensor memory" MMAs the TMEM-A full-rank kernel relies on: A rows written with
// tcgen05.st (lane = row, 2 x fp16 / 4 x e5m2 per 32-bit column), B in shared memory (no-swizzle K-major), kind::f16 and
// kind::f8f6f4 adding into one fp32 accumulator; exact integer check against a CPU GEMM; then cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I poi_recommendation_models_b200/csrc -o tests/umma_probe_ts.bin tests/umma_probe_ts.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp8.h>

#include "umma.cuh"

using namespace nais::umma;

constexpr int M = 128, N = 144, K = 64;
constexpr int B16 = K / 8 * N * 16, B8 = K / 16 * N * 16;
constexpr int COL_A16 = 320, COL_A8 = 352, COL_D = 0, COL_T = 160;  // TMEM columns

__global__ void __launch_bounds__(192, 1) probe_kernel(const uint8_t* img, const uint32_t* a16, const uint32_t* a8, float* D, int reps,
                                                       long long* clk) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB16 = smem;
  uint8_t* sB8 = sB16 + B16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB8 + B8);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tslot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  if (warp == 5 && lane == 0) {
    mbar_expect_tx(&bars[0], B16 + B8);
    bulk_g2s(sB16, img, B16 + B8, &bars[0]);
  }
  if (warp < 4) {  // A rows -> TMEM: fp16 K = 64 -> 32 columns, e5m2 K = 64 -> 16 columns
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    uint32_t v[16];
    for (int h = 0; h < 2; ++h) {
      for (int i = 0; i < 16; ++i) v[i] = a16[row * 32 + h * 16 + i];
      tmem_st16(tmem + lane_addr + COL_A16 + h * 16, v);
    }
    for (int i = 0; i < 16; ++i) v[i] = a8[row * 16 + i];
    tmem_st16(tmem + lane_addr + COL_A8, v);
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t b0 = smem_u32(sB16), b8 = smem_u32(sB8);
    const uint32_t idh = idesc_f16(M, N), id8 = idesc_e5m2(M, N);
    auto step16 = [&](uint32_t d, int s, uint32_t acc) { mma_f16_ts(d, tmem + COL_A16 + s * 8, smem_desc(b0 + s * 2 * N * 16, N * 16, 128), idh, acc); };
    auto step8 = [&](uint32_t d, int s, uint32_t acc) { mma_f8_ts(d, tmem + COL_A8 + s * 8, smem_desc(b8 + s * 2 * N * 16, N * 16, 128), id8, acc); };
    for (int s = 0; s < K / 16; ++s) step16(tmem + COL_D, s, s > 0);
    for (int s = 0; s < K / 32; ++s) step8(tmem + COL_D, s, 1);
    mma_commit(&bars[1]);
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) step16(tmem + COL_T, r & 3, r > 0);
    mma_commit(&bars[2]);
    mbar_wait(&bars[2], 0);
    long long t1 = clock64();
    for (int r = 0; r < reps; ++r) step8(tmem + COL_T, r & 1, 1);
    mma_commit(&bars[3]);
    mbar_wait(&bars[3], 0);
    long long t2 = clock64();
    clk[0] = t1 - t0;
    clk[1] = t2 - t1;
  }
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + COL_D + c0, r);
      tmem_wait_ld();
      for (int i = 0; i < 16; ++i) D[row * N + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

int main() {
  std::vector<float> A(M * K), B(N * K), Af(M * K), Bf(N * K);
  srand(3);
  auto rnd = []() { return (float)((rand() % 9) - 4); };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  for (auto& v : Af) v = rnd() * 0.25f;
  for (auto& v : Bf) v = rnd() * 2.f;
  std::vector<uint8_t> img(B16 + B8);
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < K; ++k) {
      reinterpret_cast<__half*>(img.data())[((size_t)(k / 8) * N + r) * 8 + (k % 8)] = __float2half(B[r * K + k]);
      img[B16 + ((size_t)(k / 16) * N + r) * 16 + (k % 16)] = (uint8_t)__nv_cvt_float_to_fp8(Bf[r * K + k], __NV_SATFINITE, __NV_E5M2);
    }
  std::vector<uint32_t> a16(M * 32), a8(M * 16);
  for (int r = 0; r < M; ++r) {
    for (int c = 0; c < 32; ++c) {
      __half lo = __float2half(A[r * K + 2 * c]), hi = __float2half(A[r * K + 2 * c + 1]);
      a16[r * 32 + c] = (uint32_t)(*reinterpret_cast<uint16_t*>(&lo)) | ((uint32_t)(*reinterpret_cast<uint16_t*>(&hi)) << 16);
    }
    for (int c = 0; c < 16; ++c) {
      uint32_t w = 0;
      for (int e = 0; e < 4; ++e) w |= (uint32_t)(uint8_t)__nv_cvt_float_to_fp8(Af[r * K + 4 * c + e], __NV_SATFINITE, __NV_E5M2) << (8 * e);
      a8[r * 16 + c] = w;
    }
  }
  std::vector<float> ref(M * N);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[i * K + k] * B[j * K + k] + Af[i * K + k] * Bf[j * K + k];
      ref[i * N + j] = s;
    }
  uint8_t* dimg; uint32_t *da16, *da8; float* dD; long long* dclk;
  cudaMalloc(&dimg, img.size()); cudaMalloc(&da16, a16.size() * 4); cudaMalloc(&da8, a8.size() * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dclk, 32);
  cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(da16, a16.data(), a16.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(da8, a8.data(), a8.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);
  const int smem = B16 + B8 + 128;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 1800;
  probe_kernel<<<1, 192, smem>>>(dimg, da16, da8, dD, reps, dclk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("PROBE-TS CUDA ERROR: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> D(M * N);
  long long clk[4];
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(clk, dclk, 32, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      if (D[i * N + j] != ref[i * N + j]) {
        if (bad < 12) printf("mismatch D[%d][%d] = %g, expected %g\n", i, j, D[i * N + j], ref[i * N + j]);
        ++bad;
      }
  printf("PROBE-TS %s: %d mismatches of %d\n", bad ? "FAIL" : "OK", bad, M * N);
  printf("clk per MMA, A from TMEM (M=128,N=144): f16 K16 %.1f | e5m2 K32 %.1f\n", (double)clk[0] / reps, (double)clk[1] / reps);
  return bad ? 1 : 0;
}
