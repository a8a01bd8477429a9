// Stand-alone probe of the mixed-kind accumulation the NAIS_PREC_TC_MIX path relies on: kind::f16 MMAs (K = 16) and
// kind::f8f6f4 e5m2 MMAs (K = 32) adding into the SAME fp32 TMEM accumulator, no-swizzle K-major operands, against an
// exact integer CPU GEMM; then cycles per MMA for both kinds and for the interleaved issue pattern of one MIX step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I poi_recommendation_models_b200/csrc -o tests/umma_probe_f8.bin tests/umma_probe_f8.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp8.h>

#include "umma.cuh"

using namespace nais::umma;

constexpr int M = 128, N = 144, K = 64;
constexpr int A16 = K / 8 * M * 16, B16 = K / 8 * N * 16, A8 = K / 16 * M * 16, B8 = K / 16 * N * 16;

__global__ void __launch_bounds__(192, 1) probe_kernel(const uint8_t* img, float* D, int reps, long long* clk) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA16 = smem;
  uint8_t* sB16 = sA16 + A16;
  uint8_t* sA8 = sB16 + B16;
  uint8_t* sB8 = sA8 + A8;
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
    mbar_expect_tx(&bars[0], A16 + B16 + A8 + B8);
    bulk_g2s(sA16, img, A16 + B16 + A8 + B8, &bars[0]);
  }
  if (warp == 4 && lane == 0) {
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA16), b0 = smem_u32(sB16), a8 = smem_u32(sA8), b8 = smem_u32(sB8);
    const uint32_t idh = idesc_f16(M, N), id8 = idesc_e5m2(M, N);
    auto step16 = [&](uint32_t d, int s, uint32_t acc) {
      mma_f16(d, smem_desc(a0 + s * 2 * M * 16, M * 16, 128), smem_desc(b0 + s * 2 * N * 16, N * 16, 128), idh, acc);
    };
    auto step8 = [&](uint32_t d, int s, uint32_t acc) {
      mma_f8(d, smem_desc(a8 + s * 2 * M * 16, M * 16, 128), smem_desc(b8 + s * 2 * N * 16, N * 16, 128), id8, acc);
    };
    // correctness: columns [0,144) = fp16 GEMM + e5m2 GEMM (same accumulator) ; columns [160,304) = e5m2 GEMM alone
    for (int s = 0; s < K / 16; ++s) step16(tmem, s, s > 0);
    for (int s = 0; s < K / 32; ++s) step8(tmem, s, 1);
    for (int s = 0; s < K / 32; ++s) step8(tmem + 160, s, s > 0);
    mma_commit(&bars[1]);
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    // timing (results land in columns 320.. and are not checked)
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) step16(tmem + 320, r & 3, r > 0);
    mma_commit(&bars[2]);
    mbar_wait(&bars[2], 0);
    long long t1 = clock64();
    for (int r = 0; r < reps; ++r) step8(tmem + 320, r & 1, 1);
    mma_commit(&bars[3]);
    mbar_wait(&bars[3], 0);
    long long t2 = clock64();
    clk[0] = t1 - t0;
    clk[1] = t2 - t1;
    // one MIX step = 5 fp16 + 4 e5m2 ; one SPLIT step = 15 fp16
    long long t3 = clock64();
    for (int r = 0; r < reps / 9; ++r) {
      for (int s = 0; s < 5; ++s) step16(tmem + 320, s & 3, 1);
      for (int s = 0; s < 4; ++s) step8(tmem + 320, s & 1, 1);
    }
    mma_commit(&bars[2]);
    mbar_wait(&bars[2], 1);
    long long t4 = clock64();
    clk[2] = t4 - t3;
  }
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    const int row = warp * 32 + lane;
    for (int half = 0; half < 2; ++half)
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + half * 160 + c0, r);
        tmem_wait_ld();
        for (int i = 0; i < 16; ++i) D[(half * M + row) * N + c0 + i] = __uint_as_float(r[i]);
      }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

int main() {
  std::vector<float> A(M * K), B(N * K), Af(M * K), Bf(N * K);
  srand(2);
  auto rnd = []() { return (float)((rand() % 9) - 4); };   // |x| <= 4: exact in e5m2 (3 significant bits)
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  for (auto& v : Af) v = rnd() * 0.25f;                     // quarter steps: exercises fractional e5m2 values
  for (auto& v : Bf) v = rnd() * 2.f;
  std::vector<uint8_t> img(A16 + B16 + A8 + B8);
  auto put16 = [&](size_t base, int rows, int r, int k, float v) {
    reinterpret_cast<__half*>(img.data() + base)[((size_t)(k / 8) * rows + r) * 8 + (k % 8)] = __float2half(v);
  };
  auto put8 = [&](size_t base, int rows, int r, int k, float v) {
    img[base + ((size_t)(k / 16) * rows + r) * 16 + (k % 16)] = (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E5M2);
  };
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) { put16(0, M, r, k, A[r * K + k]); put8(A16 + B16, M, r, k, Af[r * K + k]); }
  for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) { put16(A16, N, r, k, B[r * K + k]); put8(A16 + B16 + A8, N, r, k, Bf[r * K + k]); }
  std::vector<float> ref(2 * M * N);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      float s16 = 0, s8 = 0;
      for (int k = 0; k < K; ++k) { s16 += A[i * K + k] * B[j * K + k]; s8 += Af[i * K + k] * Bf[j * K + k]; }
      ref[i * N + j] = s16 + s8;
      ref[(M + i) * N + j] = s8;
    }
  uint8_t* dimg; float* dD; long long* dclk;
  cudaMalloc(&dimg, img.size()); cudaMalloc(&dD, 2 * M * N * 4); cudaMalloc(&dclk, 32);
  cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 2 * M * N * 4);
  const int smem = A16 + B16 + A8 + B8 + 128;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 1800;
  probe_kernel<<<1, 192, smem>>>(dimg, dD, reps, dclk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("PROBE-F8 CUDA ERROR: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> D(2 * M * N);
  long long clk[4];
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(clk, dclk, 32, cudaMemcpyDeviceToHost);
  int bad = 0, bad_mixed = 0;
  for (int i = 0; i < 2 * M; ++i)
    for (int j = 0; j < N; ++j)
      if (D[i * N + j] != ref[i * N + j]) {
        if (bad < 12) printf("mismatch D[%d][%d] = %g, expected %g\n", i, j, D[i * N + j], ref[i * N + j]);
        ++bad;
        if (i < M) ++bad_mixed;
      }
  printf("PROBE-F8 %s: %d mismatches (mixed-kind accumulator %d, e5m2 alone %d) of %d\n", bad ? "FAIL" : "OK", bad, bad_mixed,
         bad - bad_mixed, 2 * M * N);
  printf("clk per MMA (M=128,N=144): f16 K16 %.1f | e5m2 K32 %.1f | MIX step (5 f16 + 4 e5m2) %.1f per step vs SPLIT 15 f16 = %.1f\n",
         (double)clk[0] / reps, (double)clk[1] / reps, (double)clk[2] / (reps / 9), 15.0 * clk[0] / reps);
  return bad ? 1 : 0;
}
