// Stand-alone probe of the tcgen05 building blocks (descriptors, no-swizzle K-major layout, bulk copy + mbarrier,
// TMEM alloc / ld mapping, N=144 and column-offset N=16 MMAs, zero-chunk aliasing through LBO) against a CPU GEMM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I poi_recommendation_models_b200/csrc -o /tmp/umma_probe tests/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma.cuh"

using namespace nais::umma;

constexpr int M = 128, N = 144, KX = 64, KC = KX / 8 + 1;  // 8 x-chunks + 1 ext chunk (second ext chunk aliases zeros)
constexpr int A_BYTES = KC * M * 16, B_BYTES = KC * N * 16, BLO_BYTES = KC * 16 * 16, ZERO_BYTES = N * 16;

__global__ void __launch_bounds__(192, 1) probe_kernel(const __half* Aimg, const __half* Bimg, const __half* Bloimg, float* D) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = sA + A_BYTES;
  uint8_t* sBlo = sB + B_BYTES;
  uint8_t* sZero = sBlo + BLO_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sZero + ZERO_BYTES);  // [0]=data full, [1]=mma done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < ZERO_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sZero)[i] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tslot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 5 && lane == 0) {  // "TMA" thread
    mbar_expect_tx(&bars[0], A_BYTES + B_BYTES + BLO_BYTES);
    bulk_g2s(sA, Aimg, A_BYTES, &bars[0]);
    bulk_g2s(sB, Bimg, B_BYTES, &bars[0]);
    bulk_g2s(sBlo, Bloimg, BLO_BYTES, &bars[0]);
  }
  if (warp == 4 && lane == 0) {  // MMA thread
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB), bl0 = smem_u32(sBlo), z0 = smem_u32(sZero);
    const uint32_t idN = idesc_f16(M, N), id16 = idesc_f16(M, 16);
    // main: 4 K-steps of 16 over the x part
    for (int s = 0; s < KX / 16; ++s)
      mma_f16(tmem, smem_desc(a0 + s * 2 * M * 16, M * 16, 128), smem_desc(b0 + s * 2 * N * 16, N * 16, 128), idN, s > 0);
    // ext step: first chunk = chunk 8, second chunk aliases the zero region through LBO
    {
      const uint32_t ae = a0 + 8 * M * 16, be = b0 + 8 * N * 16;
      mma_f16(tmem, smem_desc(ae, z0 - ae, 128), smem_desc(be, z0 - be, 128), idN, 1);
    }
    // aux: N=16 at column 128, B = lo image (16 rows), all 72 K columns
    for (int s = 0; s < KX / 16; ++s)
      mma_f16(tmem + 128, smem_desc(a0 + s * 2 * M * 16, M * 16, 128), smem_desc(bl0 + s * 2 * 16 * 16, 16 * 16, 128), id16, 1);
    {
      const uint32_t ae = a0 + 8 * M * 16, be = bl0 + 8 * 16 * 16;
      mma_f16(tmem + 128, smem_desc(ae, z0 - ae, 128), smem_desc(be, z0 - be, 128), id16, 1);
    }
    // aux2: N=16 at column 128 again, B = rows 128..143 of the hi image (descriptor offset inside the plane)
    for (int s = 0; s < KX / 16; ++s)
      mma_f16(tmem + 128, smem_desc(a0 + s * 2 * M * 16, M * 16, 128), smem_desc(b0 + s * 2 * N * 16 + 128 * 16, N * 16, 128), id16, 1);
    mma_commit(&bars[1]);
  }
  if (warp < 4) {  // epilogue: warp w owns TMEM lanes 32w..32w+31
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_wait_ld();
      for (int i = 0; i < 16; ++i) D[row * N + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

static void put(std::vector<__half>& img, int rows, int r, int k, float v) { img[((size_t)(k / 8) * rows + r) * 8 + (k % 8)] = __float2half(v); }

int main() {
  const int K = KX + 8;
  std::vector<float> A(M * K), B(N * K), Bl(16 * K);
  srand(1);
  auto rnd = []() { return (float)((rand() % 9) - 4); };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  for (auto& v : Bl) v = rnd();
  std::vector<__half> Ai((size_t)KC * M * 8), Bi((size_t)KC * N * 8), Bli((size_t)KC * 16 * 8);
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) put(Ai, M, r, k, A[r * K + k]);
  for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) put(Bi, N, r, k, B[r * K + k]);
  for (int r = 0; r < 16; ++r) for (int k = 0; k < K; ++k) put(Bli, 16, r, k, Bl[r * K + k]);
  std::vector<float> ref(M * N, 0.f);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[i * K + k] * B[j * K + k];
      if (j >= 128) {
        for (int k = 0; k < K; ++k) s += A[i * K + k] * Bl[(j - 128) * K + k];
        for (int k = 0; k < KX; ++k) s += A[i * K + k] * B[j * K + k];  // aux2: x part only
      }
      ref[i * N + j] = s;
    }
  __half *dA, *dB, *dBl;
  float* dD;
  cudaMalloc(&dA, Ai.size() * 2);
  cudaMalloc(&dB, Bi.size() * 2);
  cudaMalloc(&dBl, Bli.size() * 2);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, Ai.data(), Ai.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bi.data(), Bi.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBl, Bli.data(), Bli.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);
  const int smem = A_BYTES + B_BYTES + BLO_BYTES + ZERO_BYTES + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 192, smem>>>(dA, dB, dBl, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("PROBE CUDA ERROR: %s\n", cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> D(M * N);
  cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  int bad = 0, bad_main = 0, bad_aux = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j)
      if (D[i * N + j] != ref[i * N + j]) {
        if (bad < 12) printf("mismatch D[%d][%d] = %g, expected %g\n", i, j, D[i * N + j], ref[i * N + j]);
        ++bad;
        (j < 128 ? bad_main : bad_aux)++;
      }
  printf("PROBE %s: %d mismatches (main %d, aux %d) of %d\n", bad ? "FAIL" : "OK", bad, bad_main, bad_aux, M * N);
  return bad ? 1 : 0;
}
