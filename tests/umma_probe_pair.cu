// Stand-alone probe of the CTA-pair (cta_group::2) building blocks the full-rank kernel's pair variant needs, against a CPU GEMM:
//   * tcgen05.alloc / mma / commit with cta_group::2: D[256 x 144] = A[256 x K] B[144 x K]^T, rows 0..127 of A and of D in CTA 0,
//     rows 128..255 in CTA 1, B rows 0..71 staged by CTA 0 and 72..143 by CTA 1 (each CTA holds N/2 rows of B),
//   * no-swizzle K-major operand images with LBO = 72 * 16 for the half B image,
//   * the peer's "operands have landed" relayed to the leader's mbarrier by a remote arrive (mapa + mbarrier.arrive.release.cluster),
//   * one K-step whose A operand is written with GENERIC stores by both CTAs' threads (fence.proxy.async + remote arrive on a
//     16-arrival barrier in the leader), like the A_ext lanes of the real kernel,
//   * an e5m2 (kind::f8f6f4) pass on top of the fp16 one,
//   * tcgen05.commit ... multicast::cluster to the same barrier offset in both CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I poi_recommendation_models_b200/csrc -o tests/umma_probe_pair.bin tests/umma_probe_pair.cu
#include <cuda_fp8.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma.cuh"

using namespace nais::umma;

constexpr int M = 128, N = 144, NH = N / 2, KX = 64, KC = KX / 8 + 2;  // 8 x-chunks + 2 "ext" chunks written by generic stores
constexpr int A_BYTES = KC * M * 16, B_BYTES = KC * NH * 16;
constexpr int A8_BYTES = (KX / 16) * M * 16, B8_BYTES = (KX / 16) * NH * 16;  // e5m2: 16 K elements per 16-byte chunk

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
    probe_kernel(const __half* Aimg, const __half* Bimg, const uint8_t* A8img, const uint8_t* B8img, const float* Aext, float* D) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = sA + A_BYTES;
  uint8_t* sA8 = sB + B_BYTES;
  uint8_t* sB8 = sA8 + A8_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB8 + B8_BYTES);  // [0] local data full, [1] mma done, [2] peer data full (leader), [3] ext written (leader, 8 arrivals)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 8);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc2(tslot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp == 5 && lane == 0) {  // producer: this CTA's A rows and its half of B
    mbar_expect_tx(&bars[0], (KX / 8) * M * 16 + (KX / 8) * NH * 16 + A8_BYTES + B8_BYTES);
    for (int c = 0; c < KX / 8; ++c) {  // (the ext chunks KX/8, KX/8+1 of A are written by generic stores; B's come with the image)
      bulk_g2s(sA + (size_t)c * M * 16, Aimg + ((size_t)rank * KC * M + (size_t)c * M) * 8, M * 16, &bars[0]);
      bulk_g2s(sB + (size_t)c * NH * 16, Bimg + ((size_t)rank * KC * NH + (size_t)c * NH) * 8, NH * 16, &bars[0]);
    }
    bulk_g2s(sA8, A8img + (size_t)rank * A8_BYTES, A8_BYTES, &bars[0]);
    bulk_g2s(sB8, B8img + (size_t)rank * B8_BYTES, B8_BYTES, &bars[0]);
  }
  if (warp == 5 && lane == 1) {  // the ext chunks of B (plain stores by one thread here: they are "constants")
    for (int i = 0; i < 2 * NH * 8; ++i)
      reinterpret_cast<__half*>(sB + (size_t)(KX / 8) * NH * 16)[i] = Bimg[((size_t)rank * KC * NH + (size_t)(KX / 8) * NH) * 8 + i];
    fence_proxy_async();
  }
  if (warp < 4) {  // epilogue warps write this CTA's rows of the two ext chunks of A with generic stores, then arrive on the LEADER's barrier
    const int row = warp * 32 + lane;
    __align__(16) __half v[16];
    for (int k = 0; k < 16; ++k) v[k] = __float2half(Aext[((size_t)rank * M + row) * 16 + k]);
    *reinterpret_cast<uint4*>(sA + ((size_t)(KX / 8) * M + row) * 16) = *reinterpret_cast<uint4*>(v);
    *reinterpret_cast<uint4*>(sA + ((size_t)(KX / 8 + 1) * M + row) * 16) = *reinterpret_cast<uint4*>(v + 8);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(&bars[3], 0);
  }
  __syncthreads();  // (lane 1 of warp 5 has written the B ext chunks: ordered before this CTA's relay / the leader's issue below)
  if (warp == 4) {
    if (rank == 1) {  // relay: my operands have landed -> the leader's barrier
      if (lane == 0) {
        mbar_wait(&bars[0], 0);
        mbar_arrive_cluster(&bars[2], 0);
      }
    } else if (lane == 0) {  // MMA thread of the leader
      mbar_wait(&bars[0], 0);
      mbar_wait_cluster(&bars[2], 0);
      mbar_wait_cluster(&bars[3], 0);
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB), a8 = smem_u32(sA8), b8 = smem_u32(sB8);
      const uint32_t idN = idesc_f16(2 * M, N), idN8 = idesc_e5m2(2 * M, N);
      for (int s = 0; s < KC / 2; ++s)  // 4 K-steps over x + 1 over the two ext chunks
        mma2_f16(tmem, smem_desc(a0 + s * 2 * M * 16, M * 16, 128), smem_desc(b0 + s * 2 * NH * 16, NH * 16, 128), idN, s > 0);
      for (int s = 0; s < KX / 32; ++s)  // e5m2 pass: K = 32 per instruction = two 16-byte chunks
        mma2_f8(tmem, smem_desc(a8 + s * 2 * M * 16, M * 16, 128), smem_desc(b8 + s * 2 * NH * 16, NH * 16, 128), idN8, 1);
      mma2_commit_multicast(&bars[1], 3);
    }
  }
  if (warp < 4) {  // epilogue of BOTH CTAs: warp w owns TMEM lanes 32w..32w+31 = rows rank*128 + 32w..
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_wait_ld();
      for (int i = 0; i < 16; ++i) D[((size_t)rank * M + row) * N + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync();
  if (warp == 4) tmem_dealloc2(tmem, 256);
}

static void put(std::vector<__half>& img, size_t base, int rows, int r, int k, float v) { img[base + ((size_t)(k / 8) * rows + r) * 8 + (k % 8)] = __float2half(v); }
static uint8_t e5m2(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E5M2); }

int main() {
  const int K = KX + 16;
  std::vector<float> A(2 * M * K), B(N * K), A8(2 * M * KX), B8(N * KX);
  srand(1);
  auto rnd = []() { return (float)((rand() % 9) - 4); };
  auto rnd8 = []() { return (float)((rand() % 5) - 2); };  // exactly representable in e5m2
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  for (auto& v : A8) v = rnd8();
  for (auto& v : B8) v = rnd8();
  // images: per CTA rank [k-chunk][rows][8 halves]; A: 128 rows of its M half; B: 72 rows of its N half
  std::vector<__half> Ai((size_t)2 * KC * M * 8), Bi((size_t)2 * KC * NH * 8);
  std::vector<float> Aext((size_t)2 * M * 16);
  for (int r = 0; r < 2 * M; ++r)
    for (int k = 0; k < K; ++k) {
      put(Ai, (size_t)(r / M) * KC * M * 8, M, r % M, k, A[r * K + k]);
      if (k >= KX) Aext[(size_t)r * 16 + (k - KX)] = A[r * K + k];
    }
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < K; ++k) put(Bi, (size_t)(r / NH) * KC * NH * 8, NH, r % NH, k, B[r * K + k]);
  std::vector<uint8_t> A8i((size_t)2 * A8_BYTES), B8i((size_t)2 * B8_BYTES);
  for (int r = 0; r < 2 * M; ++r)
    for (int k = 0; k < KX; ++k) A8i[(size_t)(r / M) * A8_BYTES + ((size_t)(k / 16) * M + r % M) * 16 + (k % 16)] = e5m2(A8[r * KX + k]);
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < KX; ++k) B8i[(size_t)(r / NH) * B8_BYTES + ((size_t)(k / 16) * NH + r % NH) * 16 + (k % 16)] = e5m2(B8[r * KX + k]);
  std::vector<float> ref((size_t)2 * M * N, 0.f);
  for (int i = 0; i < 2 * M; ++i)
    for (int j = 0; j < N; ++j) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[i * K + k] * B[j * K + k];
      for (int k = 0; k < KX; ++k) s += A8[i * KX + k] * B8[j * KX + k];
      ref[(size_t)i * N + j] = s;
    }
  __half *dA, *dB;
  uint8_t *dA8, *dB8;
  float *dD, *dAe;
  cudaMalloc(&dA, Ai.size() * 2);
  cudaMalloc(&dB, Bi.size() * 2);
  cudaMalloc(&dA8, A8i.size());
  cudaMalloc(&dB8, B8i.size());
  cudaMalloc(&dAe, Aext.size() * 4);
  cudaMalloc(&dD, (size_t)2 * M * N * 4);
  cudaMemcpy(dA, Ai.data(), Ai.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bi.data(), Bi.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dA8, A8i.data(), A8i.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB8, B8i.data(), B8i.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dAe, Aext.data(), Aext.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)2 * M * N * 4);
  const int smem = A_BYTES + B_BYTES + A8_BYTES + B8_BYTES + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<2, 192, smem>>>(dA, dB, dA8, dB8, dAe, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("PAIR PROBE CUDA ERROR: %s\n", cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> D((size_t)2 * M * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0, bad0 = 0, bad1 = 0;
  for (int i = 0; i < 2 * M; ++i)
    for (int j = 0; j < N; ++j)
      if (D[(size_t)i * N + j] != ref[(size_t)i * N + j]) {
        if (bad < 12) printf("mismatch D[%d][%d] = %g, expected %g\n", i, j, D[(size_t)i * N + j], ref[(size_t)i * N + j]);
        ++bad;
        (i < M ? bad0 : bad1)++;
      }
  printf("PAIR PROBE %s: %d mismatches (CTA 0 rows %d, CTA 1 rows %d) of %d\n", bad ? "FAIL" : "OK", bad, bad0, bad1, 2 * M * N);
  return bad ? 1 : 0;
}
