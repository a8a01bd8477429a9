// Stand-alone probe for the planned tensor-core backward: can the SAME K-major SWIZZLE_NONE operand images
// that the forward GEMMs use be read MN-major (transpose bits of the instruction descriptor) for dW += dt^T X, where the
// contraction runs over the ROWS (cells) of both images?
//
//   P image: 128 rows x 64 cols, Q image: 128 rows x 80 cols, both [col-chunk of 8][row][8 x 16-bit] (umma.cuh), integer data.
//   Wanted:  C[m, n] = sum_r P[r, m] Q[r, n]      (M = 64 or 128 (P2: 128 cols), N = 80, K = 128 rows -> 8 MMAs of K = 16)
//
// In the MN-major canonical layout a core matrix is 8 (k) x 8 (mn) elements with the 8 mn elements contiguous (16 B) — which
// is what a K-major image IS when rows are read as k: offset(mn, k) = (mn%8)*2 + (k%8)*16 + (mn/8)*[chunk stride] + (k/8)*128.
// The probe tries both assignments of (chunk stride, 128) to (LBO, SBO), fp16 and bf16, M = 64 and M = 128, dumps the whole
// accumulator and reports which (descriptor variant, TMEM row mapping) reproduces C exactly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I poi_recommendation_models_b200/csrc -o tests/umma_probe_mn.bin tests/umma_probe_mn.cu
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma.cuh"

using namespace nais::umma;

constexpr int R = 128, NQ = 80;
constexpr int TMEM_COLS = 128;

__host__ __device__ constexpr uint32_t idesc_mn(int M, int N, int bf16) {
  // kind::f16: c_format F32 (bit 4), a/b format (bits 7, 10: 0 = f16, 1 = bf16), a_major (bit 15) = b_major (bit 16) = MN
  return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// variant 0: LBO = 128 B (next 8 rows = next k block), SBO = R*16 B (next col-chunk = next mn block); variant 1: swapped.
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint16_t* Pimg, int pcols, const uint16_t* Qimg, float* D, int M, int variant,
                                                       int bf16) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int p_bytes = (pcols / 8) * R * 16, q_bytes = (NQ / 8) * R * 16;
  uint8_t* sP = smem;
  uint8_t* sQ = sP + p_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sQ + q_bytes);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < p_bytes / 16; i += 128) reinterpret_cast<uint4*>(sP)[i] = reinterpret_cast<const uint4*>(Pimg)[i];
  for (int i = tid; i < q_bytes / 16; i += 128) reinterpret_cast<uint4*>(sQ)[i] = reinterpret_cast<const uint4*>(Qimg)[i];
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tslot, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  // zero the accumulator first so untouched lanes / columns read back as 0
  {
    uint32_t z[16];
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    for (int c0 = 0; c0 < TMEM_COLS; c0 += 16) tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + c0, z);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      tc_fence_after();
      const uint32_t p0 = smem_u32(sP), q0 = smem_u32(sQ);
      const uint32_t chunk = R * 16, rows8 = 128;
      const uint32_t lbo = variant == 0 ? rows8 : chunk, sbo = variant == 0 ? chunk : rows8;
      const uint32_t idesc = idesc_mn(M, NQ, bf16);
      for (int s = 0; s < R / 16; ++s)  // K-step s = rows 16s .. 16s+15 of both images = 2 blocks of 8 rows, 256 B further on
        mma_f16(tmem, smem_desc(p0 + s * 256, lbo, sbo), smem_desc(q0 + s * 256, lbo, sbo), idesc, s > 0);
      mma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < NQ; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_wait_ld16(r);
    for (int i = 0; i < 16; ++i) D[tid * NQ + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

static uint16_t enc(float v, int bf16) {
  if (bf16) {
    __nv_bfloat16 h = __float2bfloat16(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

int main() {
  srand(3);
  int all_ok = 1;
  for (int bf16 = 0; bf16 < 2; ++bf16)
    for (int M : {64, 128}) {
      const int pcols = M;
      std::vector<float> P(R * pcols), Q(R * NQ);
      for (auto& v : P) v = (float)((rand() % 7) - 3);
      for (auto& v : Q) v = (float)((rand() % 7) - 3);
      std::vector<uint16_t> Pi((size_t)(pcols / 8) * R * 8), Qi((size_t)(NQ / 8) * R * 8);
      for (int r = 0; r < R; ++r) {
        for (int c = 0; c < pcols; ++c) Pi[((size_t)(c / 8) * R + r) * 8 + c % 8] = enc(P[r * pcols + c], bf16);
        for (int c = 0; c < NQ; ++c) Qi[((size_t)(c / 8) * R + r) * 8 + c % 8] = enc(Q[r * NQ + c], bf16);
      }
      std::vector<float> C((size_t)M * NQ, 0.f);
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < NQ; ++n) {
          float s = 0;
          for (int r = 0; r < R; ++r) s += P[r * pcols + m] * Q[r * NQ + n];
          C[(size_t)m * NQ + n] = s;
        }
      uint16_t *dP, *dQ;
      float* dD;
      cudaMalloc(&dP, Pi.size() * 2);
      cudaMalloc(&dQ, Qi.size() * 2);
      cudaMalloc(&dD, 128 * NQ * 4);
      cudaMemcpy(dP, Pi.data(), Pi.size() * 2, cudaMemcpyHostToDevice);
      cudaMemcpy(dQ, Qi.data(), Qi.size() * 2, cudaMemcpyHostToDevice);
      const int smem = (pcols / 8) * R * 16 + (NQ / 8) * R * 16 + 64;
      cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      int found = 0;
      for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dD, 0, 128 * NQ * 4);
        probe_kernel<<<1, 128, smem>>>(dP, pcols, dQ, dD, M, variant, bf16);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("PROBE_MN CUDA ERROR (bf16=%d M=%d variant=%d): %s\n", bf16, M, variant, cudaGetErrorString(e));
          return 2;
        }
        std::vector<float> D(128 * NQ);
        cudaMemcpy(D.data(), dD, 128 * NQ * 4, cudaMemcpyDeviceToHost);
        // TMEM row mappings to try: identity (lane = m) and the 16-rows-per-subpartition layout (lane = 32*(m/16) + m%16)
        for (int map = 0; map < (M == 64 ? 2 : 1); ++map) {
          int bad = 0;
          for (int m = 0; m < M; ++m)
            for (int n = 0; n < NQ; ++n) {
              const int lane = map == 0 ? m : 32 * (m / 16) + m % 16;
              bad += D[lane * NQ + n] != C[(size_t)m * NQ + n];
            }
          printf("  %s M=%d variant=%d (LBO=%s) map=%s: %d mismatches of %d\n", bf16 ? "bf16" : "fp16", M, variant,
                 variant == 0 ? "128B rows8, SBO=chunk" : "chunk, SBO=128B rows8", map == 0 ? "lane=m" : "lane=32*(m/16)+m%16", bad, M * NQ);
          if (!bad) found = 1;
        }
        if (M == 64 && variant == 0) {  // where did row 17 of C land? (helps if neither mapping matched)
          for (int lane = 0; lane < 128; ++lane) {
            int eq = 1;
            for (int n = 0; n < NQ; ++n) eq &= D[lane * NQ + n] == C[(size_t)17 * NQ + n];
            if (eq) printf("    C row 17 found in TMEM lane %d\n", lane);
          }
        }
      }
      printf("PROBE_MN %s M=%d: %s\n", bf16 ? "bf16" : "fp16", M, found ? "OK (a variant reproduces C exactly)" : "FAIL");
      all_ok &= found;
      cudaFree(dP);
      cudaFree(dQ);
      cudaFree(dD);
    }
  printf("PROBE_MN %s\n", all_ok ? "ALL OK" : "SOME FAILED");
  return all_ok ? 0 : 1;
}
