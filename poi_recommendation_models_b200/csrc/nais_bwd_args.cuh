// Arguments and per-CTA partial layout shared by the FP32 backward (nais_bwd.cu) and the tensor-core backward
// (nais_pairs_tc_bwd.cu): both write the same workspace (dq rows, dp rows, parameter partials) for the same reduce kernels.
#pragma once
#include "nais_common.cuh"

namespace nais {

constexpr int BWD_MAXROWS = 16;  // == PAIR_MAXROWS (nais_pairs_tile.cuh)

struct BwdArgs {
  NaisParams p;
  NaisPairs b;
  int bi;                 // branch
  const float* parts;     // [B] per-branch score of this branch
  const float* row_sum;   // [B]
  const float* dscore;    // [B]
  const unsigned long long* act_mask;  // [B*H] ReLU pattern saved by the tcgen05 forward (hid <= 64), or NULL
  // Embedding-row contributions are written straight at their position in KEY order (the ids were sorted before this kernel
  // ran), so the segment reduce that follows streams contiguous rows instead of gathering 128-byte pieces through a permutation
  // (r1: the gather ran at 20 % of the HBM peak).  A NULL destination = that table needs no gradient: nothing is written.
  float* dq_h;            // [B*H, w_poi]      POI half of every cell's dq row, at pos_h[cell] (history-id order)
  float* dq_r;            // [B*H + B, w_reg]  region half of every cell's dq row at pos_r[cell], of every row's dp at pos_r[B*H + row]
  float* dp_t;            // [B, w_poi]        POI half of every row's dp, at pos_t[row] (target-id order)
  const uint32_t* pos_h;  // [B*H]
  const uint32_t* pos_r;  // [B*H + B]
  const uint32_t* pos_t;  // [B]
  float* ws_part;         // [grid, part_stride]
  int part_stride;
  int64_t n_items;        // work items (tiles of rows: nais_pairs_tile.cuh)
  int* bad;               // the library's bad-index word (nais_common.cuh)
};

// 4 consecutive elements d .. d+3 (d % 4 == 0) of cell `cell`'s dq row; ph / pr = pos_h[cell] / pos_r[cell]
__device__ __forceinline__ void store_dq4(const BwdArgs& A, int w_poi, int w_reg, uint32_t ph, uint32_t pr, int d, float4 v) {
  if ((w_poi & 3) == 0) {  // the group lies on one side of the POI | region boundary
    if (d < w_poi) {
      if (A.dq_h) *reinterpret_cast<float4*>(A.dq_h + (size_t)ph * w_poi + d) = v;
    } else if (A.dq_r) {
      *reinterpret_cast<float4*>(A.dq_r + (size_t)pr * w_reg + (d - w_poi)) = v;
    }
  } else {
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (d + i < w_poi) {
        if (A.dq_h) A.dq_h[(size_t)ph * w_poi + d + i] = e[i];
      } else if (A.dq_r) {
        A.dq_r[(size_t)pr * w_reg + (d + i - w_poi)] = e[i];
      }
    }
  }
}
// element d of row `row`'s dp; pt / pr = pos_t[row] / pos_r[B*H + row]
__device__ __forceinline__ void store_dp(const BwdArgs& A, int w_poi, int w_reg, uint32_t pt, uint32_t pr, int d, float v) {
  if (d < w_poi) {
    if (A.dp_t) A.dp_t[(size_t)pt * w_poi + d] = v;
  } else if (A.dq_r) {
    A.dq_r[(size_t)pr * w_reg + (d - w_poi)] = v;
  }
}

// layout of one CTA's parameter partial: w1 [hid][D+lanes] | b1 [hid] | w2 [hid] | dist_w[4] dist_b[2] km[1] pad[1]
__host__ __device__ inline int part_floats(int hid, int D, int lanes) { return hid * (D + lanes) + 2 * hid + 8; }

}  // namespace nais
