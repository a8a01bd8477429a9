// Arguments and per-CTA partial layout shared by the FP32 backward (nais_bwd.cu) and the tensor-core backward
// (nais_pairs_tc_bwd.cu): both write the same workspace (dq rows, dp rows, parameter partials) for the same reduce kernels.
#pragma once
#include "nais_common.cuh"

namespace nais {

constexpr int BWD_MAXROWS = 16;

struct BwdArgs {
  NaisParams p;
  NaisPairs b;
  int bi;                 // branch
  const float* parts;     // [B] per-branch score of this branch
  const float* row_sum;   // [B]
  const float* dscore;    // [B]
  float* ws_dq;           // [B*H, D]
  float* ws_dp;           // [B, D]
  float* ws_part;         // [grid, part_stride]
  int part_stride;
  int rows_per_tile;
  int64_t n_items;        // work items (tiles of rows)
  int* bad;               // the library's bad-index word (nais_common.cuh)
};

// layout of one CTA's parameter partial: w1 [hid][D+lanes] | b1 [hid] | w2 [hid] | dist_w[4] dist_b[2] km[1] pad[1]
__host__ __device__ inline int part_floats(int hid, int D, int lanes) { return hid * (D + lanes) + 2 * hid + 8; }

}  // namespace nais
