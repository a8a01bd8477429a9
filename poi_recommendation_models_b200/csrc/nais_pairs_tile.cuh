// Work tiles of the pair kernels: which rows / history cells a 128-cell tile covers, for both layouts of NaisPairs
// (include/nais_b200.h): dense [B, H] rows that each carry their own history, and the segmented (multi-user) layout where the
// rows of a segment share one history stored once.
#pragma once
#include "nais_common.cuh"

namespace nais {

constexpr int PAIR_MAXROWS = 16;  // rows sharing one 128-cell tile when the history is short

__host__ __device__ inline int pair_rows_per_tile(int H) {
  int r = H <= TC ? TC / (H > 0 ? H : 1) : 1;
  return r > PAIR_MAXROWS ? PAIR_MAXROWS : r;
}

struct PairTile {
  int64_t row0;     // first row of the tile (global row index)
  int nrows;        // rows in the tile (all of one segment)
  int H;            // their history length
  int64_t hist0;    // index of (row0, h = 0) in hist / hreg / hist_coords
  int64_t hist_rs;  // history-index stride between consecutive rows: H (dense) or 0 (segmented: shared history)
  int64_t cell0;    // index of (row0, h = 0) in the per-cell arrays (aux, act_mask, dq positions): cell (r, h) = cell0 + r*H + h
};

__host__ __device__ inline bool pairs_segmented(const NaisPairs& b) { return b.seg_offsets != nullptr; }
__host__ __device__ inline int64_t pairs_n_cells(const NaisPairs& b) { return pairs_segmented(b) ? b.n_cells : b.B * (int64_t)b.H; }
__host__ __device__ inline int64_t pairs_n_tiles(const NaisPairs& b) {
  if (pairs_segmented(b)) return b.n_tiles;
  const int rpt = pair_rows_per_tile(b.H);
  return (b.B + rpt - 1) / rpt;
}

__device__ __forceinline__ PairTile pair_tile(const NaisPairs& b, int64_t item) {
  PairTile t;
  if (!pairs_segmented(b)) {
    const int rpt = pair_rows_per_tile(b.H);
    t.row0 = item * rpt;
    const int64_t left = b.B - t.row0;
    t.nrows = (int)(left < rpt ? left : rpt);
    t.H = b.H;
    t.hist0 = t.row0 * (int64_t)b.H;
    t.hist_rs = b.H;
    t.cell0 = t.hist0;
  } else {
    const int s = __ldg(b.tile_seg + item);
    t.row0 = __ldg(b.tile_row0 + item);
    const int64_t h0 = __ldg(b.seg_offsets + s), r0 = __ldg(b.row_offsets + s), r1 = __ldg(b.row_offsets + s + 1);
    t.H = (int)(__ldg(b.seg_offsets + s + 1) - h0);
    const int rpt = pair_rows_per_tile(t.H);
    const int64_t left = r1 - t.row0;
    t.nrows = (int)(left < rpt ? left : rpt);
    t.hist0 = h0;
    t.hist_rs = 0;
    t.cell0 = __ldg(b.seg_cell_offsets + s) + (t.row0 - r0) * (int64_t)t.H;
  }
  return t;
}

// |dlat|, |dlon| (degrees) of one cell: the dense layout reads the tensor the reference gathers from latlon_mat (run.py:239-247),
// the segmented layout forms them from the centred fp32 coordinates like the full-rank path (exact to ~1e-8 degrees)
__device__ __forceinline__ void pair_latlon(const NaisPairs& b, int64_t cidx, int64_t hidx, int64_t row, float& l0, float& l1) {
  if (b.aux) {
    l0 = __ldg(b.aux + cidx * 2);
    l1 = __ldg(b.aux + cidx * 2 + 1);
  } else {
    l0 = fabsf(__ldg(b.tgt_coords + row * 2) - __ldg(b.hist_coords + hidx * 2));
    l1 = fabsf(__ldg(b.tgt_coords + row * 2 + 1) - __ldg(b.hist_coords + hidx * 2 + 1));
  }
}

}  // namespace nais
