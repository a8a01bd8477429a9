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

// ---- software pipelining of a tile's inputs (the two-threads-per-cell tcgen05 pair kernels): Ampere-style cp.async into shared
// memory, issued one work unit ahead; a thread waits for ITS OWN copies with cp_async_wait_all(), other threads' after a barrier.
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// cell `cell` (0..127) of a work unit of those kernels (128 cells = chunk `ch` of a tile; a tile with H > 128 has several chunks,
// else one): which row of the tile / history position it is, and where its inputs live
struct PairCell {
  bool valid;
  int r, h;
  int64_t cidx, hidx;
};
__device__ __forceinline__ PairCell pair_cell(const PairTile& T, int ch, int cell) {
  PairCell c;
  if (T.H <= TC) {
    c.r = cell / T.H;
    c.h = cell - c.r * T.H;
    c.valid = c.r < T.nrows;
  } else {
    c.r = 0;
    c.h = ch * TC + cell;
    c.valid = c.h < T.H;
  }
  c.cidx = c.valid ? T.cell0 + c.r * (int64_t)T.H + c.h : 0;
  c.hidx = c.valid ? T.hist0 + c.r * T.hist_rs + c.h : 0;
  return c;
}
__device__ __forceinline__ int pair_chunks(const PairTile& T) { return T.H <= TC ? 1 : (T.H + TC - 1) / TC; }

}  // namespace nais
