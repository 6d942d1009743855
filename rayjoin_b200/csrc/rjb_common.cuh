// Shared declarations of the device side of librjb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "rjb200.h"

namespace rjb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define RJB_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess)                                                   \
      throw ::rjb::Error(RJB_ERR_CUDA, std::string(#call) + ": " +            \
                                           cudaGetErrorString(e__) + " at " + \
                                           __FILE__ + ":" +                   \
                                           std::to_string(__LINE__));         \
  } while (0)

#define RJB_REQUIRE(cond, msg)                          \
  do {                                                  \
    if (!(cond)) throw ::rjb::Error(RJB_ERR_INVALID, msg); \
  } while (0)

#define RJB_HD __host__ __device__ __forceinline__

constexpr int kNumSMs = 148;  // B200
constexpr int kQuantShift = 16;  // 47-bit coordinate -> 31-bit box coordinate

// Device view of one loaded map.  Points are interleaved int64 (x, y) so one
// 16-byte load fetches a vertex; an edge is (pts[eid + chain], pts[eid + chain + 1]).
struct MapView {
  const longlong2* pts;
  const uint32_t* edge_chain;  // per edge: eid -> chain (start point = eid + chain)
  const uint32_t* point_chain; // per point: p -> chain (eid of the edge starting at p = p - chain)
  const uint32_t* edge_desc;   // per point: occupancy descriptor of the edge starting there (edge_desc_of)
  const uint32_t* tile_desc;   // per kTileT points: cell box of the edges ENDING at them (tile_desc_of)
  const uint32_t* row_index;   // per chain, CSR into pts
  const uint32_t* last_bits;   // bit p set <=> point p is the last of its chain (no edge starts there)
  const int32_t* left;         // per chain
  const int32_t* right;        // per chain
  uint32_t n_points, n_edges, n_chains;
};

// BVH2 over leaves of <= leaf_size consecutive chain edges.  Internal node i
// keeps the boxes of BOTH children (quantised int32: coord >> kQuantShift) so
// one visit = 2 x 16 B + 8 B.  child < 0  ->  leaf ~child (sorted order).
struct BvhView {
  const int4* node_box;    // 2 per internal node: {xmin, ymin, xmax, ymax}
  const int2* node_child;  // {left, right}
  const uint2* leaf_rec;   // {first_eid, (count << 28) | chain}
  // 32-ary top tree: the binary nodes at depth 5, 10 and 15, addressed by their
  // root path (bit string), so that a warp resolves 5 levels per step with one
  // lane per slot.  Level k has 32^(k+1) slots; empty slots have an empty box.
  const int4* top_box;     // [32 | 1024 | 32768 | 1048576 (only when top_levels == 4)]
  const int* top_code;     // child code per slot (>= 0 internal node, < 0 ~leaf)
  // occupancy bitmap: kOccDim x kOccDim cells over the whole coordinate range, bit
  // set <=> some base edge's box touches the cell.  2 MB, cache resident: a query
  // edge whose box touches no occupied cell cannot intersect anything.  Followed by the
  // dilated bitmap occ2 (bit (x, y) = OR of occ over {x, x+1} x {y, y+1}), see edge_desc_of.
  const uint32_t* occ;
  // Cell directory over the OCCUPIED cells (sparse base maps only, else nullptr): occ_dir[word]
  // = {bitmap word, occupied cells before it}, cell id = rank + popcount of the lower bits; the
  // leaves whose box touches cell id are cell_item[cell_begin[id] .. cell_begin[id + 1]), each
  // item = the leaf's quantised box clipped to the cell + its first point and edge count
  // (cell_item_of).  A short query edge finds its candidate leaves with three dependent loads
  // and no tree walk (k_lsi_cells).
  const uint2* occ_dir;
  const uint32_t* cell_begin;
  const uint4* cell_item;
  int top_levels;          // 3, or 4 for big trees (>= 2^18 leaves): depth 20 resolved in 4 steps
  int4 root_box;
  uint32_t n_leaves;
};

constexpr int kOccBits = 12;             // 4096 x 4096 cells (8192^2 measured no better)
constexpr int kOccDim = 1 << kOccBits;
constexpr int kOccShift = 31 - kOccBits;  // quantised coordinates span 31 bits

// occupancy cell of a quantised coordinate (monotone, in [0, kOccDim))
static __host__ __device__ __forceinline__ int occ_cell(int q) {
  return (int) (((unsigned) q + (1u << 30)) >> kOccShift) & (kOccDim - 1);
}

// occupancy cell code of a vertex straight from its 47-bit coordinates:
// (v + 2^46) >> 35 equals occ_cell(quant(v)); packed as (cy << kOccBits | cx)
static __host__ __device__ __forceinline__ uint32_t occ_code(long long x, long long y) {
  const int sh = kQuantShift + kOccShift;
  const uint32_t cx = (uint32_t) ((unsigned long long) (x + (1ll << 46)) >> sh) & (kOccDim - 1);
  const uint32_t cy = (uint32_t) ((unsigned long long) (y + (1ll << 46)) >> sh) & (kOccDim - 1);
  return (cy << kOccBits) | cx;
}

// Occupancy descriptor of an edge, computed once when the map is loaded and streamed
// by k_lsi_filter (4 bytes per edge instead of two 16-byte vertices):
//   bits 0..23  code of the min corner cell of the edge's box
//   bits 24..25 class: kDescSame  both vertices in that cell -> test bitmap `occ`
//                      kDescSmall box within the 2 x 2 cells at the min corner -> test the
//                                 dilated bitmap `occ2` (bit = OR of those 2 x 2 cells)
//                      kDescBig   anything longer (rare) -> rectangle loop over `occ`
//                      kDescNone  no edge starts at this point (last point of a chain)
// Both bitmaps are one array: occ2 starts kOccWords words after occ.
constexpr uint32_t kDescSame = 0, kDescSmall = 1, kDescBig = 2, kDescNone = 3;
constexpr uint32_t kOccWords = (uint32_t) kOccDim * kOccDim / 32;

static __host__ __device__ __forceinline__ uint32_t edge_desc_of(uint32_t c1, uint32_t c2) {
  const uint32_t x1 = c1 & (kOccDim - 1), x2 = c2 & (kOccDim - 1), y1 = c1 >> kOccBits, y2 = c2 >> kOccBits;
  const uint32_t xm = x1 < x2 ? x1 : x2, ym = y1 < y2 ? y1 : y2;
  const uint32_t ex = (x1 < x2 ? x2 : x1) - xm, ey = (y1 < y2 ? y2 : y1) - ym;
  const uint32_t cls = (ex | ey) == 0 ? kDescSame : (ex <= 1 && ey <= 1) ? kDescSmall : kDescBig;
  return (cls << 24) | (ym << kOccBits) | xm;
}

// Tile descriptor: the union of the cell boxes of the kTileT edges that END at points
// kTileT t .. kTileT t + kTileT - 1 (= the edges STARTING one point earlier, whose descriptors are
// edge_desc[kTileT t - 1 .. kTileT t + kTileT - 2]).  k_lsi_filter_tiles decides a whole tile with
// ONE look-up in a bitmap dilated to the size of the box, and reads the edge descriptors only of
// the tiles that survive:
//   bits 0..23  code of the min corner cell of the box
//   bits 24..26 class: kTileNone no edge; 1..4 box within 1 / 2 x 2 / 4 x 4 / 8 x 8 cells at the
//               min corner -> bitmap (class - 1) of the four {occ, occ2, occ4, occ8}, where
//               occK(x, y) = OR of occ over [x, x + K) x [y, y + K);  kTileBig larger: kept
// Tile size (CPU experiment on the County x Zipcode-scale workload, tight boxes): 32 edges: 61 %
// of the tiles exceed 8 x 8 cells and 72 % stay live (measured: no gain); 16: 12 % / 35 %;
// 8: 5 % / 20 %.
constexpr int kTileT = 8;
constexpr uint32_t kTileNone = 0, kTileBig = 5;
constexpr int kOccMaps = 4;  // occ, occ2, occ4, occ8: one array, kOccWords words apart

static __host__ __device__ __forceinline__ uint32_t tile_desc_of(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) {
  const uint32_t e = (x1 - x0) > (y1 - y0) ? (x1 - x0) : (y1 - y0);
  const uint32_t cls = e == 0 ? 1u : e <= 1 ? 2u : e <= 3 ? 3u : e <= 7 ? 4u : kTileBig;
  return (cls << 24) | (y0 << kOccBits) | x0;
}

// Cell directory items.  Inside ONE cell the quantised coordinates of a box need kOccShift = 19
// bits each once they are clipped to the cell, so a leaf box, "starts left of / below the cell"
// flags, the leaf's edge count and its first point fit one 16-byte record:
//   x = x0 | x1[0..12] << 19      y = y0 | y1[0..12] << 19
//   z = x1[13..18] | y1[13..18] << 6 | starts_left << 12 | starts_below << 13 | (count - 1) << 14
//   w = first point of the leaf (eid + chain of its first edge)
// Two boxes that both touch cell (cx, cy) overlap iff their clipped boxes do (clipping is
// monotone, and where they do not overlap the separating coordinates lie inside the cell).
static __host__ __device__ __forceinline__ uint32_t cell_rel(int q, int c) {
  const long long r = (long long) ((unsigned) q + (1u << 30)) - ((long long) c << kOccShift);
  const long long hi = (1ll << kOccShift) - 1;
  return (uint32_t) (r < 0 ? 0 : (r > hi ? hi : r));
}

struct CellBoxQ {
  uint32_t x0, y0, x1, y1;
  bool left, below;  // the box starts left of / below the cell
};

static __host__ __device__ __forceinline__ CellBoxQ cell_clip(const int4& b, int cx, int cy) {
  CellBoxQ r;
  r.x0 = cell_rel(b.x, cx); r.y0 = cell_rel(b.y, cy);
  r.x1 = cell_rel(b.z, cx); r.y1 = cell_rel(b.w, cy);
  r.left = occ_cell(b.x) < cx;
  r.below = occ_cell(b.y) < cy;
  return r;
}

static __host__ __device__ __forceinline__ uint4 cell_item_of(const int4& b, int cx, int cy, uint32_t first_point,
                                                              uint32_t count) {
  const CellBoxQ r = cell_clip(b, cx, cy);
  return make_uint4(r.x0 | (r.x1 << 19), r.y0 | (r.y1 << 19),
                    (r.x1 >> 13) | ((r.y1 >> 13) << 6) | ((r.left ? 1u : 0u) << 12) | ((r.below ? 1u : 0u) << 13) |
                        ((count - 1) << 14),
                    first_point);
}

// item x query box (clipped to the same cell): boxes overlap, and this is the cell that reports
// the pair -- the one holding the min corner of the intersection of the two cell boxes, i.e. not
// both boxes start left of it and not both below it
static __host__ __device__ __forceinline__ bool cell_item_hit(const uint4& it, const CellBoxQ& q) {
  const uint32_t m19 = (1u << 19) - 1;
  const uint32_t ix0 = it.x & m19, iy0 = it.y & m19;
  const uint32_t ix1 = (it.x >> 19) | ((it.z & 63u) << 13), iy1 = (it.y >> 19) | (((it.z >> 6) & 63u) << 13);
  const bool il = (it.z >> 12) & 1u, ib = (it.z >> 13) & 1u;
  return ix0 <= q.x1 && q.x0 <= ix1 && iy0 <= q.y1 && q.y0 <= iy1 && !(il && q.left) && !(ib && q.below);
}

// (query, leaf) pairs handed to the exact pass come in two formats: {query start point, leaf id}
// from the tree walk, or DIRECT {query start point | (count - 1) << 29, first point of the leaf}
// from the cell directory (no leaf record to look up; maps of < 2^29 points)
constexpr uint32_t kDirectShift = 29;

constexpr int kTopOff0 = 0, kTopOff1 = 32, kTopOff2 = 32 + 1024, kTopOff3 = 32 + 1024 + 32768;
constexpr int kTopSlots3 = kTopOff3, kTopSlots4 = kTopOff3 + 32 * 32768;

static __device__ __forceinline__ int quant(long long v) {
  return (int) (v >> kQuantShift);  // arithmetic shift = floor, monotone
}

// A box that overlaps nothing: quantised coordinates are 31-bit, so no real box
// reaches INT_MAX / INT_MIN.  (An "inverted" box such as {1,1,0,0} is NOT safe:
// it passes the overlap test against any box that contains [0,1]^2.)
static __host__ __device__ __forceinline__ int4 empty_box() {
  return make_int4(0x7fffffff, 0x7fffffff, (int) 0x80000000, (int) 0x80000000);
}

static __device__ __forceinline__ bool box_overlap(const int4& a, const int4& b) {
  // closed boxes {xmin, ymin, xmax, ymax}
  return a.x <= b.z && b.x <= a.z && a.y <= b.w && b.y <= a.w;
}

// Query-wide counters that every warp adds to (candidate counts) are spread over kCtrSlots
// addresses, picked by the CTA: millions of atomics on ONE address serialise in the L2 (the
// 3.1 M warps of a 100 M-point PIP query each added their count to a single word).
constexpr int kCtrSlots = 64;

static inline unsigned div_up(uint64_t a, uint64_t b) {
  return (unsigned) ((a + b - 1) / b);
}

}  // namespace rjb
