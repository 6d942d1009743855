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

constexpr int kNumSMs = 148;  // B200
constexpr int kQuantShift = 16;  // 47-bit coordinate -> 31-bit box coordinate

// Device view of one loaded map.  Points are interleaved int64 (x, y) so one
// 16-byte load fetches a vertex; an edge is (pts[eid + chain], pts[eid + chain + 1]).
struct MapView {
  const longlong2* pts;
  const uint32_t* edge_chain;  // per edge
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
  // edge whose box touches no occupied cell cannot intersect anything.
  const uint32_t* occ;
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

static inline unsigned div_up(uint64_t a, uint64_t b) {
  return (unsigned) ((a + b - 1) / b);
}

}  // namespace rjb
