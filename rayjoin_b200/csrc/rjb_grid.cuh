// Uniform-grid variant (RayJoin's -mode=grid).
//
// Replaces UniformGrid::AddMapToGrid (reference: src/grid/uniform_grid.h:131-358: a dense
// 12-byte Cell per grid cell, filled by two rounds of global atomics, 6 GB at 15000^2),
// LSIGrid::Query (src/app/lsi_grid.h:19-78,97-159: one thread per CELL doing ne0 x ne1 tests
// -> 1.3 s of load imbalance) and PIPGrid::Query (src/app/pip_grid.h:21-77 +
// src/algo/pip.h:14-115).
//
// Index: SPARSE and built by sorting.  Every base edge is registered in every cell of the
// cell-box of its end points (iterate_cell, uniform_grid.h:44-86) with the reference's own cell
// function (cell.h:15-22, factor 0.999 and all); the (cell, edge) incidences are radix-sorted
// by cell with the engine's onesweep and become a CSR over the OCCUPIED cells only.  A cell is
// found through a bitmap (one bit per cell) and the rank of its word: id = rank[word] +
// popcount(lower bits).  Cells are numbered column-major (bit = cx * stride + cy) so that the
// PIP walk "next occupied cell above" is a scan of consecutive bits.  Items are the START
// POINT index of the edge (eid + chain): the two vertices load without the eid -> chain ->
// point chain.  Size: g^2 / 4 bytes (bitmap + ranks) + 4 B per occupied cell + 4 B per incidence.
//
// LSI semantics are the reference's grid semantics, not the LBVH's: the predicate is always
// evaluated as intersect_test(map-0 edge, map-1 edge) whatever the query map is
// (lsi_grid.h:62, "fixme: respect query map id" at :96), and a pair is kept only in the cell
// of its intersection point (:64-67), i.e. iff that cell lies in the cell-boxes of both edges.
#pragma once
#include "rjb_exact.cuh"
#include "rjb_lsi.cuh"
#include "rjb_prims.cuh"

namespace rjb {

// The cells of the INDEX are not the reference's square cells: the same number of them
// (-grid_size squared) is laid out as 4 x as many columns of a quarter of the rows.  PIP walks a
// column upward, and in a narrow column nearly every edge met spans the whole column, so the
// first occupied cell above the point almost always holds the answer (square cells of one edge
// length: 16 occupied cells and 53 screened edges per point on the 28 M-edge map, most of them
// edges that run alongside the column without covering px).  The index is only a filter; the
// reference's square cells enter where they define the RESULT: the ownership rule of the grid
// LSI (grid_owned) is evaluated arithmetically with ref_scale, whatever cells the index has.
struct GridView {
  const uint32_t* bits;        // occupancy, column-major: bit (cx * gs + cy)
  const uint32_t* rank;        // occupied cells before each bitmap word
  const uint32_t* cell_begin;  // [n_occ + 1] CSR into items
  const uint2* items;          // per incidence: {start point (eid + chain) of the base edge,
                               //  its box relative to the cell in 1/128 steps (grid_qbox)}
  uint32_t gx, gy, gs;         // columns, rows; column stride in bits (gy rounded up to 32)
  uint32_t gsize;              // the caller's -grid_size (reference cells: gsize x gsize)
  long long imin;
  double sx, sy;               // index cells: (double) gx / internal_range * 0.999, likewise gy
  double inv_sy;
  double ref_scale;            // (double) gsize / internal_range * 0.999   (cell.h:19)
};

struct Grid {
  DBuf<uint32_t> bits, pop, rank, cell_begin, cnt, off, big;
  DBuf<uint2> items;
  DBuf<uint64_t> keys_a, keys_b;
  DBuf<unsigned long long> totals;
  ScanTemp scan_tmp;
  SortTemp sort_tmp;
  uint32_t gsize = 0, gx = 0, gy = 0, gs = 0, n_occ = 0;
  uint64_t n_items = 0;
  long long imin = 0;
  double sx = 0, sy = 0, ref_scale = 0;
  bool built = false;
  uint32_t n_words() const { return gx * (gs / 32); }
  GridView view() const {
    GridView v;
    v.bits = bits.p;
    v.rank = rank.p;
    v.cell_begin = cell_begin.p;
    v.items = items.p;
    v.gsize = gsize;
    v.gx = gx;
    v.gy = gy;
    v.gs = gs;
    v.imin = imin;
    v.sx = sx;
    v.sy = sy;
    v.inv_sy = sy > 0 ? 1.0 / sy : 0;
    v.ref_scale = ref_scale;
    return v;
  }
  size_t index_bytes() const {
    return 2 * (size_t) (n_words() + 1) * sizeof(uint32_t) + ((size_t) n_occ + 1) * sizeof(uint32_t) +
           n_items * sizeof(uint2);
  }
};

// cell of a coordinate for FILTERING: the reference's function clamped into the grid (query
// points may lie outside the bounding box of the maps); monotone
static __device__ __forceinline__ int grid_cx(const GridView& g, long long v) {
  return min(max(ref_cell(v, g.imin, g.sx), 0), (int) g.gx - 1);
}
static __device__ __forceinline__ int grid_cy(const GridView& g, long long v) {
  return min(max(ref_cell(v, g.imin, g.sy), 0), (int) g.gy - 1);
}

// Position of a coordinate inside cell c of its axis in 1/128 steps, clamped to [0, 127]:
// monotone in v (every floating-point step is), so v1 <= v2 implies grid_q(v1) <= grid_q(v2) and
// comparisons of these numbers are CONSERVATIVE stand-ins for comparisons of the coordinates.
constexpr int kGridQ = 128;   // steps per cell row (y)
constexpr int kGridQX = 64;   // steps per cell column (x)

static __device__ __forceinline__ uint32_t grid_q(long long v, long long imin, double scale, int c, int steps) {
  const double t = ((double) (v - imin) * scale - (double) c) * (double) steps;
  return (uint32_t) min(max((int) floor(t), 0), steps - 1);
}
static __device__ __forceinline__ uint32_t grid_qx(const GridView& g, long long v, int c) {
  return grid_q(v, g.imin, g.sx, c, kGridQX);
}
static __device__ __forceinline__ uint32_t grid_qy(const GridView& g, long long v, int c) {
  return grid_q(v, g.imin, g.sy, c, kGridQ);
}

// Box of an edge as seen from cell (cx, cy), 32 bits:
//   bits  0.. 5  qx0   x range inside the cell's column in 1/64 steps (clamped)
//   bits  6..11  qx1
//   bit  12      the edge starts LEFT of the column  (so xmin < px for every px in the column)
//   bit  13      the edge ends RIGHT of the column
//   bits 14..20  qy0   ymin inside the cell (0 when the edge starts below it)
//   bits 21..27  qy1   ymax inside ITS cell, which is cell cy + dy
//   bits 28..31  dy    rows between this cell and the cell of ymax (15 = that many or more)
// The PIP walk screens and bounds an edge with this word alone; the vertices are loaded
// afterwards, and only for the edges that pass.
static __device__ __forceinline__ uint32_t grid_qbox(const GridView& g, const longlong2& a, const longlong2& b,
                                                     int cx, int cy) {
  const long long ymax = max(a.y, b.y);
  const int cy1 = grid_cy(g, ymax);
  const uint32_t dy = (uint32_t) min(max(cy1 - cy, 0), 15);
  const long long xmin = min(a.x, b.x), xmax = max(a.x, b.x);
  const uint32_t xl = grid_cx(g, xmin) < cx ? 1u : 0u, xr = grid_cx(g, xmax) > cx ? 1u : 0u;
  return grid_qx(g, xmin, cx) | (grid_qx(g, xmax, cx) << 6) | (xl << 12) | (xr << 13) |
         (grid_qy(g, min(a.y, b.y), cy) << 14) | (grid_qy(g, ymax, cy1) << 21) | (dy << 28);
}

struct CellBox {
  int x0, x1, y0, y1;
  __device__ __forceinline__ unsigned long long cells() const {
    return (unsigned long long) (x1 - x0 + 1) * (unsigned long long) (y1 - y0 + 1);
  }
};

static __device__ __forceinline__ CellBox edge_cell_box(const GridView& g, const longlong2& a, const longlong2& b) {
  CellBox c;
  c.x0 = grid_cx(g, min(a.x, b.x));
  c.x1 = grid_cx(g, max(a.x, b.x));
  c.y0 = grid_cy(g, min(a.y, b.y));
  c.y1 = grid_cy(g, max(a.y, b.y));
  return c;
}

// occupied cell -> [begin, end) of its items
static __device__ __forceinline__ bool grid_cell_range(const GridView& g, uint32_t bit, uint32_t& beg,
                                                       uint32_t& end) {
  const uint32_t w = __ldg(&g.bits[bit >> 5]);
  if (!((w >> (bit & 31)) & 1u)) return false;
  const uint32_t id = __ldg(&g.rank[bit >> 5]) + __popc(w & ((1u << (bit & 31)) - 1));
  beg = __ldg(&g.cell_begin[id]);
  end = __ldg(&g.cell_begin[id + 1]);
  return true;
}

// ---- build ------------------------------------------------------------------------------
// An edge that covers more than kGridBigCells cells is not walked by its thread: it goes to a
// list that a CTA per edge works off (a map-spanning edge is g^2 incidences).
constexpr uint32_t kGridBigCells = 256;

// per START POINT: number of cells of the edge's cell-box (0 for the last point of a chain)
__global__ void __launch_bounds__(256)
k_grid_count(MapView B, GridView g, uint32_t* __restrict__ cnt, unsigned long long* __restrict__ totals) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long n = 0;
  if (p < B.n_points) {
    const bool last = (__ldg(&B.last_bits[p >> 5]) >> (p & 31)) & 1u;
    if (!last) n = edge_cell_box(g, B.pts[p], B.pts[p + 1]).cells();
    cnt[p] = (uint32_t) min(n, 0xFFFFFFFFull);
  }
  unsigned long long s = n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(&totals[0], s);
  if (n > kGridBigCells) atomicAdd(&totals[1], 1ull);
}

__global__ void __launch_bounds__(256)
k_grid_emit(MapView B, GridView g, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off,
            uint64_t* __restrict__ key, uint32_t* __restrict__ big_list,
            unsigned long long* __restrict__ totals) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B.n_points) return;
  const uint32_t n = cnt[p];
  if (n == 0) return;
  if (n > kGridBigCells) {
    big_list[atomicAdd(&totals[2], 1ull)] = p;
    return;
  }
  const CellBox c = edge_cell_box(g, B.pts[p], B.pts[p + 1]);
  uint32_t o = off[p];
  for (int x = c.x0; x <= c.x1; x++)
    for (int y = c.y0; y <= c.y1; y++) {
      key[o] = (((uint64_t) x * g.gs + y) << 32) | p;  // packed (cell, start point)
      o++;
    }
}

__global__ void __launch_bounds__(256)
k_grid_emit_big(MapView B, GridView g, const uint32_t* __restrict__ off, const uint32_t* __restrict__ big_list,
                const unsigned long long* __restrict__ totals, uint64_t* __restrict__ key) {
  const uint32_t n_big = (uint32_t) totals[2];
  for (uint32_t i = blockIdx.x; i < n_big; i += gridDim.x) {
    const uint32_t p = big_list[i];
    const CellBox c = edge_cell_box(g, B.pts[p], B.pts[p + 1]);
    const uint32_t ny = c.y1 - c.y0 + 1;
    const uint32_t total = (uint32_t) c.cells();
    const uint32_t o = off[p];
    for (uint32_t t = threadIdx.x; t < total; t += blockDim.x) {
      key[o + t] = (((uint64_t) (c.x0 + t / ny) * g.gs + (c.y0 + t % ny)) << 32) | p;
    }
  }
}

// sorted keys -> occupancy bits (the first item of every run sets its cell's bit)
__global__ void k_grid_mark(const uint64_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ bits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = (uint32_t) (key[i] >> 32);
  if (i == 0 || (uint32_t) (key[i - 1] >> 32) != k) atomicOr(&bits[k >> 5], 1u << (k & 31));
}

__global__ void k_grid_popc(const uint32_t* __restrict__ bits, uint32_t n_words, uint32_t* __restrict__ pop) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < n_words) pop[w] = __popc(bits[w]);
}

__global__ void k_grid_cell_begin(const uint64_t* __restrict__ key, uint32_t n, const uint32_t* __restrict__ bits,
                                  const uint32_t* __restrict__ rank, uint32_t n_words,
                                  uint32_t* __restrict__ cell_begin) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = (uint32_t) (key[i] >> 32);
  if (i == 0 || (uint32_t) (key[i - 1] >> 32) != k) {
    const uint32_t id = rank[k >> 5] + __popc(bits[k >> 5] & ((1u << (k & 31)) - 1));
    cell_begin[id] = i;
  }
  if (i == n - 1) cell_begin[rank[n_words]] = n;  // rank[n_words] = number of occupied cells
}

__global__ void k_grid_items(MapView B, GridView g, const uint64_t* __restrict__ key, uint32_t n,
                             uint2* __restrict__ items) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = (uint32_t) (key[i] >> 32);
  const uint32_t p = (uint32_t) key[i];
  items[i] = make_uint2(p, grid_qbox(g, B.pts[p], B.pts[p + 1], (int) (k / g.gs), (int) (k % g.gs)));
}

static inline void build_grid(Grid& g, const MapView& B, uint32_t gsize, long long imin,
                              long long irange, cudaStream_t st) {
  RJB_REQUIRE(gsize >= 1 && gsize <= 32768, "grid_size must be in 1..32768");
  g.built = false;  // until the last call below has succeeded
  g.gsize = gsize;
  // the same number of cells as the reference's gsize x gsize, 4 x the columns, 1/4 of the rows
  g.gx = gsize >= 4 ? std::min(gsize * 4u, 65536u) : gsize;
  g.gy = std::max(1u, (uint32_t) (((uint64_t) gsize * gsize + g.gx - 1) / g.gx));
  g.gs = (g.gy + 31u) & ~31u;
  g.imin = imin;
  g.ref_scale = (double) gsize / (double) irange * 0.999;  // src/grid/cell.h:19
  g.sx = (double) g.gx / (double) irange * 0.999;
  g.sy = (double) g.gy / (double) irange * 0.999;
  g.n_items = 0;
  g.n_occ = 0;
  const uint32_t n_words = g.n_words();
  uint32_t* bits = g.bits.ensure(n_words + 1);
  uint32_t* pop = g.pop.ensure(n_words + 1);
  uint32_t* rank = g.rank.ensure(n_words + 1);
  RJB_CUDA(cudaMemsetAsync(bits, 0, (n_words + 1) * sizeof(uint32_t), st));
  unsigned long long* totals = g.totals.ensure(3);  // incidences, big edges, big-list cursor
  RJB_CUDA(cudaMemsetAsync(totals, 0, 3 * sizeof(unsigned long long), st));
  uint32_t* cnt = g.cnt.ensure(B.n_points + 1);
  uint32_t* off = g.off.ensure(B.n_points + 1);
  GridView v = g.view();
  unsigned long long h_tot[2] = {0, 0};
  if (B.n_points) {
    k_grid_count<<<div_up(B.n_points, 256), 256, 0, st>>>(B, v, cnt, totals);
    exclusive_scan_u32(cnt, off, B.n_points, g.scan_tmp, st);
    RJB_CUDA(cudaMemcpyAsync(h_tot, totals, sizeof(h_tot), cudaMemcpyDeviceToHost, st));
    RJB_CUDA(cudaStreamSynchronize(st));
  }
  // (the 32-bit scan above is only valid below 2^32; the 64-bit total says whether it is)
  RJB_REQUIRE(h_tot[0] < (1ull << 30), "grid: more than 2^30 edge-cell incidences; lower -grid_size");
  const uint32_t n = (uint32_t) h_tot[0];
  g.n_items = n;
  uint32_t* cbeg = g.cell_begin.ensure((size_t) std::min<uint64_t>(n, (uint64_t) n_words * 32) + 2);
  uint2* items = g.items.ensure(n ? n : 1);
  if (n) {
    uint64_t* ka = g.keys_a.ensure(n);
    g.keys_b.ensure(n);
    uint32_t* big = g.big.ensure(h_tot[1] ? h_tot[1] : 1);
    k_grid_emit<<<div_up(B.n_points, 256), 256, 0, st>>>(B, v, cnt, off, ka, big, totals);
    if (h_tot[1])
      k_grid_emit_big<<<(unsigned) std::min<uint64_t>(h_tot[1], 4 * kNumSMs), 256, 0, st>>>(B, v, off, big, totals, ka);
    int key_bits = 1;
    while (key_bits < 32 && (((uint64_t) g.gx * g.gs) >> key_bits)) key_bits++;
    const uint64_t* kb = sort_packed(ka, g.keys_b.p, n, 0, key_bits, g.sort_tmp, st);
    k_grid_items<<<div_up(n, 256), 256, 0, st>>>(B, g.view(), kb, n, items);
    k_grid_mark<<<div_up(n, 256), 256, 0, st>>>(kb, n, bits);
    k_grid_popc<<<div_up(n_words, 256), 256, 0, st>>>(bits, n_words, pop);
    exclusive_scan_u32(pop, rank, n_words, g.scan_tmp, st);
    k_grid_cell_begin<<<div_up(n, 256), 256, 0, st>>>(kb, n, bits, rank, n_words, cbeg);
    RJB_CUDA(cudaGetLastError());
    RJB_CUDA(cudaMemcpyAsync(&g.n_occ, rank + n_words, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RJB_CUDA(cudaStreamSynchronize(st));
  } else {
    RJB_CUDA(cudaMemsetAsync(rank, 0, (n_words + 1) * sizeof(uint32_t), st));
    RJB_CUDA(cudaMemsetAsync(cbeg, 0, 2 * sizeof(uint32_t), st));
    RJB_CUDA(cudaStreamSynchronize(st));
  }
  g.built = true;
}

// ---- LSI ----------------------------------------------------------------------------------
// Pass 1 streams the query map (lane = start point, coalesced vertex loads) and emits one
// work item (start point, cell) per OCCUPIED cell of the edge's cell-box: most query edges
// fall into empty cells and end after one to four bitmap look-ups.  One atomic per warp.
// Pass 2 resolves the work items densely (exact integer boxes, every pair in exactly one of
// its common cells, intersect_test with every lane busy), pass 3 is the point pass of the
// LBVH path.
constexpr uint32_t kGridSmallCells = 16;

__global__ void __launch_bounds__(256)
k_grid_lsi_filter(MapView Q, const uint32_t* __restrict__ order, uint32_t p_lo, uint32_t p_hi, GridView g,
                  uint2* __restrict__ work, uint32_t work_cap, unsigned int* work_n,
                  uint32_t* __restrict__ big_list, unsigned int* big_n) {
  const int lane = threadIdx.x & 31;
  // slots: the start points of the query window [p_lo, p_hi) in map order, or (order != null, no
  // window) the p_hi = n_edges entries of the map's Morton-ordered edge list
  const uint32_t tile = (order ? 0u : p_lo / 32) + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (tile * 32 >= p_hi) return;
  const QTile t = load_tile(Q, order, p_hi, tile, lane, 32, order ? 0u : p_lo);
  CellBox c = {0, -1, 0, -1};
  uint32_t hits = 0;  // bit k = k-th cell of the box (column by column) is occupied
  bool big = false;
  if (t.valid) {
    c = edge_cell_box(g, t.a, t.b);
    big = c.cells() > kGridSmallCells;
    if (!big) {
      const int ny = c.y1 - c.y0 + 1;
      const int n = (int) c.cells();
#pragma unroll 4
      for (int k = 0; k < n; k++) {
        const uint32_t bit = (uint32_t) (c.x0 + k / ny) * g.gs + (uint32_t) (c.y0 + k % ny);
        if ((__ldg(&g.bits[bit >> 5]) >> (bit & 31)) & 1u) hits |= 1u << k;
      }
    }
  }
  // long edges: a list of their own, resolved by a CTA each (k_grid_lsi_big)
  const unsigned mb = __ballot_sync(0xffffffffu, big);
  if (mb) {
    unsigned base = 0;
    const int leader = __ffs(mb) - 1;
    if (lane == leader) base = atomicAdd(big_n, (unsigned) __popc(mb));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (big) big_list[base + __popc(mb & ((1u << lane) - 1))] = t.p;
  }
  const uint32_t cnt = __popc(hits);
  const uint32_t inc = warp_incl_scan(cnt, lane);
  const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
  if (total == 0) return;
  unsigned base = 0;
  if (lane == 31) base = atomicAdd(work_n, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  uint32_t pos = base + inc - cnt;
  const int ny = c.y1 - c.y0 + 1;
  while (hits) {
    const int k = __ffs(hits) - 1;
    hits &= hits - 1;
    if (pos < work_cap) work[pos] = make_uint2(t.p, (uint32_t) (c.x0 + k / ny) * g.gs + (uint32_t) (c.y0 + k % ny));
    pos++;
  }
}

__global__ void __launch_bounds__(256)
k_grid_lsi_big(MapView Q, GridView g, const uint32_t* __restrict__ big_list, const unsigned int* __restrict__ big_n,
               uint2* __restrict__ work, uint32_t work_cap, unsigned int* work_n) {
  const int lane = threadIdx.x & 31;
  const uint32_t n_big = *big_n;
  for (uint32_t i = blockIdx.x; i < n_big; i += gridDim.x) {
    const uint32_t p = big_list[i];
    const CellBox c = edge_cell_box(g, Q.pts[p], Q.pts[p + 1]);
    const uint32_t ny = c.y1 - c.y0 + 1;
    const uint32_t total = (uint32_t) min(c.cells(), 0xFFFFFFFFull);
    // block-uniform trip count (warp collectives inside)
    for (uint32_t t0 = 0; t0 < total; t0 += blockDim.x) {
      const uint32_t t = t0 + threadIdx.x;
      uint32_t bit = 0;
      bool hit = false;
      if (t < total) {
        bit = (uint32_t) (c.x0 + t / ny) * g.gs + (c.y0 + t % ny);
        hit = (__ldg(&g.bits[bit >> 5]) >> (bit & 31)) & 1u;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m == 0) continue;
      unsigned base = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(work_n, (unsigned) __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      const unsigned pos = base + __popc(m & ((1u << lane) - 1));
      if (hit && pos < work_cap) work[pos] = make_uint2(p, bit);
    }
  }
}

// the pair (query edge, base edge) is examined in ONE of the cells both are registered in:
// the min corner of the intersection of their cell-boxes
static __device__ __forceinline__ bool grid_pair_here(const GridView& g, const Seg& q, const Seg& b, uint32_t bit) {
  const int cx = max(grid_cx(g, min(q.x1, q.x2)), grid_cx(g, min(b.x1, b.x2)));
  const int cy = max(grid_cy(g, min(q.y1, q.y2)), grid_cy(g, min(b.y1, b.y2)));
  return (uint32_t) cx * g.gs + (uint32_t) cy == bit;
}

// reference rule (lsi_grid.h:64-67): kept iff the cell of the intersection point is a cell both
// edges are registered in.  e0 = map-0 edge, e1 = map-1 edge.
static __device__ __forceinline__ bool grid_owned(const GridView& g, const Seg& e0, const Seg& e1) {
  // the REFERENCE's square cells (gsize x gsize), not the cells of the index
  const int cx = lsi_xsect_ref_cell(e0, e1, 0, g.imin, g.ref_scale);
  const int cy = lsi_xsect_ref_cell(e0, e1, 1, g.imin, g.ref_scale);
  auto lo = [&](long long a, long long b) { return ref_cell(min(a, b), g.imin, g.ref_scale); };
  auto hi = [&](long long a, long long b) { return ref_cell(max(a, b), g.imin, g.ref_scale); };
  return cx >= max(lo(e0.x1, e0.x2), lo(e1.x1, e1.x2)) && cx <= min(hi(e0.x1, e0.x2), hi(e1.x1, e1.x2)) &&
         cy >= max(lo(e0.y1, e0.y2), lo(e1.y1, e1.y2)) && cy <= min(hi(e0.y1, e0.y2), hi(e1.y1, e1.y2));
}

// intersect_test(map-0 edge, map-1 edge) + the ownership rule for the lanes with `have`, hits
// appended to the result queue as start-point index pairs (one atomic per warp)
static __device__ __forceinline__ void grid_test_emit(const MapView& Q, const MapView& B, const GridView& g, int q,
                                                      bool have, uint2 it, rjb_xsect* __restrict__ out,
                                                      uint32_t cap, unsigned int* counter, int lane) {
  bool found = false;
  if (have) {
    const longlong2 a = __ldg(&Q.pts[it.x]), b = __ldg(&Q.pts[it.x + 1]);
    const longlong2 c = __ldg(&B.pts[it.y]), d = __ldg(&B.pts[it.y + 1]);
    const Seg eq = {a.x, a.y, b.x, b.y}, eb = {c.x, c.y, d.x, d.y};
    // always (map 0, map 1), whatever the query side is
    const Seg& e0 = q == 0 ? eq : eb;
    const Seg& e1 = q == 0 ? eb : eq;
    found = lsi_intersect(e0, e1) && grid_owned(g, e0, e1);
  }
  const unsigned m = __ballot_sync(0xffffffffu, found);
  if (m == 0) return;
  unsigned base = 0;
  const int leader = __ffs(m) - 1;
  if (lane == leader) base = atomicAdd(counter, (unsigned) __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (found) {
    const unsigned pos = base + __popc(m & ((1u << lane) - 1));
    if (pos < cap) {
      out[pos].eid[0] = it.x;  // point indices for now; the point pass turns them into eids
      out[pos].eid[1] = it.y;
    }
  }
}

static __device__ __forceinline__ void grid_drain(const MapView& Q, const MapView& B, const GridView& g, int q,
                                                  const uint2* list, unsigned n_list, rjb_xsect* __restrict__ out,
                                                  uint32_t cap, unsigned int* counter) {
  const int lane = threadIdx.x & 31;
  for (unsigned t0 = threadIdx.x - lane; t0 < n_list; t0 += kExactThreads) {
    const unsigned t = t0 + lane;
    grid_test_emit(Q, B, g, q, t < n_list, t < n_list ? list[t] : make_uint2(0, 0), out, cap, counter, lane);
  }
}

constexpr int kGridInLane = 8;  // items of a cell a lane walks itself; longer lists: the whole warp

// box of a query edge relative to cell `bit`, in the steps of the item records: {qx0, qx1, qy0, qy1}
static __device__ __forceinline__ uint4 grid_query_qbox(const GridView& g, const Seg& e, uint32_t bit) {
  const int cx = (int) (bit / g.gs), cy = (int) (bit % g.gs);
  return make_uint4(grid_qx(g, min(e.x1, e.x2), cx), grid_qx(g, max(e.x1, e.x2), cx),
                    grid_qy(g, min(e.y1, e.y2), cy), grid_qy(g, max(e.y1, e.y2), cy));
}

// Conservative box test on the item's packed 32-bit box alone: both boxes are clamped into the
// cell by the same monotone map, so boxes that overlap still overlap.  In dense cells most items
// end here, without a vertex load or a 64-bit compare.
static __device__ __forceinline__ bool grid_qbox_overlap(const uint4& q, uint32_t m) {
  const uint32_t ix0 = m & 63u, ix1 = (m >> 6) & 63u, iy0 = (m >> 14) & 127u;
  const uint32_t iy1 = (m >> 28) ? 127u : ((m >> 21) & 127u);  // ends in a higher cell: up to the top
  return ix0 <= q.y && q.x <= ix1 && iy0 <= q.w && q.z <= iy1;
}

__global__ void __launch_bounds__(kExactThreads)
k_grid_lsi_exact(MapView Q, MapView B, GridView g, int q, const uint2* __restrict__ work,
                 const unsigned int* __restrict__ work_n_dev, uint32_t work_cap, rjb_xsect* __restrict__ out,
                 uint32_t cap, unsigned int* counter, unsigned long long* n_cand) {
  // < kExactThreads carried over + <= kGridInLane pushes per thread and round
  __shared__ uint2 s_list[(kGridInLane + 1) * kExactThreads];
  __shared__ unsigned s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const uint32_t n = min(*work_n_dev, work_cap);
  const int lane = threadIdx.x & 31;
  unsigned cand = 0;
  // block-uniform trip count: every thread reaches the barriers
  for (uint64_t i0 = (uint64_t) blockIdx.x * kExactThreads; i0 < n; i0 += (uint64_t) gridDim.x * kExactThreads) {
    const uint64_t i = i0 + threadIdx.x;
    uint32_t pq = 0, bit = 0, beg = 0, end = 0;
    Seg eq = {0, 0, 0, 0};
    uint4 qq = make_uint4(1, 0, 1, 0);  // empty
    if (i < n) {
      const uint2 w = work[i];
      pq = w.x;
      bit = w.y;
      grid_cell_range(g, bit, beg, end);
      const longlong2 a = __ldg(&Q.pts[pq]), b = __ldg(&Q.pts[pq + 1]);
      eq = {a.x, a.y, b.x, b.y};
      qq = grid_query_qbox(g, eq, bit);
    }
    // the first kGridInLane items of the lane's own cell: the item records in flight together;
    // vertices are loaded only for the items whose packed box overlaps the query's
    uint2 it[kGridInLane];
#pragma unroll
    for (int k = 0; k < kGridInLane; k++) it[k] = beg + k < end ? __ldg(&g.items[beg + k]) : make_uint2(0xFFFFFFFFu, 0);
#pragma unroll
    for (int k = 0; k < kGridInLane; k++) {
      bool pass = false;
      if (it[k].x != 0xFFFFFFFFu && grid_qbox_overlap(qq, it[k].y)) {
        const longlong2 c = __ldg(&B.pts[it[k].x]), d = __ldg(&B.pts[it[k].x + 1]);
        const Seg eb = {c.x, c.y, d.x, d.y};
        pass = seg_boxes_overlap(eq, eb) && grid_pair_here(g, eq, eb, bit);
      }
      exact_push(pass, make_uint2(pq, it[k].x), s_list, &s_n, lane, cand);
    }
    // longer lists (dense cells): one lane's cell at a time, the warp strides over its items and
    // tests what passes the boxes at once
    unsigned more = __ballot_sync(0xffffffffu, end - beg > (uint32_t) kGridInLane);
    while (more) {
      const int src = __ffs(more) - 1;
      more &= more - 1;
      const uint32_t b0 = __shfl_sync(0xffffffffu, beg, src) + kGridInLane, e0 = __shfl_sync(0xffffffffu, end, src);
      const uint32_t spq = __shfl_sync(0xffffffffu, pq, src), sbit = __shfl_sync(0xffffffffu, bit, src);
      const Seg sq = {__shfl_sync(0xffffffffu, eq.x1, src), __shfl_sync(0xffffffffu, eq.y1, src),
                      __shfl_sync(0xffffffffu, eq.x2, src), __shfl_sync(0xffffffffu, eq.y2, src)};
      const uint4 sqq = make_uint4(__shfl_sync(0xffffffffu, qq.x, src), __shfl_sync(0xffffffffu, qq.y, src),
                                   __shfl_sync(0xffffffffu, qq.z, src), __shfl_sync(0xffffffffu, qq.w, src));
      for (uint32_t k0 = b0; k0 < e0; k0 += 32) {
        const uint32_t k = k0 + lane;
        bool pass = false;
        uint32_t p = 0;
        if (k < e0) {
          const uint2 item = __ldg(&g.items[k]);
          p = item.x;
          if (grid_qbox_overlap(sqq, item.y)) {
            const longlong2 c = __ldg(&B.pts[p]), d = __ldg(&B.pts[p + 1]);
            const Seg eb = {c.x, c.y, d.x, d.y};
            pass = seg_boxes_overlap(sq, eb) && grid_pair_here(g, sq, eb, sbit);
          }
        }
        cand += __popc(__ballot_sync(0xffffffffu, pass));
        grid_test_emit(Q, B, g, q, pass, make_uint2(spq, p), out, cap, counter, lane);
      }
    }
    __syncthreads();
    const unsigned n_list = s_n;
    if (n_list >= kExactThreads) {
      grid_drain(Q, B, g, q, s_list, n_list, out, cap, counter);
      __syncthreads();
      if (threadIdx.x == 0) s_n = 0;
      __syncthreads();
    }
  }
  __syncthreads();
  grid_drain(Q, B, g, q, s_list, s_n, out, cap, counter);
  if (lane == 0 && cand) atomicAdd(n_cand, (unsigned long long) cand);
}

// ---- PIP ----------------------------------------------------------------------------------
// One thread per point (points sorted by cell, so a warp works in one or two columns), walking
// the column upward from the cell of (py - 1): the next OCCUPIED cell is a scan of consecutive
// bitmap bits, an edge first registered above cell cy has ymin >= the lower boundary of cell
// cy, and y* carries < 2^-5 of rounding, so the walk ends once the best hit lies provably below
// everything not yet seen.  (Reference: src/app/pip_grid.h walks cell by cell, empty or not,
// through a dense Cell array.)
//
// The WALK runs on the 32-bit boxes of the item records alone (grid_qbox) and the exact update
// rule runs afterwards, converged.  While walking, a lane screens the edges of its cells -- x
// range contains px, reaches up to py - 1, both in 1/128 steps of the cell, conservative -- and
// only NOTES the ones that pass in a four-entry register queue: no vertex is loaded, nothing
// wider than 32 bits is compared.  The walk needs no exact crossing to know when to stop: an
// edge with px strictly inside its x range that starts above py + 1 is certainly accepted by the
// rule and crosses at or below its ymax, so U = min(ymax) over such edges (in 1/128 cell rows,
// rounded up) bounds the best crossing from above, and the walk ends at the first occupied cell
// that lies wholly above U.  Then all lanes put their noted edges through the rule together
// (any order: the rule's tie-breaks make it order independent).
// ncu history (20 M points, g = 8192): rule evaluated where the screen passed, every lane in a
// cell of its own: 5.0 of 32 threads active per instruction, 176 warp instructions per point;
// rule deferred, vertices screened in the walk: 9.1 threads, 105 per point.
struct PipCandQueue {
  uint32_t c0, c1, c2, c3;
  int n;
  __device__ __forceinline__ void push(uint32_t v) {
    c3 = c2; c2 = c1; c1 = c0; c0 = v;
    n++;
  }
  __device__ __forceinline__ uint32_t pop() {
    const uint32_t v = c0;
    c0 = c1; c1 = c2; c2 = c3;
    n--;
    return v;
  }
};

// best_pb: start point of the best edge so far (its chain is best_pb - best.eid)
static __device__ __forceinline__ void pip_eval(const MapView& B, PipBest& best, uint32_t& best_pb, int q,
                                                const longlong2& p, uint32_t pb, unsigned long long& cand) {
  const longlong2 a = __ldg(&B.pts[pb]), b = __ldg(&B.pts[pb + 1]);
  const Seg e = {a.x, a.y, b.x, b.y};
  cand++;
  if (pip_update(best, q, p.x, p.y, e, pb - __ldg(&B.point_chain[pb]))) best_pb = pb;
}

template <bool kPacked>
__global__ void __launch_bounds__(256)
k_pip_grid(const longlong2* __restrict__ pts, uint32_t n, const uint64_t* __restrict__ order, MapView B,
           GridView g, int query_map_id, uint32_t* __restrict__ out_eid, int32_t* __restrict__ out_face,
           uint2* __restrict__ out_packed, unsigned long long* n_cand) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cand = 0;
  const bool valid = slot < n;
  uint32_t i = 0;
  longlong2 p = make_longlong2(0, 0);
  if (valid) {
    i = order ? (uint32_t) order[slot] : slot;  // packed (key, point index) words
    p = pts[i];
  }
  PipBest best;
  pip_init(best);
  uint32_t best_pb = 0;
  const int gmax = (int) g.gy;
  const int cxp = grid_cx(g, p.x);
  const uint32_t col = (uint32_t) cxp * g.gs;
  const uint32_t qpx = grid_qx(g, p.x, cxp);
  int cy = grid_cy(g, p.y - 1);
  PipCandQueue Q = {0, 0, 0, 0, 0};
  if (valid) {
    // 1/128 cell rows: row of cell c = c * 128.  (py + 1, rounded up to the next step, must lie
    // BELOW the start of an edge for the edge to count as certainly above the point.)
    const int cy_lo = grid_cy(g, p.y + 1);
    const uint32_t y_above = (uint32_t) cy_lo * kGridQ + grid_qy(g, p.y + 1, cy_lo);
    uint32_t U = 0xFFFFFF00u;  // upper bound of the best crossing (none yet; U + 1 must not wrap)
    const uint32_t qpy = grid_qy(g, p.y - 1, cy);
    const int cy_start = cy;
    while (cy < gmax) {
      // next occupied cell at or above cy in this column
      uint32_t bit = col + cy;
      uint32_t w = __ldg(&g.bits[bit >> 5]) >> (bit & 31);
      while (w == 0) {
        bit = (bit | 31u) + 1;
        if ((int) (bit - col) >= gmax) break;
        w = __ldg(&g.bits[bit >> 5]);
      }
      if (w == 0) break;
      bit += __ffs(w) - 1;
      cy = (int) (bit - col);
      if (cy >= gmax) break;
      // every edge first met in this or a higher cell starts at or above row cy * 128; one step
      // of slack covers the +-3 units between the rows of the cell function and the rounding of y*
      if ((uint32_t) cy * kGridQ > U + 1) break;
      uint32_t beg = 0, end = 0;
      grid_cell_range(g, bit, beg, end);
      const uint32_t row0 = (uint32_t) cy * kGridQ;
      const uint32_t qy_min = cy == cy_start ? qpy : 0u;  // higher cells lie above py - 1 anyway
      for (uint32_t k = beg; k < end; k++) {
        const uint2 it = __ldg(&g.items[k]);
        const uint32_t qx0 = it.y & 63u, qx1 = (it.y >> 6) & 63u, qy0 = (it.y >> 14) & 127u;
        const uint32_t qy1 = (it.y >> 21) & 127u, dy = it.y >> 28;
        if (qx0 > qpx || qx1 < qpx || (dy == 0 && qy1 < qy_min)) continue;
        // certainly above the point, crossing at or below the end of step qy1 of row cy + dy.
        // (qy0 == 0 may mean "starts below this cell": such an edge was judged in the cell of its
        // ymin, or starts below the walk altogether; it is not judged again.)
        const bool x_strict = (((it.y >> 12) & 1u) || qx0 < qpx) && (((it.y >> 13) & 1u) || qpx < qx1);
        if (x_strict && qy0 > 0 && row0 + qy0 > y_above && dy < 15) U = min(U, row0 + dy * kGridQ + qy1 + 1);
        // an edge spanning several cells is met again: keep it once
        if (Q.n > 0 && (Q.c0 == it.x || (Q.n > 1 && Q.c1 == it.x))) continue;
        if (Q.n == 4) pip_eval(B, best, best_pb, query_map_id, p, Q.c3, cand), Q.n = 3;  // queue full (rare)
        Q.push(it.x);
      }
      cy++;
    }
  }
  // the exact update rule, converged: first every lane's first noted edge, then the second ...
  while (__any_sync(0xffffffffu, Q.n > 0)) {
    if (Q.n > 0) pip_eval(B, best, best_pb, query_map_id, p, Q.pop(), cand);
  }
  if (valid) {
    int32_t face = RJB_EXTERIOR_FACE;
    if ((kPacked || out_face) && best.eid != RJB_NO_HIT) {
      // get_face_id, src/map/map.h:79-87 (chain = start point - eid)
      const uint32_t ch = best_pb - best.eid;
      const longlong2 a = __ldg(&B.pts[best_pb]), bb = __ldg(&B.pts[best_pb + 1]);
      face = a.x < bb.x ? __ldg(&B.right[ch]) : __ldg(&B.left[ch]);
    }
    if (kPacked) {
      out_packed[i] = make_uint2(best.eid, (uint32_t) face);
    } else {
      out_eid[i] = best.eid;
      if (out_face) out_face[i] = face;
    }
  }
  if (n_cand) {  // kCtrSlots slots, summed on the host
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
    if ((threadIdx.x & 31) == 0 && cand) atomicAdd(n_cand + (blockIdx.x & (kCtrSlots - 1)), cand);
  }
}

}  // namespace rjb
