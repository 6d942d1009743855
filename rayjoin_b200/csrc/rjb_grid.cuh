// Uniform-grid variant (RayJoin's -mode=grid): the base map's edges are binned
// into gsize x gsize cells (every cell the edge's bounding box touches, like
// reference src/grid/uniform_grid.h:44-86), stored as CSR; LSI is
// query-edge-centric, PIP walks the point's column upward.
//
// Replaces UniformGrid::AddMapToGrid (src/grid/uniform_grid.h:131-358),
// LSIGrid::Query (src/app/lsi_grid.h:19-78,97-159: one thread per CELL doing
// ne0 x ne1 tests -> 1.3 s load imbalance) and PIPGrid::Query
// (src/app/pip_grid.h:21-77 + src/algo/pip.h:14-115).
// The grid is only a filter; results come from the exact predicates, so they
// equal the LBVH / brute-force results pair for pair.
#pragma once
#include "rjb_exact.cuh"
#include "rjb_lsi.cuh"
#include "rjb_prims.cuh"

namespace rjb {

struct GridView {
  const uint32_t* cell_begin;  // gsize*gsize + 1
  const uint32_t* items;       // base eids
  uint32_t gsize;
  long long imin;
};

struct Grid {
  DBuf<uint32_t> cell_begin, items, cursor;
  ScanTemp scan_tmp;
  uint32_t gsize = 0;
  uint64_t n_items = 0;
  long long imin = 0;
  bool built = false;
  GridView view() const {
    GridView v;
    v.cell_begin = cell_begin.p;
    v.items = items.p;
    v.gsize = gsize;
    v.imin = imin;
    return v;
  }
  size_t index_bytes() const {
    return ((size_t) gsize * gsize + 1) * sizeof(uint32_t) + n_items * sizeof(uint32_t);
  }
};

// cell of an internal coordinate: floor((v - imin) * gsize / 2^47), monotone,
// in [0, gsize - 1] for the whole 47-bit range
static __device__ __forceinline__ int grid_cell(long long v, long long imin, uint32_t gsize) {
  long long d = v - imin;
  d = d < 0 ? 0 : (d > (1ll << 47) - 1 ? (1ll << 47) - 1 : d);  // points outside the box
  return (int) (((unsigned long long) d * gsize) >> 47);
}

// smallest internal coordinate that lies in cell c (c may be gsize: upper end)
static __device__ __forceinline__ long long grid_cell_lo(int c, long long imin, uint32_t gsize) {
  unsigned long long t = (((unsigned long long) c << 47) + gsize - 1) / gsize;
  return imin + (long long) t;
}

template <bool kFill>
__global__ void k_grid_bin(MapView B, GridView g, uint32_t* __restrict__ counts_or_cursor,
                           uint32_t* __restrict__ items) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B.n_edges) return;
  Seg s = load_seg(B, e);
  int cx0 = grid_cell(min(s.x1, s.x2), g.imin, g.gsize), cx1 = grid_cell(max(s.x1, s.x2), g.imin, g.gsize);
  int cy0 = grid_cell(min(s.y1, s.y2), g.imin, g.gsize), cy1 = grid_cell(max(s.y1, s.y2), g.imin, g.gsize);
  for (int cy = cy0; cy <= cy1; cy++)
    for (int cx = cx0; cx <= cx1; cx++) {
      size_t c = (size_t) cy * g.gsize + cx;
      uint32_t pos = atomicAdd(&counts_or_cursor[c], 1u);
      if (kFill) items[pos] = e;
    }
}

static inline void build_grid(Grid& g, const MapView& B, uint32_t gsize, long long imin,
                              long long /*irange*/, cudaStream_t st) {
  RJB_REQUIRE(gsize >= 1 && gsize <= 32768, "grid_size must be in 1..32768");
  g.gsize = gsize;
  g.imin = imin;
  g.built = true;
  g.n_items = 0;
  size_t ncell = (size_t) gsize * gsize;
  RJB_REQUIRE(ncell < 0xFFFFFFF0ull, "grid too large");
  uint32_t* begin = g.cell_begin.ensure(ncell + 1);
  uint32_t* cursor = g.cursor.ensure(ncell + 1);
  RJB_CUDA(cudaMemsetAsync(cursor, 0, (ncell + 1) * sizeof(uint32_t), st));
  GridView v = g.view();
  if (B.n_edges) k_grid_bin<false><<<div_up(B.n_edges, 256), 256, 0, st>>>(B, v, cursor, nullptr);
  exclusive_scan_u32(cursor, begin, (uint32_t) ncell, g.scan_tmp, st);
  uint32_t total = 0;
  RJB_CUDA(cudaMemcpyAsync(&total, begin + ncell, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  RJB_CUDA(cudaStreamSynchronize(st));
  g.n_items = total;
  uint32_t* items = g.items.ensure(total ? total : 1);
  RJB_CUDA(cudaMemcpyAsync(cursor, begin, ncell * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  v = g.view();
  if (B.n_edges) k_grid_bin<true><<<div_up(B.n_edges, 256), 256, 0, st>>>(B, v, cursor, items);
  RJB_CUDA(cudaGetLastError());
}

// one thread per query edge; every (query, base) pair is examined in exactly
// one cell: the cell of the lower-left corner of the intersection of the two
// bounding boxes (both edges are registered there)
__global__ void __launch_bounds__(256)
k_lsi_grid(MapView Q, MapView B, GridView g, uint2* __restrict__ out, uint32_t cap,
           unsigned int* counter, unsigned long long* n_cand) {
  uint32_t qe = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cand = 0;
  if (qe < Q.n_edges) {
    Seg q = load_seg(Q, qe);
    long long qx0 = min(q.x1, q.x2), qx1 = max(q.x1, q.x2);
    long long qy0 = min(q.y1, q.y2), qy1 = max(q.y1, q.y2);
    int cx0 = grid_cell(qx0, g.imin, g.gsize), cx1 = grid_cell(qx1, g.imin, g.gsize);
    int cy0 = grid_cell(qy0, g.imin, g.gsize), cy1 = grid_cell(qy1, g.imin, g.gsize);
    for (int cy = cy0; cy <= cy1; cy++)
      for (int cx = cx0; cx <= cx1; cx++) {
        size_t c = (size_t) cy * g.gsize + cx;
        uint32_t b = g.cell_begin[c], e = g.cell_begin[c + 1];
        for (uint32_t k = b; k < e; k++) {
          uint32_t be = g.items[k];
          Seg s = load_seg(B, be);
          long long bx0 = min(s.x1, s.x2), bx1 = max(s.x1, s.x2);
          long long by0 = min(s.y1, s.y2), by1 = max(s.y1, s.y2);
          if (bx1 < qx0 || qx1 < bx0 || by1 < qy0 || qy1 < by0) continue;
          if (grid_cell(max(qx0, bx0), g.imin, g.gsize) != cx ||
              grid_cell(max(qy0, by0), g.imin, g.gsize) != cy)
            continue;
          cand++;
          if (lsi_intersect(q, s)) {
            unsigned pos = atomicAdd(counter, 1u);
            if (pos < cap) out[pos] = make_uint2(qe, be);
          }
        }
      }
  }
  if (n_cand) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
    if ((threadIdx.x & 31) == 0 && cand) atomicAdd(n_cand, cand);
  }
}

static inline void lsi_grid(const Grid& g, const MapView& Q, const MapView& B, uint2* out,
                            uint32_t cap, unsigned int* counter, unsigned long long* n_cand,
                            cudaStream_t st) {
  k_lsi_grid<<<div_up(Q.n_edges, 256), 256, 0, st>>>(Q, B, g.view(), out, cap, counter, n_cand);
}

// one thread per point: walk the column upward from the cell of (py - 1) and
// stop once the best hit is provably below the top of the visited cell
__global__ void __launch_bounds__(256)
k_pip_grid(const longlong2* __restrict__ pts, uint32_t n, MapView B, GridView g, int query_map_id,
           uint32_t* __restrict__ out_eid, int32_t* __restrict__ out_face,
           unsigned long long* n_cand) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cand = 0;
  if (i < n) {
    longlong2 p = pts[i];
    PipBest best;
    pip_init(best);
    int cx = grid_cell(p.x, g.imin, g.gsize);
    for (int cy = grid_cell(p.y - 1, g.imin, g.gsize); cy < (int) g.gsize; cy++) {
      size_t c = (size_t) cy * g.gsize + cx;
      uint32_t b = g.cell_begin[c], e = g.cell_begin[c + 1];
      for (uint32_t k = b; k < e; k++) {
        uint32_t be = g.items[k];
        cand++;
        pip_update(best, query_map_id, p.x, p.y, load_seg(B, be), be);
      }
      if (best.eid != RJB_NO_HIT) {
        double top = (double) grid_cell_lo(cy + 1, g.imin, g.gsize) - 1.0;
        if (best.y < top) break;
      }
    }
    out_eid[i] = best.eid;
    if (out_face) {
      int32_t face = RJB_EXTERIOR_FACE;
      if (best.eid != RJB_NO_HIT) {
        uint32_t ch = B.edge_chain[best.eid];
        longlong2 a = B.pts[best.eid + ch], bb = B.pts[best.eid + ch + 1];
        face = a.x < bb.x ? B.right[ch] : B.left[ch];
      }
      out_face[i] = face;
    }
  }
  if (n_cand) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
    if ((threadIdx.x & 31) == 0 && cand) atomicAdd(n_cand, cand);
  }
}

static inline void pip_grid(const Grid& g, const longlong2* pts, uint32_t n, const MapView& B,
                            int query_map_id, uint32_t* out_eid, int32_t* out_face,
                            unsigned long long* n_cand, cudaStream_t st) {
  k_pip_grid<<<div_up(n, 256), 256, 0, st>>>(pts, n, B, g.view(), query_map_id, out_eid, out_face,
                                             n_cand);
}

}  // namespace rjb
