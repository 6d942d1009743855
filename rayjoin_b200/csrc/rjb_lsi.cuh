// LSI kernels: warp-cooperative BVH traversal, all-pairs reference kernel,
// and the intersection-point post-pass.
//
// Replaces LSILBVH::Query (reference: src/app/lsi_lbvh.h:27-98 +
// deps/lbvh/lbvh/query.cuh:8-51, one thread per query edge with a 64-entry
// local-memory stack and three dependent AoS loads per node).
#pragma once
#include "rjb_exact.cuh"

namespace rjb {

constexpr int kLsiWarps = 8;    // warps per CTA
constexpr int kStackDepth = 96; // >= max LBVH depth (64 key bits + 32 index bits)

static __device__ __forceinline__ Seg load_seg(const MapView& m, uint32_t eid) {
  uint32_t p = eid + m.edge_chain[eid];
  longlong2 a = m.pts[p], b = m.pts[p + 1];
  Seg s = {a.x, a.y, b.x, b.y};
  return s;
}

// Warp-aggregated append of (query eid, base eid) to the result queue.
static __device__ __forceinline__ void emit_pair(bool found, uint32_t q, uint32_t b,
                                                 uint2* __restrict__ out, uint32_t cap,
                                                 unsigned int* counter, int lane) {
  unsigned m = __ballot_sync(0xffffffffu, found);
  if (m == 0) return;
  unsigned base = 0;
  int leader = __ffs(m) - 1;
  if (lane == leader) base = atomicAdd(counter, (unsigned) __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (found) {
    unsigned pos = base + __popc(m & ((1u << lane) - 1));
    if (pos < cap) out[pos] = make_uint2(q, b);
  }
}

// One warp = 32 query edges.  The warp walks the BVH ONCE for all of them:
// node records are loaded at a warp-uniform address (one transaction,
// broadcast), every lane tests its own query box against both child boxes,
// and __ballot_sync decides warp-uniformly which children to enter.  No lane
// ever diverges in the traversal loop and the stack is a single warp-shared
// array in shared memory.  `order` (optional) maps slot -> query eid so that
// the 32 edges of a warp are spatial neighbours.
template <bool kStats>
__global__ void __launch_bounds__(kLsiWarps * 32)
k_lsi_bvh(MapView Q, MapView B, BvhView bvh, const uint32_t* __restrict__ order,
          uint2* __restrict__ out, uint32_t cap, unsigned int* counter,
          unsigned long long* n_cand) {
  __shared__ int s_stack[kLsiWarps][kStackDepth];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stack = s_stack[warp];
  const uint32_t slot = (blockIdx.x * kLsiWarps + warp) * 32 + lane;
  const bool valid = slot < Q.n_edges;
  uint32_t qe = 0;
  Seg q = {0, 0, 0, 0};
  int4 qb = make_int4(1, 1, 0, 0);  // empty box for idle lanes
  if (valid) {
    qe = order ? order[slot] : slot;
    q = load_seg(Q, qe);
    qb = make_int4(quant(min(q.x1, q.x2)), quant(min(q.y1, q.y2)),
                   quant(max(q.x1, q.x2)), quant(max(q.y1, q.y2)));
  }
  unsigned long long cand = 0;
  // traversal statistics (kStats only): node visits, leaf visits, length of the
  // initial single-child descent, lane-level leaf tests, deepest stack
  unsigned st_nodes = 0, st_leaves = 0, st_prefix = 0, st_lane_leaf = 0, st_maxsp = 0;
  bool st_in_prefix = true;
  if (__ballot_sync(0xffffffffu, box_overlap(qb, bvh.root_box)) != 0) {
    int sp = 0;
    int node = 0;
    while (true) {
      const int4 lb = __ldg(&bvh.node_box[2 * node]);
      const int4 rb = __ldg(&bvh.node_box[2 * node + 1]);
      const int2 ch = __ldg(&bvh.node_child[node]);
      const bool hl = box_overlap(qb, lb), hr = box_overlap(qb, rb);
      const unsigned ml = __ballot_sync(0xffffffffu, hl);
      const unsigned mr = __ballot_sync(0xffffffffu, hr);
      if (kStats) {
        st_nodes++;
        bool single = ((ml != 0) != (mr != 0)) && ((ml ? ch.x : ch.y) >= 0);
        if (st_in_prefix && single) st_prefix++; else st_in_prefix = false;
      }
      int next = -1;
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const unsigned m = side ? mr : ml;
        const int c = side ? ch.y : ch.x;
        const bool h = side ? hr : hl;
        if (m == 0) continue;
        if (c >= 0) {  // internal child: enter now or later
          if (next < 0) next = c; else stack[sp++] = c;
          continue;
        }
        // leaf: the hit lanes test their edge against its <= 8 base edges,
        // whose vertices are one contiguous run of points (uniform loads)
        if (kStats) { st_leaves++; st_lane_leaf += __popc(m); }
        const uint2 rec = __ldg(&bvh.leaf_rec[~c]);
        const uint32_t first_eid = rec.x, cnt = rec.y >> 28, chain = rec.y & 0x0FFFFFFFu;
        const longlong2* bp = B.pts + (first_eid + chain);
        longlong2 p1 = __ldg(bp);
        for (uint32_t k = 0; k < cnt; k++) {
          const longlong2 p2 = __ldg(bp + k + 1);
          const Seg e2 = {p1.x, p1.y, p2.x, p2.y};
          bool found = false;
          if (h && seg_boxes_overlap(q, e2)) {
            cand++;
            found = lsi_intersect(q, e2);
          }
          emit_pair(found, qe, first_eid + k, out, cap, counter, lane);
          p1 = p2;
        }
      }
      if (kStats) st_maxsp = max(st_maxsp, (unsigned) sp);
      if (next >= 0) { node = next; continue; }
      if (sp == 0) break;
      node = stack[--sp];
    }
  }
  if (n_cand) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
    if (lane == 0 && cand) atomicAdd(n_cand, cand);
    if (kStats && lane == 0) {  // n_cand points at counters[1]; stats live at counters[2..7]
      atomicAdd(n_cand + 1, (unsigned long long) st_nodes);
      atomicAdd(n_cand + 2, (unsigned long long) st_leaves);
      atomicAdd(n_cand + 3, (unsigned long long) st_prefix);
      atomicAdd(n_cand + 4, (unsigned long long) st_lane_leaf);
      atomicAdd(n_cand + 5, (unsigned long long) (st_leaves ? 1 : 0));
      atomicMax(n_cand + 6, (unsigned long long) st_maxsp);
    }
  }
}

// All |Q| x |B| pairs, no index: pins the exact arithmetic (RJB_MODE_BRUTE).
// Each CTA stages 256 base edges in shared memory; each thread owns one query edge.
__global__ void __launch_bounds__(256)
k_lsi_brute(MapView Q, MapView B, uint2* __restrict__ out, uint32_t cap,
            unsigned int* counter, unsigned long long* n_cand) {
  __shared__ Seg s_b[256];
  const int lane = threadIdx.x & 31;
  const uint32_t qe = blockIdx.x * 256 + threadIdx.x;
  const bool valid = qe < Q.n_edges;
  Seg q = {0, 0, 0, 0};
  if (valid) q = load_seg(Q, qe);
  unsigned long long cand = 0;
  for (uint32_t b0 = blockIdx.y * 256; b0 < B.n_edges; b0 += gridDim.y * 256) {
    __syncthreads();
    uint32_t be = b0 + threadIdx.x;
    if (be < B.n_edges) s_b[threadIdx.x] = load_seg(B, be);
    __syncthreads();
    uint32_t nb = min(256u, B.n_edges - b0);
    for (uint32_t k = 0; k < nb; k++) {
      bool found = false;
      if (valid) {
        cand++;
        found = lsi_intersect(q, s_b[k]);
      }
      emit_pair(found, qe, b0 + k, out, cap, counter, lane);
    }
  }
  if (n_cand) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
    if (lane == 0 && cand) atomicAdd(n_cand, cand);
  }
}

// Intersection points of the found pairs -> rjb_xsect records
// (reference: src/app/lsi_rt.h:66-112 does the same as a post-pass; the LBVH
// backend computes it inside the traversal callback, lsi_lbvh.h:69-78).
// The pair count is read from the device counter so that no host round trip
// separates the traversal from this pass (grid is sized by the capacity).
__global__ void __launch_bounds__(128)
k_xsect_points_dyn(MapView Q, MapView B, int query_map_id, const uint2* __restrict__ pairs,
                   const unsigned int* __restrict__ counter, uint32_t cap,
                   rjb_xsect* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = min(*counter, cap);
  if (i >= n) return;
  uint2 pr = pairs[i];
  Seg e1 = load_seg(Q, pr.x), e2 = load_seg(B, pr.y);
  long long x, y;
  lsi_point(e1, e2, x, y);
  rjb_xsect r;
  r.x = x;
  r.y = y;
  r.eid[0] = query_map_id == 0 ? pr.x : pr.y;
  r.eid[1] = query_map_id == 0 ? pr.y : pr.x;
  r.mid_point_polygon_id = RJB_DONTKNOW;
  r._pad = 0;
  out[i] = r;
}

}  // namespace rjb
