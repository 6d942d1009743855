// PIP kernels: "closest edge above the point" (RayJoin's PIP rule), as a
// warp-cooperative, front-to-back BVH traversal pruned by the best hit so far.
//
// Replaces PIPLBVH::Query (reference: src/app/pip_lbvh.h:25-142), which visits
// EVERY leaf whose box is above the point (box y in [y - 3ulp, FLT_MAX]) and
// is therefore O(edges above the point) per query.
#pragma once
#include "rjb_exact.cuh"
#include "rjb_lsi.cuh"

namespace rjb {

// quantised upper bound of everything that can still tie or beat best_y:
// y*(double) differs from the exact crossing by < 2^-5, and the exact
// crossing is >= the edge's (and node's) ymin, so a node whose quantised ymin
// exceeds quant(floor(best_y) + 2) cannot contain a better or tying edge.
static __device__ __forceinline__ int pip_y_bound(double best_y) {
  if (!(best_y < 1.0e15)) return 0x7fffffff;
  return quant((long long) floor(best_y) + 2);
}

// per-lane query state of the PIP traversal
struct PipLane {
  long long px, py;
  int qx, qy_lo, y_hi;  // quantised: x, lower y bound (py - 1), pruning bound
  bool valid;
  PipBest best;
  // one leaf whose box the lane's ray meets, parked for pip_flush (or RJB_NO_HIT)
  uint32_t pend_leaf;
};

// lane wants a box: its x range contains px, it reaches up to py, and it starts
// below the lane's current best hit
static __device__ __forceinline__ bool pip_wants(const PipLane& L, const int4& b) {
  return L.valid && b.x <= L.qx && L.qx <= b.z && b.w >= L.qy_lo && b.y <= L.y_hi;
}

// Opens the parked leaf of every lane that has one -- all lanes in parallel, each on
// ITS OWN leaf (per-lane gathers of <= 9 consecutive points).  Stage 1 screens the
// leaf's edges with integer tests (x range contains px, reaches up to py - 1, starts
// below the pruning bound) into a bit mask; stage 2 runs the exact update rule on the
// pending edges, every lane on its own, until no lane has one left.
static __device__ __forceinline__ void pip_flush(const MapView& B, const BvhView& bvh, PipLane& L,
                                                 int q, unsigned long long& cand) {
  unsigned mask = 0;
  uint32_t first_eid = 0;
  const longlong2* bp = B.pts;
  if (L.pend_leaf != RJB_NO_HIT) {
    const uint2 rec = __ldg(&bvh.leaf_rec[L.pend_leaf]);
    first_eid = rec.x;
    const uint32_t cnt = rec.y >> 28;
    bp = B.pts + (first_eid + (rec.y & 0x0FFFFFFFu));
    longlong2 p1 = __ldg(bp);
    for (uint32_t k = 0; k < cnt; k++) {
      const longlong2 p2 = __ldg(bp + k + 1);
      if (min(p1.x, p2.x) <= L.px && L.px <= max(p1.x, p2.x) && max(p1.y, p2.y) >= L.py - 1 &&
          quant(min(p1.y, p2.y)) <= L.y_hi)
        mask |= 1u << k;
      p1 = p2;
    }
    L.pend_leaf = RJB_NO_HIT;
  }
  while (__ballot_sync(0xffffffffu, mask != 0)) {
    if (mask) {
      const uint32_t k = __ffs(mask) - 1;
      mask &= mask - 1;
      const longlong2 a = __ldg(bp + k), b = __ldg(bp + k + 1);
      if (quant(min(a.y, b.y)) <= L.y_hi) {  // the bound may have shrunk meanwhile
        const Seg e = {a.x, a.y, b.x, b.y};
        cand++;
        if (pip_update(L.best, q, L.px, L.py, e, first_eid + k)) L.y_hi = pip_y_bound(L.best.y);
      }
    }
  }
}

// A leaf is <= 8 consecutive edges of one chain.  kPark (option "pip_park"): ncu showed
// this kernel to be instruction bound (13 k warp-instructions per warp, 15 of 32
// threads active): on the 100 M-point workload the 32 points of a warp need ~30
// DIFFERENT leaves, so opening a leaf when the walk reaches it runs the screening and
// the int128/double update rule with one or two active lanes, 30 times over.  Now the
// walk only PARKS the leaf in the slot of every lane whose ray meets its box; parked
// leaves are opened together (pip_flush) when a lane needs its slot again, when most
// lanes hold one, and at the end.  A parked lane keeps its stale (larger) pruning
// bound in the meantime, which only makes the warp visit more, never less.
template <bool kStats, bool kPark>
static __device__ __forceinline__ void pip_leaf(const MapView& B, const BvhView& bvh, int leaf, bool h,
                                                PipLane& L, int q, unsigned long long& cand,
                                                TravStats& st) {
  if (kStats) st.leaves++;
  if (!kPark) {
    // open the leaf at once: every lane whose ray meets the box scans its edges
    const uint2 rec = __ldg(&bvh.leaf_rec[leaf]);
    const uint32_t first_eid = rec.x, cnt = rec.y >> 28, chain = rec.y & 0x0FFFFFFFu;
    const longlong2* bp = B.pts + (first_eid + chain);
    longlong2 p1 = __ldg(bp);
    for (uint32_t k = 0; k < cnt; k++) {
      const longlong2 p2 = __ldg(bp + k + 1);
      // cheap integer rejections before the int128 / double arithmetic: the edge must
      // span px in x, reach up to py - 1, and start below the best crossing so far
      // (same +-1 margins as the box test, exact in the integer domain)
      if (h && min(p1.x, p2.x) <= L.px && L.px <= max(p1.x, p2.x) && max(p1.y, p2.y) >= L.py - 1 &&
          quant(min(p1.y, p2.y)) <= L.y_hi) {
        const Seg e = {p1.x, p1.y, p2.x, p2.y};
        cand++;
        if (pip_update(L.best, q, L.px, L.py, e, first_eid + k)) L.y_hi = pip_y_bound(L.best.y);
      }
      p1 = p2;
    }
    return;
  }
  const unsigned mh = __ballot_sync(0xffffffffu, h);
  if (__ballot_sync(0xffffffffu, h && L.pend_leaf != RJB_NO_HIT)) pip_flush(B, bvh, L, q, cand);
  if (h) L.pend_leaf = (uint32_t) leaf;
  // open at once when the leaf is shared by many lanes (coherent queries such as the
  // vertices of a chain: fresh bounds prune better) or when most lanes hold a leaf
  if (__popc(mh) >= 12 || __popc(__ballot_sync(0xffffffffu, L.pend_leaf != RJB_NO_HIT)) >= 24)
    pip_flush(B, bvh, L, q, cand);
}

// binary subtree, near child (lower ymin) first; boxes are re-tested when a node
// is reached because the pruning bounds shrink while the warp works
template <bool kStats, bool kPark>
static __device__ __forceinline__ void pip_subtree(const MapView& B, const BvhView& bvh, int root,
                                                   int* stack, PipLane& L, int q,
                                                   unsigned long long& cand, TravStats& st) {
  int sp = 0;
  int node = root;
  while (true) {
    const int4 lb = __ldg(&bvh.node_box[2 * node]);
    const int4 rb = __ldg(&bvh.node_box[2 * node + 1]);
    const int2 ch = __ldg(&bvh.node_child[node]);
    const bool swap = rb.y < lb.y;
    const int4 b0 = swap ? rb : lb, b1 = swap ? lb : rb;
    const int c0 = swap ? ch.y : ch.x, c1 = swap ? ch.x : ch.y;
    if (kStats) st.nodes++;
    int next = -1;
    {
      const bool h = pip_wants(L, b0);
      const unsigned m = __ballot_sync(0xffffffffu, h);
      if (m) {
        if (c0 >= 0) next = c0;
        else {
          if (kStats) st.lane_leaf += __popc(m);
          pip_leaf<kStats, kPark>(B, bvh, ~c0, h, L, q, cand, st);
        }
      }
    }
    {
      const bool h = pip_wants(L, b1);  // after the near child: bounds may have shrunk
      const unsigned m = __ballot_sync(0xffffffffu, h);
      if (m) {
        if (c1 >= 0) {
          if (next < 0) next = c1; else stack[sp++] = c1;
        } else {
          if (kStats) st.lane_leaf += __popc(m);
          pip_leaf<kStats, kPark>(B, bvh, ~c1, h, L, q, cand, st);
        }
      }
    }
    if (kStats) st.maxsp = max(st.maxsp, (unsigned) sp);
    if (next >= 0) { node = next; continue; }
    if (sp == 0) break;
    node = stack[--sp];
  }
}

// the slot of a 32-ary top level that starts lowest among the candidates in m
static __device__ __forceinline__ int lowest_slot(unsigned m, int ymin, int lane) {
  const int key = ((m >> lane) & 1u) ? ymin : 0x7fffffff;
  const int best = __reduce_min_sync(0xffffffffu, key);
  return __ffs(__ballot_sync(0xffffffffu, key == best && ((m >> lane) & 1u))) - 1;
}

#ifndef RJB_PIP_MIN_CTAS
#define RJB_PIP_MIN_CTAS 8
#endif
template <bool kStats, bool kPark>
__global__ void __launch_bounds__(kLsiWarps * 32, RJB_PIP_MIN_CTAS)
k_pip_bvh(const longlong2* __restrict__ pts, uint32_t n_pts, const uint64_t* __restrict__ order,
          MapView B, BvhView bvh, int query_map_id, uint32_t* __restrict__ out_eid,
          int32_t* __restrict__ out_face, uint2* __restrict__ out_packed, unsigned long long* counters) {
  __shared__ int s_stack[kLsiWarps][kStackDepth];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stack = s_stack[warp];
  const uint32_t slot = (blockIdx.x * kLsiWarps + warp) * 32 + lane;
  PipLane L;
  L.valid = slot < n_pts;
  L.px = L.py = 0;
  L.qx = L.qy_lo = 0;
  L.y_hi = 0x7fffffff;
  uint32_t pi = 0;
  if (L.valid) {
    pi = order ? (uint32_t) order[slot] : slot;  // packed (key, point index) words
    const longlong2 p = pts[pi];
    L.px = p.x;
    L.py = p.y;
    L.qx = quant(p.x);
    L.qy_lo = quant(p.y - 1);
  }
  pip_init(L.best);
  L.pend_leaf = RJB_NO_HIT;
  unsigned long long cand = 0;
  TravStats st = {0, 0, 0, 0, 0};
  // the union of the lanes' rays: x range of the points, from the lowest point up
  const int ux0 = __reduce_min_sync(0xffffffffu, L.valid ? L.qx : 0x7fffffff);
  const int ux1 = __reduce_max_sync(0xffffffffu, L.valid ? L.qx : (int) 0x80000000);
  const int uy0 = __reduce_min_sync(0xffffffffu, L.valid ? L.qy_lo : 0x7fffffff);
  auto pre = [&](const int4& b) { return b.x <= ux1 && ux0 <= b.z && b.w >= uy0; };
  if (bvh.n_leaves > 0 && pre(bvh.root_box)) {
    // three levels of the 32-ary top tree (one lane per slot); slots are taken in
    // order of their lower y bound so that near hits prune far slots
    const int4 b0 = __ldg(&bvh.top_box[kTopOff0 + lane]);
    const int c0 = __ldg(&bvh.top_code[kTopOff0 + lane]);
    unsigned m0 = __ballot_sync(0xffffffffu, pre(b0));
    if (kStats) st.top_steps++;
    while (m0) {
      const int g = lowest_slot(m0, b0.y, lane);
      m0 &= ~(1u << g);
      const bool h0 = pip_wants(L, shfl_box(b0, g));
      if (__ballot_sync(0xffffffffu, h0) == 0) continue;
      const int code0 = __shfl_sync(0xffffffffu, c0, g);
      if (code0 < 0) { pip_leaf<kStats, kPark>(B, bvh, ~code0, h0, L, query_map_id, cand, st); continue; }
      const int4 b1 = __ldg(&bvh.top_box[kTopOff1 + g * 32 + lane]);
      const int c1 = __ldg(&bvh.top_code[kTopOff1 + g * 32 + lane]);
      unsigned m1 = __ballot_sync(0xffffffffu, pre(b1));
      if (kStats) st.top_steps++;
      while (m1) {
        const int h = lowest_slot(m1, b1.y, lane);
        m1 &= ~(1u << h);
        const bool h1 = pip_wants(L, shfl_box(b1, h));
        if (__ballot_sync(0xffffffffu, h1) == 0) continue;
        const int code1 = __shfl_sync(0xffffffffu, c1, h);
        if (code1 < 0) { pip_leaf<kStats, kPark>(B, bvh, ~code1, h1, L, query_map_id, cand, st); continue; }
        const int4 b2 = __ldg(&bvh.top_box[kTopOff2 + (g * 32 + h) * 32 + lane]);
        const int c2 = __ldg(&bvh.top_code[kTopOff2 + (g * 32 + h) * 32 + lane]);
        unsigned m2 = __ballot_sync(0xffffffffu, pre(b2));
        if (kStats) st.top_steps++;
        while (m2) {
          const int i = lowest_slot(m2, b2.y, lane);
          m2 &= ~(1u << i);
          const bool h2 = pip_wants(L, shfl_box(b2, i));
          if (__ballot_sync(0xffffffffu, h2) == 0) continue;
          const int code2 = __shfl_sync(0xffffffffu, c2, i);
          if (code2 < 0) { pip_leaf<kStats, kPark>(B, bvh, ~code2, h2, L, query_map_id, cand, st); continue; }
          if (bvh.top_levels < 4) {
            pip_subtree<kStats, kPark>(B, bvh, code2, stack, L, query_map_id, cand, st);
            continue;
          }
          // level 3 (big trees only): the 32 depth-20 nodes below slot (g, h, i)
          const uint32_t s3 = kTopOff3 + ((uint32_t) ((g * 32 + h) * 32 + i)) * 32 + lane;
          const int4 b3 = __ldg(&bvh.top_box[s3]);
          const int c3 = __ldg(&bvh.top_code[s3]);
          unsigned m3 = __ballot_sync(0xffffffffu, pre(b3));
          if (kStats) st.top_steps++;
          while (m3) {
            const int j = lowest_slot(m3, b3.y, lane);
            m3 &= ~(1u << j);
            const bool h3 = pip_wants(L, shfl_box(b3, j));
            if (__ballot_sync(0xffffffffu, h3) == 0) continue;
            const int code3 = __shfl_sync(0xffffffffu, c3, j);
            if (code3 < 0) pip_leaf<kStats, kPark>(B, bvh, ~code3, h3, L, query_map_id, cand, st);
            else pip_subtree<kStats, kPark>(B, bvh, code3, stack, L, query_map_id, cand, st);
          }
        }
      }
    }
  }
  if (kPark) pip_flush(B, bvh, L, query_map_id, cand);  // whatever is still parked
  if (L.valid) {
    int32_t face = RJB_EXTERIOR_FACE;
    if ((out_face || out_packed) && L.best.eid != RJB_NO_HIT) {
      // get_face_id, src/map/map.h:79-87
      uint32_t c = B.edge_chain[L.best.eid];
      longlong2 a = B.pts[L.best.eid + c], b = B.pts[L.best.eid + c + 1];
      face = a.x < b.x ? B.right[c] : B.left[c];
    }
    if (out_packed) {
      // ordered queries: ONE scattered 8-byte store per point (k_pip_split separates the
      // two result arrays afterwards with coalesced accesses) instead of two 4-byte ones
      out_packed[pi] = make_uint2(L.best.eid, (uint32_t) face);
    } else {
      out_eid[pi] = L.best.eid;
      if (out_face) out_face[pi] = face;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
  if (lane == 0) {
    if (cand) atomicAdd(counters + 16 + (blockIdx.x & (kCtrSlots - 1)), cand);  // spread, summed on the host
    if (kStats) {
      atomicAdd(counters + 2, (unsigned long long) st.nodes);
      atomicAdd(counters + 3, (unsigned long long) st.leaves);
      atomicAdd(counters + 4, (unsigned long long) st.top_steps);
      atomicAdd(counters + 5, (unsigned long long) st.lane_leaf);
      atomicAdd(counters + 6, (unsigned long long) (st.leaves ? 1 : 0));
      atomicMax(counters + 7, (unsigned long long) st.maxsp);
    }
  }
}

__global__ void k_pip_split(const uint2* __restrict__ packed, uint32_t n, uint32_t* __restrict__ out_eid,
                            int32_t* __restrict__ out_face) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 v = packed[i];
  out_eid[i] = v.x;
  out_face[i] = (int32_t) v.y;
}

// all points x all edges (RJB_MODE_BRUTE): pins the PIP arithmetic
__global__ void __launch_bounds__(256)
k_pip_brute(const longlong2* __restrict__ pts, uint32_t n_pts, MapView B, int query_map_id,
            uint32_t* __restrict__ out_eid, int32_t* __restrict__ out_face) {
  __shared__ Seg s_b[256];
  const uint32_t pi = blockIdx.x * 256 + threadIdx.x;
  const bool valid = pi < n_pts;
  long long px = 0, py = 0;
  if (valid) { px = pts[pi].x; py = pts[pi].y; }
  PipBest best;
  pip_init(best);
  for (uint32_t b0 = 0; b0 < B.n_edges; b0 += 256) {
    __syncthreads();
    uint32_t be = b0 + threadIdx.x;
    if (be < B.n_edges) s_b[threadIdx.x] = load_seg(B, be);
    __syncthreads();
    uint32_t nb = min(256u, B.n_edges - b0);
    if (valid)
      for (uint32_t k = 0; k < nb; k++) pip_update(best, query_map_id, px, py, s_b[k], b0 + k);
  }
  if (valid) {
    out_eid[pi] = best.eid;
    if (out_face) {
      int32_t face = RJB_EXTERIOR_FACE;
      if (best.eid != RJB_NO_HIT) {
        uint32_t c = B.edge_chain[best.eid];
        longlong2 a = B.pts[best.eid + c], b = B.pts[best.eid + c + 1];
        face = a.x < b.x ? B.right[c] : B.left[c];
      }
      out_face[pi] = face;
    }
  }
}

}  // namespace rjb
