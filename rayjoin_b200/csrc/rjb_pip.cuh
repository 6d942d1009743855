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

__global__ void __launch_bounds__(kLsiWarps * 32)
k_pip_bvh(const longlong2* __restrict__ pts, uint32_t n_pts, const uint32_t* __restrict__ order,
          MapView B, BvhView bvh, int query_map_id, uint32_t* __restrict__ out_eid,
          int32_t* __restrict__ out_face, unsigned long long* n_cand) {
  __shared__ int s_stack[kLsiWarps][kStackDepth];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stack = s_stack[warp];
  const uint32_t slot = (blockIdx.x * kLsiWarps + warp) * 32 + lane;
  const bool valid = slot < n_pts;
  uint32_t pi = 0;
  long long px = 0, py = 0;
  int qx = 1, qy_lo = 1;
  if (valid) {
    pi = order ? order[slot] : slot;
    longlong2 p = pts[pi];
    px = p.x;
    py = p.y;
    qx = quant(px);
    qy_lo = quant(py - 1);
  }
  PipBest best;
  pip_init(best);
  int y_hi = 0x7fffffff;  // quantised pruning bound, shrinks as best improves
  unsigned long long cand = 0;
  // lane wants a node box: x range contains px, box reaches up to py, and box
  // starts below the current best hit
  auto wants = [&](const int4& b) {
    return valid && b.x <= qx && qx <= b.z && b.w >= qy_lo && b.y <= y_hi;
  };
  if (bvh.n_leaves > 0 && __ballot_sync(0xffffffffu, wants(bvh.root_box)) != 0) {
    int sp = 0;
    int node = 0;
    while (true) {
      const int4 lb = __ldg(&bvh.node_box[2 * node]);
      const int4 rb = __ldg(&bvh.node_box[2 * node + 1]);
      const int2 ch = __ldg(&bvh.node_child[node]);
      // near child first: the one that starts lower in y (warp-uniform choice)
      const bool swap = rb.y < lb.y;
      const int4 b0 = swap ? rb : lb, b1 = swap ? lb : rb;
      const int c0 = swap ? ch.y : ch.x, c1 = swap ? ch.x : ch.y;
      int next = -1;
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const int4 bb = side ? b1 : b0;
        const int c = side ? c1 : c0;
        const bool h = wants(bb);  // re-evaluated: y_hi may have shrunk
        const unsigned m = __ballot_sync(0xffffffffu, h);
        if (m == 0) continue;
        if (c >= 0) {
          // far child goes to the stack and is re-tested when popped
          if (next < 0) next = c; else stack[sp++] = c;
          continue;
        }
        const uint2 rec = __ldg(&bvh.leaf_rec[~c]);
        const uint32_t first_eid = rec.x, cnt = rec.y >> 28, chain = rec.y & 0x0FFFFFFFu;
        const longlong2* bp = B.pts + (first_eid + chain);
        longlong2 p1 = __ldg(bp);
        for (uint32_t k = 0; k < cnt; k++) {
          const longlong2 p2 = __ldg(bp + k + 1);
          if (h) {
            const Seg e = {p1.x, p1.y, p2.x, p2.y};
            cand++;
            if (pip_update(best, query_map_id, px, py, e, first_eid + k))
              y_hi = pip_y_bound(best.y);
          }
          p1 = p2;
        }
      }
      if (next >= 0) { node = next; continue; }
      // pop until a node some lane still wants (its box is in the parent record,
      // so the test happens after loading; cheap because loads are broadcast)
      if (sp == 0) break;
      node = stack[--sp];
    }
  }
  if (valid) {
    out_eid[pi] = best.eid;
    if (out_face) {
      int32_t face = RJB_EXTERIOR_FACE;
      if (best.eid != RJB_NO_HIT) {
        // get_face_id, src/map/map.h:79-87
        uint32_t c = B.edge_chain[best.eid];
        longlong2 a = B.pts[best.eid + c], b = B.pts[best.eid + c + 1];
        face = a.x < b.x ? B.right[c] : B.left[c];
      }
      out_face[pi] = face;
    }
  }
  if (n_cand) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
    if (lane == 0 && cand) atomicAdd(n_cand, cand);
  }
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
