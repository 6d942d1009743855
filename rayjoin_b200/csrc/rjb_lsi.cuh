// LSI kernels: warp-cooperative BVH traversal (candidate generation), the dense
// exact pass (predicate + intersection point), and an all-pairs kernel.
//
// Replaces LSILBVH::Query (reference: src/app/lsi_lbvh.h:27-98 +
// deps/lbvh/lbvh/query.cuh:8-51, one thread per query edge with a 64-entry
// local-memory stack and three dependent AoS loads per node).
#pragma once
#include "rjb_exact.cuh"

namespace rjb {

constexpr int kLsiWarps = 4;     // warps per CTA (small CTAs: a slow warp holds few others)
constexpr int kStackDepth = 96;  // >= max LBVH depth (64 key bits + 32 index bits)

static __device__ __forceinline__ Seg load_seg(const MapView& m, uint32_t eid) {
  uint32_t p = eid + m.edge_chain[eid];
  longlong2 a = m.pts[p], b = m.pts[p + 1];
  Seg s = {a.x, a.y, b.x, b.y};
  return s;
}

// Warp-aggregated append of (query eid, base eid) to a queue: one atomic per warp.
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

static __device__ __forceinline__ int4 shfl_box(const int4& b, int src) {
  return make_int4(__shfl_sync(0xffffffffu, b.x, src), __shfl_sync(0xffffffffu, b.y, src),
                   __shfl_sync(0xffffffffu, b.z, src), __shfl_sync(0xffffffffu, b.w, src));
}

// union of the lanes' boxes (REDUX); lanes that do not take part pass the neutral box
static __device__ __forceinline__ int4 warp_union(const int4& b) {
  return make_int4(__reduce_min_sync(0xffffffffu, b.x), __reduce_min_sync(0xffffffffu, b.y),
                   __reduce_max_sync(0xffffffffu, b.z), __reduce_max_sync(0xffffffffu, b.w));
}

// statistics of one warp's traversal (option "stats")
struct TravStats {
  unsigned nodes, leaves, top_steps, lane_leaf, maxsp;
};

// A leaf that some lanes' boxes overlap is NOT opened by the traversing warp: the
// (query start point, leaf) pairs are appended to a queue (one atomic per warp)
// and resolved later by a dense kernel in which every thread is independent.
// Keeping the leaf's dependent loads and its divergent edge tests out of the
// warp-synchronous walk roughly halves the traversal time.
template <bool kStats>
static __device__ __forceinline__ void lsi_leaf(const MapView&, const BvhView&, int leaf, bool h,
                                                const Seg&, uint32_t qe, uint2* __restrict__ out,
                                                uint32_t cap, unsigned int* counter, int lane,
                                                TravStats& st) {
  if (kStats) st.leaves++;
  emit_pair(h, qe, (uint32_t) leaf, out, cap, counter, lane);
}

// Binary part of the traversal, below the 32-ary top tree: node records are read
// at a warp-uniform address (one transaction), every lane tests ITS box against
// both child boxes and __ballot_sync decides warp-uniformly where to go.
template <bool kStats>
static __device__ __forceinline__ void lsi_subtree(const MapView& B, const BvhView& bvh, int root,
                                                   int* stack, const int4& qb, const Seg& q,
                                                   uint32_t qe, uint2* __restrict__ out,
                                                   uint32_t cap, unsigned int* counter, int lane,
                                                   TravStats& st) {
  int sp = 0;
  int node = root;
  while (true) {
    const int4 lb = __ldg(&bvh.node_box[2 * node]);
    const int4 rb = __ldg(&bvh.node_box[2 * node + 1]);
    const int2 ch = __ldg(&bvh.node_child[node]);
    const bool hl = box_overlap(qb, lb), hr = box_overlap(qb, rb);
    const unsigned ml = __ballot_sync(0xffffffffu, hl);
    const unsigned mr = __ballot_sync(0xffffffffu, hr);
    if (kStats) st.nodes++;
    int next = -1;
    if (ml) {
      if (ch.x >= 0) next = ch.x;
      else {
        if (kStats) st.lane_leaf += __popc(ml);
        lsi_leaf<kStats>(B, bvh, ~ch.x, hl, q, qe, out, cap, counter, lane, st);
      }
    }
    if (mr) {
      if (ch.y >= 0) {
        if (next < 0) next = ch.y; else stack[sp++] = ch.y;
      } else {
        if (kStats) st.lane_leaf += __popc(mr);
        lsi_leaf<kStats>(B, bvh, ~ch.y, hr, q, qe, out, cap, counter, lane, st);
      }
    }
    if (kStats) st.maxsp = max(st.maxsp, (unsigned) sp);
    if (next >= 0) { node = next; continue; }
    if (sp == 0) break;
    node = stack[--sp];
  }
}

// One warp = 32 query edges.  The warp first resolves the top 15 levels of the
// tree in three steps of a 32-ary top tree (one LANE PER CHILD SLOT tests the
// warp's union box; loads are coalesced and the 21 KB of the first two levels
// stay L1-resident), then walks the remaining binary subtrees with per-lane
// boxes.  Output: candidate pairs whose exact integer boxes overlap.
//
// Query slots are POINT indices: lane p owns the edge (pts[p], pts[p+1]) unless p
// is the last point of a chain (one bit per point), so a tile is two coalesced
// 16-byte loads per lane plus one 32-bit word per warp -- no dependent
// eid -> chain -> point chain.  One tile per warp: the hardware CTA scheduler
// balances the load (a persistent, software-pipelined variant measured slower).
// With `order` (Morton-sorted queries) a slot is order[slot] instead.
struct QTile {
  longlong2 a, b;
  uint32_t p;
  bool valid;
};

static __device__ __forceinline__ QTile load_tile(const MapView& Q, const uint32_t* __restrict__ order,
                                                  uint32_t n_slots, uint32_t tile, int lane) {
  QTile t;
  t.a = make_longlong2(0, 0);
  t.b = make_longlong2(0, 0);
  const uint32_t slot = tile * 32 + lane;
  t.valid = slot < n_slots;
  t.p = 0;
  if (order) {
    if (t.valid) t.p = order[slot];
  } else {
    t.p = slot;
    if (tile * 32 < n_slots) {
      const uint32_t w = __ldg(&Q.last_bits[tile]);  // warp-uniform
      t.valid = t.valid && !((w >> lane) & 1u);
    }
  }
  if (t.valid) {
    t.a = __ldg(&Q.pts[t.p]);
    t.b = __ldg(&Q.pts[t.p + 1]);
  }
  return t;
}

// Occupancy pre-filter: streams every query edge once, keeps the start-point index
// of those whose (quantised) box touches an occupied cell of the base map's
// bitmap.  For a sparse base map (county boundaries: ~1 % of the cells) a few per
// cent of the query edges survive, and only those are traversed.
constexpr int kFilterTilesPerWarp = 16;  // a CTA of 8 warps covers 128 consecutive tiles

// keep-mask of one tile of 32 consecutive query points (bit = edge starting there
// touches an occupied cell)
static __device__ __forceinline__ unsigned filter_tile(const MapView& Q, const uint32_t* __restrict__ occ,
                                                       uint32_t tile, int lane) {
  // one 16-byte load per lane; the edge's second vertex is the next lane's point
  const uint32_t p = tile * 32 + lane;
  const bool in = p < Q.n_points;
  const longlong2 a = in ? __ldg(&Q.pts[p]) : make_longlong2(0, 0);
  // occupancy cell straight from the 47-bit coordinate: (v + 2^46) >> 35 equals
  // occ_cell(quant(v)); packed as (cy << kOccBits | cx)
  const int sh = kQuantShift + kOccShift;
  const uint32_t cx = (uint32_t) ((unsigned long long) (a.x + (1ll << 46)) >> sh) & (kOccDim - 1);
  const uint32_t cy = (uint32_t) ((unsigned long long) (a.y + (1ll << 46)) >> sh) & (kOccDim - 1);
  const uint32_t code = (cy << kOccBits) | cx;
  uint32_t code2 = __shfl_down_sync(0xffffffffu, code, 1);
  if (lane == 31 && p + 1 < Q.n_points) {
    const longlong2 b = __ldg(&Q.pts[p + 1]);
    code2 = ((uint32_t) ((unsigned long long) (b.y + (1ll << 46)) >> sh) & (kOccDim - 1)) << kOccBits |
            ((uint32_t) ((unsigned long long) (b.x + (1ll << 46)) >> sh) & (kOccDim - 1));
  }
  const uint32_t w = __ldg(&Q.last_bits[tile]);  // warp-uniform: bit set = no edge starts here
  const bool valid = in && !((w >> lane) & 1u);
  bool keep = false;
  if (valid) {
    if (code == code2) {  // the usual case: both vertices in one cell
      keep = (__ldg(&occ[code >> 5]) >> (code & 31)) & 1u;
    } else {
      const uint32_t x0 = min(code & (kOccDim - 1), code2 & (kOccDim - 1));
      const uint32_t x1 = max(code & (kOccDim - 1), code2 & (kOccDim - 1));
      const uint32_t y0 = min(code >> kOccBits, code2 >> kOccBits), y1 = max(code >> kOccBits, code2 >> kOccBits);
      for (uint32_t y = y0; y <= y1 && !keep; y++)
        for (uint32_t x = x0; x <= x1; x++) {
          const uint32_t bit = y * kOccDim + x;
          if ((__ldg(&occ[bit >> 5]) >> (bit & 31)) & 1u) { keep = true; break; }
        }
    }
  }
  return __ballot_sync(0xffffffffu, keep);
}

// A CTA filters 128 consecutive tiles (4096 points) and appends its survivors as ONE
// contiguous, map-ordered run (block scan of the tile counts, one atomic per CTA):
// the 32 survivors a traversal warp picks up are then neighbours on the map.  (With
// one atomic per warp the runs of concurrently running warps from all over the map
// interleave, and every traversal warp has to follow up to 32 separate clusters.)
__global__ void __launch_bounds__(256)
k_lsi_filter(MapView Q, const uint32_t* __restrict__ occ, uint32_t* __restrict__ survivors,
             unsigned int* counter) {
  constexpr int kTiles = 8 * kFilterTilesPerWarp;
  __shared__ unsigned s_mask[kTiles];
  __shared__ unsigned s_off[kTiles];
  __shared__ unsigned s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n_tiles = (Q.n_points + 31) / 32;
  const uint32_t tile0 = blockIdx.x * kTiles + warp * kFilterTilesPerWarp;
#pragma unroll 4
  for (int t = 0; t < kFilterTilesPerWarp; t++) {
    const uint32_t tile = tile0 + t;
    unsigned m = 0;
    if (tile < n_tiles) m = filter_tile(Q, occ, tile, lane);
    if (lane == 0) s_mask[warp * kFilterTilesPerWarp + t] = m;
  }
  __syncthreads();
  // exclusive scan of the 128 tile counts by warp 0 (4 per lane)
  if (warp == 0) {
    unsigned c[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { c[k] = __popc(s_mask[lane * 4 + k]); sum += c[k]; }
    unsigned inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    unsigned ex = inc - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) { s_off[lane * 4 + k] = ex; ex += c[k]; }
    if (lane == 31) s_base = inc ? atomicAdd(counter, inc) : 0u;
  }
  __syncthreads();
  const unsigned base = s_base;
  for (int t = 0; t < kFilterTilesPerWarp; t++) {
    const int i = warp * kFilterTilesPerWarp + t;
    const unsigned m = s_mask[i];
    if ((m >> lane) & 1u)
      survivors[base + s_off[i] + __popc(m & ((1u << lane) - 1))] = (blockIdx.x * kTiles + i) * 32 + lane;
  }
}

template <bool kStats>
__global__ void __launch_bounds__(kLsiWarps * 32)
k_lsi_bvh(MapView Q, MapView B, BvhView bvh, const uint32_t* __restrict__ order, uint32_t n_slots,
          const unsigned int* __restrict__ n_slots_dev,
          uint2* __restrict__ out, uint32_t cap, unsigned int* counter,
          unsigned long long* stats) {
  __shared__ int s_stack[kLsiWarps][kStackDepth];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stack = s_stack[warp];
  if (n_slots_dev) n_slots = *n_slots_dev;  // survivor count of the pre-filter
  const uint32_t n_tiles = (n_slots + 31) / 32;
  const uint32_t tile = blockIdx.x * kLsiWarps + warp;
  TravStats st = {0, 0, 0, 0, 0};
  const int4 kNeutral = empty_box();
  const int4 kEmpty = empty_box();
  if (tile >= n_tiles) return;
  const QTile cur = load_tile(Q, order, n_slots, tile, lane);
  const bool valid = cur.valid;
  const uint32_t qe = cur.p;
  const Seg q = {cur.a.x, cur.a.y, cur.b.x, cur.b.y};
  int4 qb = empty_box();  // empty box for idle lanes
  if (valid)
    qb = make_int4(quant(min(q.x1, q.x2)), quant(min(q.y1, q.y2)),
                   quant(max(q.x1, q.x2)), quant(max(q.y1, q.y2)));
  const int4 ub = qb;
  const int4 U = warp_union(ub);
  if (bvh.n_leaves > 0 && box_overlap(U, bvh.root_box)) {
    // level 0: the 32 nodes at depth 5.  The union box only preselects slots; each
    // preselected slot is confirmed with the lanes' own boxes (ballot) before the
    // warp descends, and the union box is re-tightened to the confirming lanes, so
    // a warp whose edges form several far-apart clusters follows each cluster
    // separately instead of sweeping everything under a huge union box.
    const int4 b0 = __ldg(&bvh.top_box[kTopOff0 + lane]);
    const int c0 = __ldg(&bvh.top_code[kTopOff0 + lane]);
    unsigned m0 = __ballot_sync(0xffffffffu, box_overlap(U, b0));
    if (kStats) st.top_steps++;
    while (m0) {
      const int g = __ffs(m0) - 1;
      m0 &= m0 - 1;
      const int4 sb0 = shfl_box(b0, g);
      const bool h0 = box_overlap(qb, sb0);
      if (__ballot_sync(0xffffffffu, h0) == 0) continue;
      const int code0 = __shfl_sync(0xffffffffu, c0, g);
      if (code0 < 0) {
        lsi_leaf<kStats>(B, bvh, ~code0, h0, q, qe, out, cap, counter, lane, st);
        continue;
      }
      const int4 U1 = warp_union(h0 ? qb : kNeutral);
      // level 1: the 32 depth-10 nodes below slot g
      const int4 b1 = __ldg(&bvh.top_box[kTopOff1 + g * 32 + lane]);
      const int c1 = __ldg(&bvh.top_code[kTopOff1 + g * 32 + lane]);
      unsigned m1 = __ballot_sync(0xffffffffu, box_overlap(U1, b1));
      if (kStats) st.top_steps++;
      while (m1) {
        const int h = __ffs(m1) - 1;
        m1 &= m1 - 1;
        const int4 sb1 = shfl_box(b1, h);
        const bool h1 = h0 && box_overlap(qb, sb1);
        if (__ballot_sync(0xffffffffu, h1) == 0) continue;
        const int code1 = __shfl_sync(0xffffffffu, c1, h);
        if (code1 < 0) {
          lsi_leaf<kStats>(B, bvh, ~code1, h1, q, qe, out, cap, counter, lane, st);
          continue;
        }
        const int4 U2 = warp_union(h1 ? qb : kNeutral);
        // level 2: the 32 depth-15 nodes below slot (g, h)
        const int4 b2 = __ldg(&bvh.top_box[kTopOff2 + (g * 32 + h) * 32 + lane]);
        const int c2 = __ldg(&bvh.top_code[kTopOff2 + (g * 32 + h) * 32 + lane]);
        unsigned m2 = __ballot_sync(0xffffffffu, box_overlap(U2, b2));
        if (kStats) st.top_steps++;
        while (m2) {
          const int i = __ffs(m2) - 1;
          m2 &= m2 - 1;
          const int4 sb2 = shfl_box(b2, i);
          const bool h2 = h1 && box_overlap(qb, sb2);
          if (__ballot_sync(0xffffffffu, h2) == 0) continue;
          const int code2 = __shfl_sync(0xffffffffu, c2, i);
          if (code2 < 0) {
            lsi_leaf<kStats>(B, bvh, ~code2, h2, q, qe, out, cap, counter, lane, st);
            continue;
          }
          if (bvh.top_levels < 4) {
            lsi_subtree<kStats>(B, bvh, code2, stack, h2 ? qb : kEmpty, q, qe, out, cap, counter,
                                lane, st);
            continue;
          }
          // level 3 (big trees only): the 32 depth-20 nodes below slot (g, h, i)
          const int4 U3 = warp_union(h2 ? qb : kNeutral);
          const uint32_t s3 = kTopOff3 + ((uint32_t) ((g * 32 + h) * 32 + i)) * 32 + lane;
          const int4 b3 = __ldg(&bvh.top_box[s3]);
          const int c3 = __ldg(&bvh.top_code[s3]);
          unsigned m3 = __ballot_sync(0xffffffffu, box_overlap(U3, b3));
          if (kStats) st.top_steps++;
          while (m3) {
            const int j = __ffs(m3) - 1;
            m3 &= m3 - 1;
            const int4 sb3 = shfl_box(b3, j);
            const bool h3 = h2 && box_overlap(qb, sb3);
            if (__ballot_sync(0xffffffffu, h3) == 0) continue;
            const int code3 = __shfl_sync(0xffffffffu, c3, j);
            if (code3 < 0)
              lsi_leaf<kStats>(B, bvh, ~code3, h3, q, qe, out, cap, counter, lane, st);
            else
              lsi_subtree<kStats>(B, bvh, code3, stack, h3 ? qb : kEmpty, q, qe, out, cap, counter,
                                  lane, st);
          }
        }
      }
    }
  }
  if (kStats && lane == 0) {
    atomicAdd(stats + 0, (unsigned long long) st.nodes);
    atomicAdd(stats + 1, (unsigned long long) st.leaves);
    atomicAdd(stats + 2, (unsigned long long) st.top_steps);
    atomicAdd(stats + 3, (unsigned long long) st.lane_leaf);
    atomicAdd(stats + 4, (unsigned long long) (st.leaves ? 1 : 0));
    atomicMax(stats + 5, (unsigned long long) st.maxsp);
  }
}

// chain of point p: last c with row_index[c] <= p
static __device__ __forceinline__ uint32_t chain_of_point(const MapView& m, uint32_t p) {
  uint32_t lo = 0, hi = m.n_chains;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(&m.row_index[mid]) <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// Exact pass 1, dense over the (query start point, leaf) pairs of the traversal:
// each thread tests its query edge against the <= 8 consecutive base edges of the
// leaf (one contiguous run of points): exact integer box test, then
// intersect_test.  Hits are compacted into the result queue (one atomic per warp
// and edge slot) as start-point index pairs.  The pair count is read on the
// device: no host round trip between the kernels.
__global__ void __launch_bounds__(256)
k_lsi_exact(MapView Q, MapView B, const uint2* __restrict__ pairs, const uint2* __restrict__ leaf_rec,
            const unsigned int* __restrict__ n_pairs_dev, uint32_t pair_cap,
            rjb_xsect* __restrict__ out, uint32_t cap, unsigned int* counter,
            unsigned long long* n_cand) {
  const uint32_t n = min(*n_pairs_dev, pair_cap);
  const int lane = threadIdx.x & 31;
  unsigned long long cand = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i - lane < n;
       i += gridDim.x * blockDim.x) {
    const bool act = i < n;
    uint32_t pq = 0, pb0 = 0, cnt = 0;
    Seg e1 = {0, 0, 0, 0};
    longlong2 p1 = make_longlong2(0, 0);
    if (act) {
      const uint2 pr = pairs[i];
      pq = pr.x;
      const uint2 rec = __ldg(&leaf_rec[pr.y]);
      cnt = rec.y >> 28;
      pb0 = rec.x + (rec.y & 0x0FFFFFFFu);
      const longlong2 a = __ldg(&Q.pts[pq]), b = __ldg(&Q.pts[pq + 1]);
      e1 = {a.x, a.y, b.x, b.y};
      p1 = __ldg(&B.pts[pb0]);
    }
    const uint32_t cmax = __reduce_max_sync(0xffffffffu, cnt);
    for (uint32_t k = 0; k < cmax; k++) {
      bool found = false;
      if (k < cnt) {
        const longlong2 p2 = __ldg(&B.pts[pb0 + k + 1]);
        const Seg e2 = {p1.x, p1.y, p2.x, p2.y};
        if (seg_boxes_overlap(e1, e2)) {
          cand++;
          found = lsi_intersect(e1, e2);
        }
        p1 = p2;
      }
      const unsigned m = __ballot_sync(0xffffffffu, found);
      if (m == 0) continue;
      unsigned base = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(counter, (unsigned) __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (found) {
        const unsigned pos = base + __popc(m & ((1u << lane) - 1));
        if (pos < cap) {
          out[pos].eid[0] = pq;  // point indices for now; pass 2 turns them into eids
          out[pos].eid[1] = pb0 + k;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
  if (lane == 0 && cand) atomicAdd(n_cand, cand);
}

// Exact pass 2, dense over the hits: rational intersection point and the final
// edge ids -> rjb_xsect (reference computes it inside the traversal callback,
// lsi_lbvh.h:69-78; its RT backend has the same post-pass, src/app/lsi_rt.h:66-112).
__global__ void __launch_bounds__(128)
k_lsi_points(MapView Q, MapView B, int query_map_id, const unsigned int* __restrict__ counter,
             uint32_t cap, rjb_xsect* __restrict__ out) {
  // two threads per hit: the even lane computes x, the odd lane y (each axis is a
  // long dependent chain: gcd + division), then the pair assembles the record
  const uint32_t n = min(*counter, cap);
  const uint32_t stride = (gridDim.x * blockDim.x) >> 1;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 1; i - ((threadIdx.x & 31) >> 1) < n;
       i += stride) {
    const int axis = threadIdx.x & 1;
    long long v = 0;
    uint32_t pq = 0, pb = 0;
    if (i < n) {
      pq = out[i].eid[0];
      pb = out[i].eid[1];
      const longlong2 a = __ldg(&Q.pts[pq]), b = __ldg(&Q.pts[pq + 1]);
      const longlong2 c = __ldg(&B.pts[pb]), d = __ldg(&B.pts[pb + 1]);
      const Seg e1 = {a.x, a.y, b.x, b.y}, e2 = {c.x, c.y, d.x, d.y};
      v = lsi_point_axis(e1, e2, axis);
    }
    // every lane of the warp takes part in the exchange (the loop bound is warp-uniform)
    const long long other = __shfl_xor_sync(0xffffffffu, v, 1);
    // the point-index fields are overwritten below: both lanes must have read them
    __syncwarp();
    if (i < n && axis == 0) {
      const uint32_t eq = pq - chain_of_point(Q, pq), eb = pb - chain_of_point(B, pb);
      rjb_xsect r;
      r.x = v;
      r.y = other;
      r.eid[0] = query_map_id == 0 ? eq : eb;
      r.eid[1] = query_map_id == 0 ? eb : eq;
      r.mid_point_polygon_id = RJB_DONTKNOW;
      r._pad = 0;
      out[i] = r;
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

// Intersection points of already verified pairs -> rjb_xsect records (grid and
// brute modes).  The pair count is read from the device counter.
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
