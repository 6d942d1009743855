// LSI kernels: warp-cooperative BVH traversal (candidate generation), the dense
// exact pass (predicate + intersection point), and an all-pairs kernel.
//
// Replaces LSILBVH::Query (reference: src/app/lsi_lbvh.h:27-98 +
// deps/lbvh/lbvh/query.cuh:8-51, one thread per query edge with a 64-entry
// local-memory stack and three dependent AoS loads per node).
#pragma once
#include "rjb_exact.cuh"

namespace rjb {

// Programmatic dependent launch: the kernels of a query are launched with the "programmatic
// stream serialization" attribute, so the next kernel's CTAs may be scheduled (and run their
// prologue) while the previous kernel drains; pdl_wait() blocks until the previous grid has
// completed and its writes are visible, pdl_launch_dependents() lets the next grid start early.
// MEASURED (B200, bench workload): slower, 0.155 against 0.124 ms per query -- the early CTAs of
// the next kernel take the registers the running one still needs -- and even the bare
// griddepcontrol instructions in normally launched kernels cost (k_lsi_cells 47 -> 67 us), so
// the whole mechanism is compiled in only with -DRJB_PDL (option lsi_pdl then switches it on).
// (Under PDL the counts the previous kernel left must be read with volatile loads -- a const
// __restrict__ load may be hoisted above the wait, which lost pairs -- and those loads are not
// free: every warp of the grid then goes to ONE L2 sector, k_lsi_cells 47 -> 67 us.)
#ifdef RJB_PDL
static __device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
static __device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
static __device__ __forceinline__ unsigned int load_count(const unsigned int* p) { return *(volatile const unsigned int*) p; }
#else
static __device__ __forceinline__ void pdl_wait() {}
static __device__ __forceinline__ void pdl_launch_dependents() {}
static __device__ __forceinline__ unsigned int load_count(const unsigned int* p) { return *p; }
#endif

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
// (query start point, leaf) pairs are resolved later by a dense kernel in which every
// thread is independent.  Keeping the leaf's dependent loads and its divergent edge
// tests out of the warp-synchronous walk roughly halves the traversal time.
//
// The pairs are staged in a per-warp shared-memory buffer and appended to the global
// queue kEmitFlush at a time: ONE atomic per flush instead of one per leaf, and its
// round trip to L2 (the walk cannot proceed past it) is paid once per warp, not once
// per leaf; the queue is written in full 256-byte runs.
constexpr int kEmitFlush = 64;              // flush at >= this many staged pairs
constexpr int kEmitBuf = kEmitFlush + 32;   // one more leaf always fits

struct Emit {
  uint2* buf;              // this warp's staging area (shared memory)
  unsigned n;              // staged pairs (warp-uniform)
  uint2* __restrict__ out;
  uint32_t cap;
  unsigned int* counter;
  // non-null: the staged pairs are {query, leaf id} and leave in the DIRECT format (kDirectShift)
  const uint2* __restrict__ conv_leaf_rec;
};

static __device__ __forceinline__ void emit_flush(Emit& E, int lane) {
  if (E.n == 0) return;
  unsigned base = 0;
  if (lane == 0) base = atomicAdd(E.counter, E.n);
  base = __shfl_sync(0xffffffffu, base, 0);
  __syncwarp();
  for (unsigned t = lane; t < E.n; t += 32) {
    uint2 v = E.buf[t];
    if (E.conv_leaf_rec) {
      const uint2 rec = __ldg(&E.conv_leaf_rec[v.y]);
      v = make_uint2(v.x | (((rec.y >> 28) - 1) << kDirectShift), rec.x + (rec.y & 0x0FFFFFFFu));
    }
    if (base + t < E.cap) E.out[base + t] = v;
  }
  __syncwarp();
  E.n = 0;
}

template <bool kStats>
static __device__ __forceinline__ void lsi_leaf(int leaf, bool h, uint32_t qe, Emit& E, int lane,
                                                TravStats& st) {
  if (kStats) st.leaves++;
  const unsigned m = __ballot_sync(0xffffffffu, h);
  if (h) E.buf[E.n + __popc(m & ((1u << lane) - 1))] = make_uint2(qe, (uint32_t) leaf);
  E.n += __popc(m);
  if (E.n >= kEmitFlush) emit_flush(E, lane);
}

// Binary part of the traversal, below the 32-ary top tree: node records are read
// at a warp-uniform address (one transaction), every lane tests ITS box against
// both child boxes and __ballot_sync decides warp-uniformly where to go.
template <bool kStats>
static __device__ __forceinline__ void lsi_subtree(const BvhView& bvh, int root, int* stack,
                                                   const int4& qb, uint32_t qe, Emit& E, int lane,
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
        lsi_leaf<kStats>(~ch.x, hl, qe, E, lane, st);
      }
    }
    if (mr) {
      if (ch.y >= 0) {
        if (next < 0) next = ch.y; else stack[sp++] = ch.y;
      } else {
        if (kStats) st.lane_leaf += __popc(mr);
        lsi_leaf<kStats>(~ch.y, hr, qe, E, lane, st);
      }
    }
    if (kStats) st.maxsp = max(st.maxsp, (unsigned) sp);
    if (next >= 0) { node = next; continue; }
    if (sp == 0) break;
    node = stack[--sp];
  }
}

// One warp = 32 query edges.  The warp first resolves the top 15 (20 for trees of
// >= 2^18 leaves) levels of the tree in three (four) steps of a 32-ary top tree (one
// LANE PER CHILD SLOT tests the warp's union box; loads are coalesced and the 21 KB
// of the first two levels stay L1-resident), then walks the remaining binary subtrees
// with per-lane boxes.  Output: (query, leaf) pairs whose quantised boxes overlap.
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

// (p_lo: plain slots only -- start points below it are outside the query window)
static __device__ __forceinline__ QTile load_tile(const MapView& Q, const uint32_t* __restrict__ order,
                                                  uint32_t n_slots, uint32_t tile, int lane,
                                                  uint32_t spw = 32, uint32_t p_lo = 0) {
  QTile t;
  t.a = make_longlong2(0, 0);
  t.b = make_longlong2(0, 0);
  // spw < 32 (list slots only): a warp takes fewer queries, the other lanes idle
  const uint32_t slot = tile * spw + lane;
  t.valid = slot < n_slots && (uint32_t) lane < spw;
  t.p = 0;
  if (order) {
    if (t.valid) t.p = order[slot];
  } else {
    t.p = slot;
    if (tile * 32 < n_slots) {
      const uint32_t w = __ldg(&Q.last_bits[tile]);  // warp-uniform
      t.valid = t.valid && !((w >> lane) & 1u) && slot >= p_lo;
    }
  }
  if (t.valid) {
    t.a = __ldg(&Q.pts[t.p]);
    t.b = __ldg(&Q.pts[t.p + 1]);
  }
  return t;
}

// Occupancy pre-filter: streams the query map once and keeps the start-point index of
// the edges whose (quantised) box touches an occupied cell of the base map's bitmap.
// For a sparse base map (county boundaries: ~1 % of the cells) a few per cent of the
// query edges survive, and only those are traversed.
//
// The kernel reads the 4-byte occupancy DESCRIPTOR of every edge (edge_desc_of, written
// once by the load kernel), not two 16-byte vertices: a quarter of the traffic, and the
// decision is one bitmap look-up per edge -- in `occ` for an edge inside one cell, in the
// 2 x 2-dilated `occ2` for an edge that crosses into a neighbour cell (conservative).
// A warp owns 512 consecutive points and works on them 32 at a time, lane = point:
// consecutive points lie in the same or neighbouring cells, so the 32 look-ups of one
// instruction fall into one or two 32-byte sectors.  (With 16 consecutive points per
// LANE every look-up instruction touched ~30 sectors and the kernel was bound by the L1
// tag stage at 88 %.)  All 16 descriptor loads and then all 16 look-ups of a thread are
// independent and in flight together.
constexpr int kFilterThreads = 256;
constexpr int kFilterPerThread = 16;
constexpr int kFilterCtaPoints = kFilterThreads * kFilterPerThread;

// edge longer than a cell (rare): every cell of its box, a bitmap word at a time.  A box of
// more than 4096 words is kept without looking (the tree walk deals with it): the loop of
// one thread stays bounded whatever the map contains.
#ifndef RJB_OCC_RECT_MAX
#define RJB_OCC_RECT_MAX 4096
#endif
constexpr uint32_t kOccRectMaxWords = RJB_OCC_RECT_MAX;
static __device__ __noinline__ bool occ_rect(const MapView& Q, const uint32_t* __restrict__ occ, uint32_t p,
                                             bool* within3) {
  const longlong2 a = __ldg(&Q.pts[p]), b = __ldg(&Q.pts[p + 1]);
  const uint32_t c1 = occ_code(a.x, a.y), c2 = occ_code(b.x, b.y);
  const uint32_t x0 = min(c1 & (kOccDim - 1), c2 & (kOccDim - 1));
  const uint32_t x1 = max(c1 & (kOccDim - 1), c2 & (kOccDim - 1));
  const uint32_t y0 = min(c1 >> kOccBits, c2 >> kOccBits), y1 = max(c1 >> kOccBits, c2 >> kOccBits);
  *within3 = x1 - x0 <= 2 && y1 - y0 <= 2;
  const uint32_t w0 = x0 >> 5, w1 = x1 >> 5;
  if ((w1 - w0 + 1) * (y1 - y0 + 1) > kOccRectMaxWords) return true;
  for (uint32_t y = y0; y <= y1; y++)
    for (uint32_t w = w0; w <= w1; w++) {
      const uint32_t lo = max(x0, 32 * w), hi = min(x1, 32 * w + 31);
      const uint32_t m = (hi - lo == 31) ? 0xFFFFFFFFu : (((1u << (hi - lo + 1)) - 1) << (lo & 31));
      if (__ldg(&occ[y * (kOccDim / 32) + w]) & m) return true;
    }
  return false;
}

// A CTA filters 4096 consecutive points and appends its survivors as ONE contiguous,
// map-ordered run (positions from the ballots, scan over the 8 warp totals, one atomic per
// CTA): the 32
// survivors a traversal warp picks up are then neighbours on the map.  (With one
// atomic per warp the runs of concurrently running warps from all over the map
// interleave, and every traversal warp has to follow up to 32 separate clusters.)
__global__ void __launch_bounds__(kFilterThreads)
k_lsi_filter(MapView Q, uint32_t p_lo, uint32_t p_hi, const uint32_t* __restrict__ occ,
             uint32_t* __restrict__ survivors, unsigned int* counter, uint32_t* __restrict__ long_list,
             unsigned int* long_counter) {
  constexpr int kWarps = kFilterThreads / 32;
  __shared__ unsigned s_wsum[kWarps];
  __shared__ unsigned s_base;
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // point of (round e, lane) = w0 + 32 * e + lane; only start points in the query window
  // [p_lo, p_hi) are looked at (a shard of the query map; the whole map by default)
  const uint32_t w0 = (p_lo & ~31u) + blockIdx.x * kFilterCtaPoints + warp * (32 * kFilterPerThread);
  // stage 1: the descriptors (edge_desc is padded with "no edge" up to a multiple of 16;
  // beyond that the index is clamped and the value ignored)
  uint32_t d[kFilterPerThread];
#pragma unroll
  for (int e = 0; e < kFilterPerThread; e++) {
    const uint32_t p = w0 + 32 * e + lane;
    d[e] = __ldg(&Q.edge_desc[min(p, Q.n_points)]);
    if (p >= p_hi || p < p_lo) d[e] = kDescNone << 24;
  }
  // stage 2: one look-up per edge; class bit 0 selects occ2 (kDescNone / kDescBig read a
  // valid word too and ignore it)
  uint32_t w[kFilterPerThread];
#pragma unroll
  for (int e = 0; e < kFilterPerThread; e++)
    w[e] = __ldg(&occ[((d[e] & 0xFFFFFFu) >> 5) + ((d[e] >> 24) & 1u) * kOccWords]);
  // stage 3: keep masks per round (bit = lane), positions from the ballots alone
  unsigned m[kFilterPerThread];
  unsigned cnt = 0;
#pragma unroll
  for (int e = 0; e < kFilterPerThread; e++) {
    const uint32_t cls = d[e] >> 24;
    bool keep = cls < kDescBig && ((w[e] >> (d[e] & 31u)) & 1u);
    if (__any_sync(0xffffffffu, cls == kDescBig)) {
      bool within3 = false;
      bool big = cls == kDescBig && occ_rect(Q, occ, w0 + 32 * e + lane, &within3);
      if (long_list) {
        // edges whose box exceeds 3 x 3 cells go to a list of their own: the tree walk
        // handles them, the cell directory (k_lsi_cells) everything else
        if (big && within3) {
          keep = true;
          big = false;
        }
        const unsigned mb = __ballot_sync(0xffffffffu, big);
        if (mb) {
          unsigned base = 0;
          const int leader = __ffs(mb) - 1;
          if (lane == leader) base = atomicAdd(long_counter, (unsigned) __popc(mb));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (big) long_list[base + __popc(mb & ((1u << lane) - 1))] = w0 + 32 * e + lane;
        }
      } else if (big) {
        keep = true;
      }
    }
    m[e] = __ballot_sync(0xffffffffu, keep);
    cnt += __popc(m[e]);
  }
  if (lane == 0) s_wsum[warp] = cnt;
  __syncthreads();
  if (warp == 0) {
    const unsigned sum = lane < kWarps ? s_wsum[lane] : 0u;
    unsigned winc = sum;
#pragma unroll
    for (int o = 1; o < kWarps; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += v;
    }
    if (lane < kWarps) s_wsum[lane] = winc - sum;
    if (lane == kWarps - 1) s_base = winc ? atomicAdd(counter, winc) : 0u;
  }
  __syncthreads();
  unsigned pos = s_base + s_wsum[warp];
  const unsigned lt = (1u << lane) - 1;
#pragma unroll
  for (int e = 0; e < kFilterPerThread; e++) {
    if ((m[e] >> lane) & 1u) survivors[pos + __popc(m[e] & lt)] = w0 + 32 * e + lane;
    pos += __popc(m[e]);
  }
}

// Two-level variant of the filter (option lsi_tile_filter, default): a warp first decides
// kTfRounds x 32 TILES of kTileT edges each with one look-up per tile (lane = tile: tile_desc_of,
// bitmaps dilated to the size of the tile's box), then reads the edge descriptors of the live
// tiles only: the live tiles of the warp are compacted into a list, and every look-up
// instruction of stage 2 serves 32 / kTileT of them (lane = point of a tile).  What makes it
// faster than the one-level filter is not the smaller volume (12 MB instead of 36 MB: neither
// is near the bandwidth) but the SHAPE: all loads of a level are issued together, so a warp pays
// four round trips to memory for 2048 edges, and the whole grid is one resident wave.  Survivors
// leave as one map-ordered run per CTA, like k_lsi_filter.  (History: tiles of 32 edges: 72 %
// live, no gain; tiles of 8, one 32-tile group per warp and round trip: no gain either, 35 us.)
// (measured, filter stage incl. its launch: 8 warps x 4 / 8 / 12 / 16 groups: 28.7 / 25.4 / 27.4 / 27.2 us;
// one group per warp, one warp-wide memory round trip per level and group: 35.6 us = the one-level filter)
#ifndef RJB_TF_ROUNDS
#define RJB_TF_ROUNDS 8
#endif
#ifndef RJB_TF_WARPS
#define RJB_TF_WARPS 8
#endif
constexpr int kTfWarps = RJB_TF_WARPS;
constexpr int kTfRounds = RJB_TF_ROUNDS;   // 32-tile groups per warp, decided TOGETHER: every level of the
                                           // chain tile_desc -> bitmap -> edge_desc -> bitmap is one round
                                           // trip to memory per warp, whatever the number of groups
constexpr int kTfCtaTiles = kTfWarps * kTfRounds * 32;
constexpr int kTfGroup = 32 / kTileT;      // tiles per stage-2 instruction
constexpr int kTfMaxSlots = kTfRounds * 32 / kTfGroup;  // stage-2 rounds of a warp at most (every tile live)
constexpr int kTfBatch = 8;                // stage-2 rounds in flight together

__global__ void __launch_bounds__(kTfWarps * 32)
k_lsi_filter_tiles(MapView Q, uint32_t p_lo, uint32_t p_hi, const uint32_t* __restrict__ occ,
                   uint32_t* __restrict__ survivors, unsigned int* counter, uint32_t* __restrict__ long_list,
                   unsigned int* long_counter) {
  __shared__ unsigned s_mask[kTfWarps][kTfMaxSlots];      // keep mask (bit = lane) per stage-2 round
  __shared__ uint32_t s_live[kTfWarps][kTfRounds * 32];   // the warp's live tiles, compacted, in map order
  __shared__ unsigned s_wsum[kTfWarps];
  __shared__ unsigned s_base;
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // tiles of the query window: start point p lies in tile (p + 1) / kTileT
  const uint32_t t_last = p_hi / kTileT;  // tile of the last start point p_hi - 1
  const uint32_t t0 = p_lo / kTileT + (blockIdx.x * kTfWarps + warp) * (kTfRounds * 32);
  const int sub = lane / kTileT, off = lane % kTileT;
  // stage 1: lane = tile, kTfRounds groups of 32 tiles; all descriptor loads, then all look-ups
  uint32_t td[kTfRounds];
#pragma unroll
  for (int r = 0; r < kTfRounds; r++) {
    const uint32_t t = t0 + r * 32 + lane;
    td[r] = t <= t_last ? __ldg(&Q.tile_desc[t]) : (kTileNone << 24);
  }
  uint32_t tw[kTfRounds];
#pragma unroll
  for (int r = 0; r < kTfRounds; r++) {
    const uint32_t cls = td[r] >> 24;
    const uint32_t code = td[r] & 0xFFFFFFu;
    tw[r] = (cls >= 1 && cls < kTileBig) ? __ldg(&occ[(cls - 1) * kOccWords + (code >> 5)]) : 0u;
  }
  int n_live = 0;
#pragma unroll
  for (int r = 0; r < kTfRounds; r++) {
    const uint32_t cls = td[r] >> 24;
    const bool live = cls == kTileBig || (cls >= 1 && cls < kTileBig && ((tw[r] >> (td[r] & 31u)) & 1u));
    const unsigned m = __ballot_sync(0xffffffffu, live);
    if (live) s_live[warp][n_live + __popc(m & ((1u << lane) - 1))] = t0 + r * 32 + lane;
    n_live += __popc(m);
  }
  __syncwarp();
  // stage 2: lane = point of one of kTfGroup live tiles per round, kTfBatch rounds in flight
  // together and only as many rounds as there are live tiles (warp-uniform trip count)
  const int n_slots = (n_live + kTfGroup - 1) / kTfGroup;
  unsigned cnt = 0;
  for (int k0 = 0; k0 < n_slots; k0 += kTfBatch) {
    uint32_t d[kTfBatch], pp[kTfBatch];
#pragma unroll
    for (int j = 0; j < kTfBatch; j++) {
      const int idx = (k0 + j) * kTfGroup + sub;
      const bool have = idx < n_live;
      const uint32_t p = (have ? s_live[warp][idx] : 0u) * kTileT - 1 + off;  // (tile 0, point 0: wraps, rejected)
      pp[j] = p;
      const bool in = have && p >= p_lo && p < p_hi;
      d[j] = in ? __ldg(&Q.edge_desc[p]) : (kDescNone << 24);
    }
    uint32_t w[kTfBatch];
#pragma unroll
    for (int j = 0; j < kTfBatch; j++)
      w[j] = __ldg(&occ[((d[j] & 0xFFFFFFu) >> 5) + ((d[j] >> 24) & 1u) * kOccWords]);
#pragma unroll
    for (int j = 0; j < kTfBatch; j++) {
      const uint32_t cls = d[j] >> 24;
      bool keep = cls < kDescBig && ((w[j] >> (d[j] & 31u)) & 1u);
      if (__any_sync(0xffffffffu, cls == kDescBig)) {
        bool within3 = false;
        bool big = cls == kDescBig && occ_rect(Q, occ, pp[j], &within3);
        if (long_list) {
          if (big && within3) {
            keep = true;
            big = false;
          }
          const unsigned mb = __ballot_sync(0xffffffffu, big);
          if (mb) {
            unsigned base = 0;
            const int leader = __ffs(mb) - 1;
            if (lane == leader) base = atomicAdd(long_counter, (unsigned) __popc(mb));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (big) long_list[base + __popc(mb & ((1u << lane) - 1))] = pp[j];
          }
        } else if (big) {
          keep = true;
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (k0 + j < n_slots) {
        if (lane == 0) s_mask[warp][k0 + j] = m;
        cnt += __popc(m);
      }
    }
  }
  if (lane == 0) s_wsum[warp] = cnt;
  __syncthreads();
  if (warp == 0) {
    const unsigned sum = lane < kTfWarps ? s_wsum[lane] : 0u;
    unsigned winc = sum;
#pragma unroll
    for (int o = 1; o < kTfWarps; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += v;
    }
    if (lane < kTfWarps) s_wsum[lane] = winc - sum;
    if (lane == kTfWarps - 1) s_base = winc ? atomicAdd(counter, winc) : 0u;
  }
  __syncthreads();
  unsigned pos = s_base + s_wsum[warp];
  const unsigned lt = (1u << lane) - 1;
  for (int k = 0; k < n_slots; k++) {
    const unsigned m = s_mask[warp][k];
    if ((m >> lane) & 1u) survivors[pos + __popc(m & lt)] = s_live[warp][k * kTfGroup + sub] * kTileT - 1 + off;
    pos += __popc(m);
  }
}

// Candidate generation through the cell directory (sparse base maps, short query edges:
// the survivors of the occupancy filter whose box lies within 3 x 3 cells).  A warp owns 32
// query edges.  Every lane looks up the cells of ITS edge -- {bitmap word, rank} in one load,
// then the list bounds: two rounds of independent loads for the 2 x 2 cells at the min corner,
// two more only if some edge spans three cells -- and the warp then works through the
// concatenation of all the lists with one lane per (query, item), 128 items at a time: the
// owner of an item is found by binary search over the prefix sums, and the 16-byte item record
// (cell_item_of: the leaf's box clipped to the cell, its first point and edge count) decides
// the pair without touching the leaf.  Four dependent rounds of loads per warp in all
// (survivor -> vertices -> directory -> bounds -> items) against eight for the tree walk,
// independent of how unevenly the lists are distributed over the lanes (a lane-per-query loop
// waits for the longest list: 8.1 tests against a mean of 3.4).
// A leaf that is listed in several cells of the query's box is reported once: in the cell
// that holds the min corner of the intersection of the two cell boxes.  The pairs are
// exactly the ones the tree walk emits (quantised boxes overlap); they leave in the DIRECT
// format, so the exact pass needs no leaf record either.
struct CellWork {
  int4 qb[32];           // query boxes of the lanes
  uint32_t p[32];        // query start points
  uint32_t beg[32][9];   // list begin per cell (cell k = (cx0 + k % 3, cy0 + k / 3))
  uint32_t cum[32][9];   // inclusive prefix of the list lengths over the lane's cells
  uint32_t pre[33];      // exclusive prefix of the lanes' totals
};

constexpr int kCellBatch = 4;  // items per lane and batch

// the directory word of cell (cx, cy): {bitmap word, rank}, first round of loads
static __device__ __forceinline__ uint2 cell_dir(const BvhView& bvh, bool in, int cx, int cy) {
  const uint32_t bit = in ? (uint32_t) cy * kOccDim + cx : 0u;
  return __ldg(&bvh.occ_dir[bit >> 5]);
}

// list bounds of the cell if it is occupied, else an empty range: second round
static __device__ __forceinline__ void cell_bounds(const BvhView& bvh, bool in, int cx, int cy, const uint2& dir,
                                                   uint32_t& beg, uint32_t& end) {
  const uint32_t bit = in ? (uint32_t) cy * kOccDim + cx : 0u;
  const bool set = in && ((dir.x >> (bit & 31)) & 1u);
  const uint32_t id = set ? dir.y + __popc(dir.x & ((1u << (bit & 31)) - 1)) : 0u;
  beg = __ldg(&bvh.cell_begin[id]);
  end = set ? __ldg(&bvh.cell_begin[id + 1]) : beg;
}

// One tile: lane = one query edge (start point p) whose box spans at most 3 x 3 cells.
static __device__ __forceinline__ void cells_tile(const MapView& Q, const BvhView& bvh, uint32_t p, bool valid,
                                                  CellWork& W, Emit& E, int lane, TravStats& st) {
  int4 qb = empty_box();
  int cx0 = 0, cy0 = 0, cx1 = -1, cy1 = -1;
  if (valid) {
    const longlong2 a = __ldg(&Q.pts[p]), b = __ldg(&Q.pts[p + 1]);
    qb = make_int4(quant(min(a.x, b.x)), quant(min(a.y, b.y)), quant(max(a.x, b.x)), quant(max(a.y, b.y)));
    cx0 = occ_cell(qb.x); cy0 = occ_cell(qb.y); cx1 = occ_cell(qb.z); cy1 = occ_cell(qb.w);
  }
  // the 2 x 2 cells at the min corner: all look-ups in flight together (two rounds of loads)
  uint32_t lb[9], le[9];
#pragma unroll
  for (int k = 0; k < 9; k++) lb[k] = le[k] = 0;
  {
    uint2 dir[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int dx = k & 1, dy = k >> 1;
      dir[k] = cell_dir(bvh, valid && cx0 + dx <= cx1 && cy0 + dy <= cy1, cx0 + dx, cy0 + dy);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int dx = k & 1, dy = k >> 1;
      cell_bounds(bvh, valid && cx0 + dx <= cx1 && cy0 + dy <= cy1, cx0 + dx, cy0 + dy, dir[k], lb[dy * 3 + dx],
                  le[dy * 3 + dx]);
    }
  }
  // third column / row: only for the rare edge spanning three cells
  if (__any_sync(0xffffffffu, valid && (cx1 - cx0 == 2 || cy1 - cy0 == 2))) {
#pragma unroll
    for (int k = 0; k < 9; k++) {
      if (!(k % 3 == 2 || k / 3 == 2)) continue;
      const bool in = valid && cx0 + k % 3 <= cx1 && cy0 + k / 3 <= cy1;
      const uint2 dir = cell_dir(bvh, in, cx0 + k % 3, cy0 + k / 3);
      cell_bounds(bvh, in, cx0 + k % 3, cy0 + k / 3, dir, lb[k], le[k]);
    }
  }
  uint32_t total = 0;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 9; k++) {
    total += le[k] - lb[k];
    W.beg[lane][k] = lb[k];
    W.cum[lane][k] = total;
  }
  W.qb[lane] = qb;
  W.p[lane] = p;
  uint32_t inc = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  W.pre[lane] = inc - total;
  if (lane == 31) W.pre[32] = inc;
  __syncwarp();
  const uint32_t all = W.pre[32];
  for (uint32_t base = 0; base < all; base += 32 * kCellBatch) {
    uint32_t owner[kCellBatch], cell[kCellBatch];
    uint4 item[kCellBatch];
    bool act[kCellBatch];
#pragma unroll
    for (int s2 = 0; s2 < kCellBatch; s2++) {
      const uint32_t i = base + s2 * 32 + lane;
      act[s2] = i < all;
      // owner = last lane whose prefix is <= i
      int lo = 0, hi = 32;
#pragma unroll
      for (int it = 0; it < 5; it++) {
        const int mid = (lo + hi) >> 1;
        if (W.pre[mid] <= i) lo = mid; else hi = mid;
      }
      owner[s2] = (uint32_t) lo;
      const uint32_t r = i - W.pre[lo];
      uint32_t k = 0, before = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t cj = W.cum[lo][j];
        if (r >= cj) {
          k = j + 1;
          before = cj;
        }
      }
      cell[s2] = k;
      item[s2] = act[s2] ? __ldg(&bvh.cell_item[W.beg[lo][k] + (r - before)]) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int s2 = 0; s2 < kCellBatch; s2++) {
      bool hit = false;
      uint32_t qp = 0;
      if (act[s2]) {
        const int4 oq = W.qb[owner[s2]];
        qp = W.p[owner[s2]];
        const CellBoxQ qc = cell_clip(oq, occ_cell(oq.x) + (int) (cell[s2] % 3u), occ_cell(oq.y) + (int) (cell[s2] / 3u));
        hit = cell_item_hit(item[s2], qc);
      }
      // DIRECT pair: {query start point | (count - 1) << 29, first point of the leaf}
      lsi_leaf<false>((int) item[s2].w, hit, qp | ((item[s2].z >> 14) << kDirectShift), E, lane, st);
    }
  }
}

__global__ void __launch_bounds__(kLsiWarps * 32)
k_lsi_cells(MapView Q, BvhView bvh, const uint32_t* __restrict__ survivors,
            const unsigned int* __restrict__ n_survivors_dev, uint2* __restrict__ out, uint32_t cap,
            unsigned int* counter) {
  __shared__ uint2 s_emit[kLsiWarps][kEmitBuf];
  __shared__ CellWork s_work[kLsiWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Emit E = {s_emit[warp], 0u, out, cap, counter, nullptr};
  TravStats st = {0, 0, 0, 0, 0};
  pdl_wait();  // the filter's survivors
  pdl_launch_dependents();
  const uint32_t n = load_count(n_survivors_dev);
  const uint32_t n_tiles = (n + 31) / 32;
  for (uint32_t tile = blockIdx.x * kLsiWarps + warp; tile < n_tiles; tile += gridDim.x * kLsiWarps) {
    const uint32_t slot = tile * 32 + lane;
    const bool valid = slot < n;
    cells_tile(Q, bvh, valid ? survivors[slot] : 0u, valid, s_work[warp], E, lane, st);
  }
  emit_flush(E, lane);
}

// The walk is a chain of dependent loads (row of level k+1 after the ballots of level k).
// When a row has SEVERAL candidate slots, their child rows are requested at once -- one
// lane per candidate -- so that the second, third ... descent finds its row in L2
// instead of paying a DRAM round trip each.
static __device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

static __device__ __forceinline__ void prefetch_row(const BvhView& bvh, uint32_t first_slot) {
  const char* b = reinterpret_cast<const char*>(bvh.top_box + first_slot);  // 32 x 16 B
  prefetch_l2(b);
  prefetch_l2(b + 128);
  prefetch_l2(b + 256);
  prefetch_l2(b + 384);
  prefetch_l2(bvh.top_code + first_slot);  // 32 x 4 B
}

static __device__ __forceinline__ void prefetch_node(const BvhView& bvh, int node) {
  prefetch_l2(bvh.node_box + 2 * node);
  prefetch_l2(bvh.node_child + node);
}

template <bool kStats>
__global__ void __launch_bounds__(kLsiWarps * 32)
k_lsi_bvh(MapView Q, MapView B, BvhView bvh, const uint32_t* __restrict__ order, uint32_t n_slots,
          uint32_t p_lo, const unsigned int* __restrict__ n_slots_dev, uint32_t spw,
          uint2* __restrict__ out, uint32_t cap, unsigned int* counter,
          unsigned long long* stats, bool direct) {
  __shared__ int s_stack[kLsiWarps][kStackDepth];
  __shared__ uint2 s_emit[kLsiWarps][kEmitBuf];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stack = s_stack[warp];
  // direct: the pairs join those of the cell directory, in its format
  Emit E = {s_emit[warp], 0u, out, cap, counter, direct ? bvh.leaf_rec : nullptr};
  pdl_wait();  // the filter's survivors / the cell kernel's long list
  pdl_launch_dependents();
  if (n_slots_dev) n_slots = load_count(n_slots_dev);  // survivor count of the pre-filter
  // spw = query slots per warp: 32, or fewer for a list of queries from all over the map
  // (the long edges the cell directory leaves over): a warp follows the clusters of its
  // queries one after the other, and 32 unrelated queries are 32 clusters
  // plain slots (no list): n_slots = end of the query window, p_lo = its begin
  const uint32_t n_tiles = (n_slots + spw - 1) / spw;
  const uint32_t tile = (order ? 0u : p_lo / 32) + blockIdx.x * kLsiWarps + warp;
  TravStats st = {0, 0, 0, 0, 0};
  const int4 kNeutral = empty_box();
  const int4 kEmpty = empty_box();
  if (tile >= n_tiles) return;
  const QTile cur = load_tile(Q, order, n_slots, tile, lane, spw, order ? 0u : p_lo);
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
        lsi_leaf<kStats>(~code0, h0, qe, E, lane, st);
        continue;
      }
      const int4 U1 = warp_union(h0 ? qb : kNeutral);
      // level 1: the 32 depth-10 nodes below slot g
      const int4 b1 = __ldg(&bvh.top_box[kTopOff1 + g * 32 + lane]);
      const int c1 = __ldg(&bvh.top_code[kTopOff1 + g * 32 + lane]);
      unsigned m1 = __ballot_sync(0xffffffffu, box_overlap(U1, b1));
      if (kStats) st.top_steps++;
      if ((m1 & (m1 - 1)) && ((m1 >> lane) & 1u) && c1 >= 0)
        prefetch_row(bvh, kTopOff2 + (uint32_t) (g * 32 + lane) * 32);
      while (m1) {
        const int h = __ffs(m1) - 1;
        m1 &= m1 - 1;
        const int4 sb1 = shfl_box(b1, h);
        const bool h1 = h0 && box_overlap(qb, sb1);
        if (__ballot_sync(0xffffffffu, h1) == 0) continue;
        const int code1 = __shfl_sync(0xffffffffu, c1, h);
        if (code1 < 0) {
          lsi_leaf<kStats>(~code1, h1, qe, E, lane, st);
          continue;
        }
        const int4 U2 = warp_union(h1 ? qb : kNeutral);
        // level 2: the 32 depth-15 nodes below slot (g, h)
        const int4 b2 = __ldg(&bvh.top_box[kTopOff2 + (g * 32 + h) * 32 + lane]);
        const int c2 = __ldg(&bvh.top_code[kTopOff2 + (g * 32 + h) * 32 + lane]);
        unsigned m2 = __ballot_sync(0xffffffffu, box_overlap(U2, b2));
        if (kStats) st.top_steps++;
        if ((m2 & (m2 - 1)) && ((m2 >> lane) & 1u) && c2 >= 0) {
          if (bvh.top_levels == 4) prefetch_row(bvh, kTopOff3 + (uint32_t) ((g * 32 + h) * 32 + lane) * 32);
          else prefetch_node(bvh, c2);
        }
        while (m2) {
          const int i = __ffs(m2) - 1;
          m2 &= m2 - 1;
          const int4 sb2 = shfl_box(b2, i);
          const bool h2 = h1 && box_overlap(qb, sb2);
          if (__ballot_sync(0xffffffffu, h2) == 0) continue;
          const int code2 = __shfl_sync(0xffffffffu, c2, i);
          if (code2 < 0) {
            lsi_leaf<kStats>(~code2, h2, qe, E, lane, st);
            continue;
          }
          if (bvh.top_levels < 4) {
            lsi_subtree<kStats>(bvh, code2, stack, h2 ? qb : kEmpty, qe, E, lane, st);
            continue;
          }
          // level 3 (big trees only): the 32 depth-20 nodes below slot (g, h, i)
          const int4 U3 = warp_union(h2 ? qb : kNeutral);
          const uint32_t s3 = kTopOff3 + ((uint32_t) ((g * 32 + h) * 32 + i)) * 32 + lane;
          const int4 b3 = __ldg(&bvh.top_box[s3]);
          const int c3 = __ldg(&bvh.top_code[s3]);
          unsigned m3 = __ballot_sync(0xffffffffu, box_overlap(U3, b3));
          if (kStats) st.top_steps++;
          if ((m3 & (m3 - 1)) && ((m3 >> lane) & 1u) && c3 >= 0) prefetch_node(bvh, c3);
          while (m3) {
            const int j = __ffs(m3) - 1;
            m3 &= m3 - 1;
            const int4 sb3 = shfl_box(b3, j);
            const bool h3 = h2 && box_overlap(qb, sb3);
            if (__ballot_sync(0xffffffffu, h3) == 0) continue;
            const int code3 = __shfl_sync(0xffffffffu, c3, j);
            if (code3 < 0)
              lsi_leaf<kStats>(~code3, h3, qe, E, lane, st);
            else
              lsi_subtree<kStats>(bvh, code3, stack, h3 ? qb : kEmpty, qe, E, lane, st);
          }
        }
      }
    }
  }
  emit_flush(E, lane);
  if (kStats && lane == 0) {
    atomicAdd(stats + 0, (unsigned long long) st.nodes);
    atomicAdd(stats + 1, (unsigned long long) st.leaves);
    atomicAdd(stats + 2, (unsigned long long) st.top_steps);
    atomicAdd(stats + 3, (unsigned long long) st.lane_leaf);
    atomicAdd(stats + 4, (unsigned long long) (st.leaves ? 1 : 0));
    atomicMax(stats + 5, (unsigned long long) st.maxsp);
  }
}

// Exact pass 1 over the (query start point, leaf) pairs of the traversal, in two dense
// steps per CTA:
//  (1) one thread per pair loads the leaf's vertices (<= 9 consecutive 16-byte points, all
//      loads of a thread in flight together) and tests the exact integer boxes of the
//      query edge against each base edge of the leaf.  The ~1 in 8 that overlap go to a
//      shared-memory list;
//  (2) whenever the list holds a CTA's worth, intersect_test runs over it with every
//      lane busy, and the hits are compacted into the result queue (one atomic per
//      warp) as start-point index pairs.
// (Running intersect_test where the box test passes kept 3 of 32 lanes busy: ncu showed
// the int128 predicate executed by nearly every warp for one or two lanes.)
// The pair count is read on the device: no host round trip between the kernels.
constexpr int kExactThreads = 256;
constexpr int kExactList = 9 * kExactThreads;  // < kExactThreads before a round, <= 8 per thread added

static __device__ __forceinline__ void exact_drain(const MapView& Q, const MapView& B, const uint2* list,
                                                   unsigned n_list, rjb_xsect* __restrict__ out,
                                                   uint32_t cap, unsigned int* counter) {
  const int lane = threadIdx.x & 31;
  for (unsigned t0 = threadIdx.x - lane; t0 < n_list; t0 += kExactThreads) {
    const unsigned t = t0 + lane;
    bool found = false;
    uint2 it = make_uint2(0, 0);
    if (t < n_list) {
      it = list[t];
      const longlong2 a = __ldg(&Q.pts[it.x]), b = __ldg(&Q.pts[it.x + 1]);
      const longlong2 c = __ldg(&B.pts[it.y]), d = __ldg(&B.pts[it.y + 1]);
      const Seg e1 = {a.x, a.y, b.x, b.y}, e2 = {c.x, c.y, d.x, d.y};
      found = lsi_intersect(e1, e2);
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
        out[pos].eid[0] = it.x;  // point indices for now; pass 2 turns them into eids
        out[pos].eid[1] = it.y;
      }
    }
  }
}

// append the lanes with `pass` to the CTA's list (one shared-memory atomic per warp)
static __device__ __forceinline__ void exact_push(bool pass, uint2 item, uint2* list, unsigned* s_n,
                                                  int lane, unsigned& cand) {
  const unsigned m = __ballot_sync(0xffffffffu, pass);
  if (m == 0) return;
  unsigned base = 0;
  const int leader = __ffs(m) - 1;
  if (lane == leader) base = atomicAdd(s_n, (unsigned) __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pass) list[base + __popc(m & ((1u << lane) - 1))] = item;
  cand += __popc(m);  // counted by every lane alike
}

template <bool kDirect>
__global__ void __launch_bounds__(kExactThreads)
k_lsi_exact(MapView Q, MapView B, const uint2* __restrict__ pairs, const uint2* __restrict__ leaf_rec,
            const unsigned int* __restrict__ n_pairs_dev, uint32_t pair_cap,
            rjb_xsect* __restrict__ out, uint32_t cap, unsigned int* counter,
            unsigned long long* n_cand) {
  __shared__ uint2 s_list[kExactList];
  __shared__ unsigned s_n;
  if (threadIdx.x == 0) s_n = 0;
  pdl_wait();
  __syncthreads();
  const uint32_t n = min(load_count(n_pairs_dev), pair_cap);
  const int lane = threadIdx.x & 31;
  unsigned cand = 0;
  // block-uniform trip count: every thread reaches the barriers
  for (uint64_t i0 = (uint64_t) blockIdx.x * kExactThreads; i0 < n; i0 += (uint64_t) gridDim.x * kExactThreads) {
    const uint64_t i = i0 + threadIdx.x;
    uint32_t pq = 0, pb0 = 0, cnt = 0;
    Seg e1 = {0, 0, 0, 0};
    longlong2 bp[5];
#pragma unroll
    for (int k = 0; k < 5; k++) bp[k] = make_longlong2(0, 0);
    if (i < n) {
      const uint2 pr = pairs[i];
      if (kDirect) {
        pq = pr.x & ((1u << kDirectShift) - 1);
        cnt = (pr.x >> kDirectShift) + 1;
        pb0 = pr.y;
      } else {
        pq = pr.x;
        const uint2 rec = __ldg(&leaf_rec[pr.y]);
        cnt = rec.y >> 28;
        pb0 = rec.x + (rec.y & 0x0FFFFFFFu);
      }
      const longlong2 a = __ldg(&Q.pts[pq]), b = __ldg(&Q.pts[pq + 1]);
      e1 = {a.x, a.y, b.x, b.y};
#pragma unroll
      for (int k = 0; k < 5; k++)
        if ((uint32_t) k <= cnt) bp[k] = __ldg(&B.pts[pb0 + k]);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const Seg e2 = {bp[k].x, bp[k].y, bp[k + 1].x, bp[k + 1].y};
      exact_push((uint32_t) k < cnt && seg_boxes_overlap(e1, e2), make_uint2(pq, pb0 + k), s_list, &s_n,
                 lane, cand);
    }
    if (__any_sync(0xffffffffu, cnt > 4)) {  // leaves of 5..8 edges (lbvh_leaf_size > 4)
      longlong2 p1 = bp[4];
      for (uint32_t k = 4; k < 8; k++) {
        bool pass = false;
        if (k < cnt) {
          const longlong2 p2 = __ldg(&B.pts[pb0 + k + 1]);
          const Seg e2 = {p1.x, p1.y, p2.x, p2.y};
          pass = seg_boxes_overlap(e1, e2);
          p1 = p2;
        }
        exact_push(pass, make_uint2(pq, pb0 + k), s_list, &s_n, lane, cand);
      }
    }
    __syncthreads();
    const unsigned n_list = s_n;
    if (n_list >= kExactThreads) {
      exact_drain(Q, B, s_list, n_list, out, cap, counter);
      __syncthreads();
      if (threadIdx.x == 0) s_n = 0;
      __syncthreads();
    }
  }
  __syncthreads();
  exact_drain(Q, B, s_list, s_n, out, cap, counter);
  if (lane == 0 && cand) atomicAdd(n_cand, (unsigned long long) cand);
}

// Exact pass 2, dense over the hits: rational intersection point and the final
// edge ids -> rjb_xsect (reference computes it inside the traversal callback,
// lsi_lbvh.h:69-78; its RT backend has the same post-pass, src/app/lsi_rt.h:66-112).
//
// One thread per hit computes both coordinates.  94 % of the coordinates are decided
// without the 128-bit gcd (lsi_point_axis<true>); the rest would keep every warp waiting
// for one or two lanes' ~10^3-instruction gcd chains, so they are parked in a
// shared-memory list and worked off densely -- 32 deferred coordinates per warp -- when
// the list fills up and at the end.
constexpr int kPointsThreads = 256;
constexpr int kDeferCap = 3 * kPointsThreads;  // < kPointsThreads before a round, <= 2 per thread added

struct DeferItem {
  uint32_t i, pq, pb, axis;
};

// (Q = map of DeferItem::pq = the e1 side, B = map of pb = the e2 side)
static __device__ __forceinline__ void points_flush(const MapView& Q, const MapView& B,
                                                    const DeferItem* list, unsigned n_list,
                                                    rjb_xsect* __restrict__ out) {
  for (unsigned t = threadIdx.x; t < n_list; t += kPointsThreads) {
    const DeferItem it = list[t];
    const longlong2 a = __ldg(&Q.pts[it.pq]), b = __ldg(&Q.pts[it.pq + 1]);
    const longlong2 c = __ldg(&B.pts[it.pb]), d = __ldg(&B.pts[it.pb + 1]);
    const Seg e1 = {a.x, a.y, b.x, b.y}, e2 = {c.x, c.y, d.x, d.y};
    const long long v = lsi_point_axis<false>(e1, e2, (int) it.axis, nullptr);
    if (it.axis == 0) out[it.i].x = v; else out[it.i].y = v;
  }
}

__global__ void __launch_bounds__(kPointsThreads)
k_lsi_points(MapView Q, MapView B, int query_map_id, const unsigned int* __restrict__ counter,
             uint32_t cap, rjb_xsect* __restrict__ out, bool base_first) {
  __shared__ DeferItem s_list[kDeferCap];
  __shared__ unsigned s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const uint32_t n = min(*counter, cap);
  // block-uniform trip count: every thread reaches the barriers
  for (uint64_t i0 = (uint64_t) blockIdx.x * kPointsThreads; i0 < n; i0 += (uint64_t) gridDim.x * kPointsThreads) {
    const uint64_t i64 = i0 + threadIdx.x;
    if (i64 < n) {
      const uint32_t i = (uint32_t) i64;
      const uint32_t pq = out[i].eid[0], pb = out[i].eid[1];
      const longlong2 a = __ldg(&Q.pts[pq]), b = __ldg(&Q.pts[pq + 1]);
      const longlong2 c = __ldg(&B.pts[pb]), d = __ldg(&B.pts[pb + 1]);
      const uint32_t eq = pq - __ldg(&Q.point_chain[pq]), eb = pb - __ldg(&B.point_chain[pb]);
      // e1 = query edge (lsi_lbvh.h:71); grid mode with query map 1: e1 = the map-0 (base) edge
      // (lsi_grid.h:62) -- the point is the same rational either way, the order is kept anyway
      const Seg sq = {a.x, a.y, b.x, b.y}, sb = {c.x, c.y, d.x, d.y};
      const Seg& e1 = base_first ? sb : sq;
      const Seg& e2 = base_first ? sq : sb;
      bool dx = false, dy = false;
      rjb_xsect r;
      r.x = lsi_point_axis<true>(e1, e2, 0, &dx);
      r.y = lsi_point_axis<true>(e1, e2, 1, &dy);
      if (dx) s_list[atomicAdd(&s_n, 1u)] = {i, base_first ? pb : pq, base_first ? pq : pb, 0u};
      if (dy) s_list[atomicAdd(&s_n, 1u)] = {i, base_first ? pb : pq, base_first ? pq : pb, 1u};
      r.eid[0] = query_map_id == 0 ? eq : eb;
      r.eid[1] = query_map_id == 0 ? eb : eq;
      r.mid_point_polygon_id = RJB_DONTKNOW;
      r._pad = 0;
      out[i] = r;
    }
    __syncthreads();  // records written, list complete for this round
    const unsigned n_list = s_n;
    if (n_list >= kPointsThreads) {
      points_flush(base_first ? B : Q, base_first ? Q : B, s_list, n_list, out);
      __syncthreads();
      if (threadIdx.x == 0) s_n = 0;
      __syncthreads();
    }
  }
  __syncthreads();
  points_flush(base_first ? B : Q, base_first ? Q : B, s_list, s_n, out);
}

// ---- fused exact + point pass (option lsi_fused, default) -----------------------------------
// k_lsi_exact and k_lsi_points as ONE kernel, and the last kernel of the query: every kernel
// of this pipeline costs ~20 us beyond its issue time (launch, a chain of dependent loads
// through a flushed L2, a tail), and the point pass re-loaded everything the predicate had in
// registers -- the hit list, four vertices and two chain ids per hit.  Here the lane that finds a
// hit computes both coordinates at once (gcd-free path) and writes the finished record; only
// the coordinates that need the 128-bit gcd (6 %) are parked in shared memory and worked off
// densely, like before.  The LAST CTA to finish copies the query's counters straight into
// mapped pinned host memory and zeroes them for the next query: no memset before and no
// memcpy after the kernels (each ~3-5 us of stream time on a 0.15 ms query).
constexpr int kLsiCtrs = 10;  // 64-bit words the host reads back

// Optional phase trace (build with -DRJB_TRACE; rjb_debug_trace reads it): thread 0 of every
// CTA notes %globaltimer at the phase boundaries of k_lsi_resolve.
#ifdef RJB_TRACE
constexpr int kTraceSlots = 16;
__device__ unsigned long long g_trace[4096 * kTraceSlots];
__device__ unsigned long long g_trace_w[4096 * 4 * kTraceSlots];  // lane 0 of every warp
static __device__ __forceinline__ void trace_mark(int& k) {
  if ((threadIdx.x & 31) == 0 && blockIdx.x < 4096 && k < kTraceSlots) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (threadIdx.x == 0) g_trace[blockIdx.x * kTraceSlots + k] = t;
    g_trace_w[(blockIdx.x * 4 + (threadIdx.x >> 5 & 3)) * kTraceSlots + k] = t;
  }
  k++;
}
#define RJB_MARK(k) trace_mark(k)
#else
#define RJB_MARK(k)
#endif

struct LsiTail {
  unsigned long long* ctr;   // device counters [0, kLsiCtrs)
  unsigned long long* host;  // their mapped pinned copy (device pointer), or nullptr: leave them
  unsigned int* ticket;      // CTAs that have finished; zero before and after the kernel
};

// (all threads of the CTA)
static __device__ __forceinline__ void lsi_tail(const LsiTail& T) {
  __shared__ bool s_last;
  if (!T.host) return;
#ifdef RJB_TRACE
  int tk = 6;  // slots 6..9 (the rounds rarely get that far)
#endif
  __syncthreads();  // the CTA's own atomics are issued
  RJB_MARK(tk);  // 6: barrier B passed
  if (threadIdx.x == 0) {
    __threadfence();
    RJB_MARK(tk);  // 7: fence done
    s_last = atomicAdd(T.ticket, 1u) == gridDim.x - 1;
    RJB_MARK(tk);  // 8: ticket back
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < kLsiCtrs) {
    T.host[threadIdx.x] = ((volatile unsigned long long*) T.ctr)[threadIdx.x];
    T.ctr[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) *T.ticket = 0;
}

// Build-time variants, measured on the B200 (k_lsi_resolve on the bench workload, us):
//   defaults below                                                   47.4
//   RJB_RESOLVE_JOINT=1  both axes at once (lsi_point_both), parked state   53.0  (88 registers)
//   RJB_RESOLVE_INL=__noinline__  one copy of drain / flush (72 KB code)    51.0  (240 B stack)
//   RJB_RESOLVE_ROT=1    flush warp rotates with the CTA                     49.1
//   RJB_RESOLVE_THREADS=128  twice as many, smaller CTAs                     49.1
// (tools/trace_resolve.py with -DRJB_TRACE shows the phases: 3 box rounds of 2 us, 2-3 drains
// of 6-9 us, a final flush of 15 us of which the gcd itself is 3.5 us.)
#ifndef RJB_RESOLVE_INL
#define RJB_RESOLVE_INL __forceinline__
#endif
#ifndef RJB_RESOLVE_ROT
#define RJB_RESOLVE_ROT 0
#endif
#ifndef RJB_RESOLVE_JOINT
#define RJB_RESOLVE_JOINT 0
#endif
#ifndef RJB_RESOLVE_THREADS
#define RJB_RESOLVE_THREADS 256
#endif
#ifndef RJB_RESOLVE_MIN_CTAS
#define RJB_RESOLVE_MIN_CTAS 1
#endif
constexpr int kResolveThreads = RJB_RESOLVE_THREADS;  // small CTAs: the gcd tail of one overlaps the others' work
constexpr int kResolveList = 9 * kResolveThreads;
constexpr int kResolveDefer = 3 * kResolveThreads;  // < kResolveThreads before a drain, <= 2 per thread added

// a parked coordinate: where it goes, and the state the gcd path continues from (or, for long
// edges, the pair to redo from scratch)
struct ResolveDefer {
  uint32_t i;       // result slot
  uint32_t axis;    // 0 / 1, + 2 when the item must be redone with lsi_point_axis<false>
  PointState st;    // (redo: X0 = query start point | base start point << 32)
};

// deferred coordinates of the CTA: the gcd path, one item per thread.  A flush usually has a
// dozen items, i.e. ONE busy warp per CTA running a ~1400-instruction dependent chain.  Warp w of
// every CTA sits on sub-partition w % 4, so with the items in warp 0 the flushes of the six
// CTAs of an SM all queued on one scheduler (ncu / %globaltimer trace: 15 us per flush against
// 3.5 us for the same code alone); the warp that takes the first items rotates with the CTA.
// (ONE copy of the code: the kernel calls it from two places, and a 170 KB kernel does not stay in
// the instruction cache)
static __device__ RJB_RESOLVE_INL void resolve_flush(const MapView& Q, const MapView& B, const ResolveDefer* list,
                                                     unsigned n_list, rjb_xsect* __restrict__ out) {
  const unsigned rot = RJB_RESOLVE_ROT ? 32u * ((blockIdx.x / kNumSMs + blockIdx.x) % (kResolveThreads / 32)) : 0u;
  for (unsigned t = (threadIdx.x + kResolveThreads - rot) % kResolveThreads; t < n_list; t += kResolveThreads) {
    const ResolveDefer it = list[t];
    long long v;
    if (it.axis & 2u) {
      const uint32_t pq = (uint32_t) (unsigned long long) it.st.X0, pb = (uint32_t) ((unsigned long long) it.st.X0 >> 32);
      const longlong2 a = __ldg(&Q.pts[pq]), b = __ldg(&Q.pts[pq + 1]);
      const longlong2 c = __ldg(&B.pts[pb]), d = __ldg(&B.pts[pb + 1]);
      const Seg e1 = {a.x, a.y, b.x, b.y}, e2 = {c.x, c.y, d.x, d.y};
      v = lsi_point_axis_slow(e1, e2, (int) (it.axis & 1u));
    } else {
      v = lsi_point_finish(it.st);
    }
    if ((it.axis & 1u) == 0) out[it.i].x = v; else out[it.i].y = v;
  }
}

// intersect_test over the CTA's list of box-overlapping pairs, every lane busy; a hit becomes a
// finished rjb_xsect record at once (deferred coordinates: s_defer, flushed whenever a CTA's
// worth has gathered).  Called by ALL threads with the same n_list; contains barriers.
static __device__ RJB_RESOLVE_INL void resolve_drain(const MapView& Q, const MapView& B, int query_map_id,
                                                     const uint2* list, unsigned n_list, ResolveDefer* s_defer,
                                                     unsigned* s_nd, rjb_xsect* __restrict__ out, uint32_t cap,
                                                     unsigned int* counter) {
  const int lane = threadIdx.x & 31;
  for (unsigned r0 = 0; r0 < n_list; r0 += kResolveThreads) {  // block-uniform trip count (barriers inside)
    const unsigned t = r0 + threadIdx.x;
    bool found = false;
    uint2 it = make_uint2(0, 0);
    Seg e1 = {0, 0, 0, 0}, e2 = {0, 0, 0, 0};
    uint32_t cq = 0, cb = 0;
    if (t < n_list) {
      it = list[t];
      const longlong2 a = __ldg(&Q.pts[it.x]), b = __ldg(&Q.pts[it.x + 1]);
      const longlong2 c = __ldg(&B.pts[it.y]), d = __ldg(&B.pts[it.y + 1]);
      cq = __ldg(&Q.point_chain[it.x]);  // with the vertices: no dependent round trip for the hits
      cb = __ldg(&B.point_chain[it.y]);
      e1 = {a.x, a.y, b.x, b.y};
      e2 = {c.x, c.y, d.x, d.y};
      found = lsi_intersect(e1, e2);
    }
    const unsigned m = __ballot_sync(0xffffffffu, found);
    unsigned base = 0;
    const int leader = __ffs(m) - 1;
    if (m && lane == leader) base = atomicAdd(counter, (unsigned) __popc(m));
    if (m) base = __shfl_sync(0xffffffffu, base, leader);
    if (found) {
      const unsigned pos = base + __popc(m & ((1u << lane) - 1));
      if (pos < cap) {
        long long xy[2];
        int code[2];
        PointState st[2];
#if RJB_RESOLVE_JOINT
        lsi_point_both(e1, e2, xy, code, st);
#else
        for (int axis = 0; axis < 2; axis++) {
          bool def = false;
          xy[axis] = lsi_point_axis<true>(e1, e2, axis, &def);
          code[axis] = def ? kPointRedo : kPointDone;
        }
#endif
        rjb_xsect r;
        r.x = xy[0];
        r.y = xy[1];
#pragma unroll
        for (int axis = 0; axis < 2; axis++)
          if (code[axis] != kPointDone) {
            ResolveDefer d;
            d.i = pos;
            d.axis = (uint32_t) axis | (code[axis] == kPointRedo ? 2u : 0u);
            d.st = st[axis];
            if (code[axis] == kPointRedo) d.st.X0 = (long long) ((unsigned long long) it.x | ((unsigned long long) it.y << 32));
            s_defer[atomicAdd(s_nd, 1u)] = d;
          }
        const uint32_t eq = it.x - cq, eb = it.y - cb;
        r.eid[0] = query_map_id == 0 ? eq : eb;
        r.eid[1] = query_map_id == 0 ? eb : eq;
        r.mid_point_polygon_id = RJB_DONTKNOW;
        r._pad = 0;
        out[pos] = r;
      }
    }
    // <= 2 deferred coordinates per thread and round: the list never exceeds kResolveDefer
    __syncthreads();
    if (*s_nd >= kResolveThreads) {
      resolve_flush(Q, B, s_defer, *s_nd, out);
      __syncthreads();
      if (threadIdx.x == 0) *s_nd = 0;
      __syncthreads();
    }
  }
}

template <bool kDirect>
__global__ void __launch_bounds__(kResolveThreads, RJB_RESOLVE_MIN_CTAS)
k_lsi_resolve(MapView Q, MapView B, int query_map_id, const uint2* __restrict__ pairs,
              const uint2* __restrict__ leaf_rec, const unsigned int* __restrict__ n_pairs_dev, uint32_t pair_cap,
              rjb_xsect* __restrict__ out, uint32_t cap, unsigned int* counter, unsigned long long* n_cand,
              LsiTail tail) {
  __shared__ uint2 s_list[kResolveList];
  __shared__ ResolveDefer s_defer[kResolveDefer];
  __shared__ unsigned s_n, s_nd;
  if (threadIdx.x == 0) s_n = s_nd = 0;
#ifdef RJB_TRACE
  int tk = 0;
  if (threadIdx.x == 0 && blockIdx.x < 4096)
    for (int z = 0; z < kTraceSlots; z++) g_trace[blockIdx.x * kTraceSlots + z] = 0;
#endif
  RJB_MARK(tk);  // 0: start
  pdl_wait();  // the (query, leaf) pairs
  __syncthreads();
  const uint32_t n = min(load_count(n_pairs_dev), pair_cap);
  const int lane = threadIdx.x & 31;
  unsigned cand = 0;
  RJB_MARK(tk);  // 1: count read
  // block-uniform trip count: every thread reaches the barriers
  for (uint64_t i0 = (uint64_t) blockIdx.x * kResolveThreads; i0 < n; i0 += (uint64_t) gridDim.x * kResolveThreads) {
    const uint64_t i = i0 + threadIdx.x;
    uint32_t pq = 0, pb0 = 0, cnt = 0;
    Seg e1 = {0, 0, 0, 0};
    longlong2 bp[5];
#pragma unroll
    for (int k = 0; k < 5; k++) bp[k] = make_longlong2(0, 0);
    if (i < n) {
      const uint2 pr = pairs[i];
      if (kDirect) {
        pq = pr.x & ((1u << kDirectShift) - 1);
        cnt = (pr.x >> kDirectShift) + 1;
        pb0 = pr.y;
      } else {
        pq = pr.x;
        const uint2 rec = __ldg(&leaf_rec[pr.y]);
        cnt = rec.y >> 28;
        pb0 = rec.x + (rec.y & 0x0FFFFFFFu);
      }
      const longlong2 a = __ldg(&Q.pts[pq]), b = __ldg(&Q.pts[pq + 1]);
      e1 = {a.x, a.y, b.x, b.y};
#pragma unroll
      for (int k = 0; k < 5; k++)
        if ((uint32_t) k <= cnt) bp[k] = __ldg(&B.pts[pb0 + k]);
    }
    // the <= 4 edges of the leaf in ONE push: a single shared-memory atomic per warp and round
    unsigned pass_m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const Seg e2 = {bp[k].x, bp[k].y, bp[k + 1].x, bp[k + 1].y};
      if ((uint32_t) k < cnt && seg_boxes_overlap(e1, e2)) pass_m |= 1u << k;
    }
    if (__any_sync(0xffffffffu, cnt > 4)) {  // leaves of 5..8 edges (lbvh_leaf_size > 4)
      longlong2 p1 = bp[4];
      for (uint32_t k = 4; k < 8; k++) {
        if (k < cnt) {
          const longlong2 p2 = __ldg(&B.pts[pb0 + k + 1]);
          const Seg e2 = {p1.x, p1.y, p2.x, p2.y};
          if (seg_boxes_overlap(e1, e2)) pass_m |= 1u << k;
          p1 = p2;
        }
      }
    }
    {
      const unsigned mine = __popc(pass_m);
      unsigned inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
      if (total) {
        unsigned base = 0;
        if (lane == 31) base = atomicAdd(&s_n, total);
        base = __shfl_sync(0xffffffffu, base, 31) + inc - mine;
        while (pass_m) {
          const uint32_t k = __ffs(pass_m) - 1;
          pass_m &= pass_m - 1;
          s_list[base++] = make_uint2(pq, pb0 + k);
        }
        cand += total;  // counted by every lane alike
      }
    }
    __syncthreads();
    RJB_MARK(tk);  // after the box stage of a round
    const unsigned n_list = s_n;
    if (n_list >= kResolveThreads) {
      resolve_drain(Q, B, query_map_id, s_list, n_list, s_defer, &s_nd, out, cap, counter);
      __syncthreads();
      if (threadIdx.x == 0) s_n = 0;
      __syncthreads();
      RJB_MARK(tk);  // after a drain
    }
  }
  __syncthreads();
#ifdef RJB_TRACE
  tk = 10;
#endif
  RJB_MARK(tk);  // 10: loop done
  resolve_drain(Q, B, query_map_id, s_list, s_n, s_defer, &s_nd, out, cap, counter);
  __syncthreads();
  RJB_MARK(tk);  // 11: final drain done
  resolve_flush(Q, B, s_defer, s_nd, out);
#ifdef RJB_TRACE
  { int k2 = 14; RJB_MARK(k2); }  // 14: back from the flush
#endif
  if (lane == 0 && cand) atomicAdd(n_cand, (unsigned long long) cand);
#ifdef RJB_TRACE
  { int k2 = 15; RJB_MARK(k2); }  // 15: atomic issued
#endif
  __syncthreads();
  RJB_MARK(tk);  // 12: final flush done
  lsi_tail(tail);
  RJB_MARK(tk);  // 13: end
}

// ---- k_lsi_resolve, warp-private variant (-DRJB_RESOLVE_WARP=1) ---------------------------------
// Same work as k_lsi_resolve, but every WARP keeps its own list of box-overlapping pairs and
// drains it -- 32 items, one per lane -- as soon as it holds 32: no CTA barrier until the very
// end (ncu on the CTA-wide version: 36 % of the warp time is barrier wait), the warps of an SM
// are in different phases and hide each other's latency.  Deferred (gcd) coordinates still go to
// one CTA list, flushed once at the end; if that list is full the lane takes the gcd path at once.
// MEASURED (option lsi_resolve_warp): 48.8 us against 47.2 us for the CTA-wide kernel -- 108
// registers (two CTAs per SM); 80 registers with spills 50.2 us, out-of-line drain 53.6 us.  The
// barrier waits were not the bound; the kernel is kept as the A/B it was.
constexpr int kWarpList = 9 * 32;  // < 32 carried over + <= 8 pushes per lane and round

#ifndef RJB_WDRAIN_INL
#define RJB_WDRAIN_INL __forceinline__
#endif
static __device__ RJB_WDRAIN_INL void warp_resolve_drain(const MapView& Q, const MapView& B, int query_map_id,
                                                          const uint2* items, unsigned n_items, ResolveDefer* s_defer,
                                                          unsigned* s_nd, rjb_xsect* __restrict__ out, uint32_t cap,
                                                          unsigned int* counter, int lane) {
  bool found = false;
  uint2 it = make_uint2(0, 0);
  Seg e1 = {0, 0, 0, 0}, e2 = {0, 0, 0, 0};
  uint32_t cq = 0, cb = 0;
  if ((unsigned) lane < n_items) {
    it = items[lane];
    const longlong2 a = __ldg(&Q.pts[it.x]), b = __ldg(&Q.pts[it.x + 1]);
    const longlong2 c = __ldg(&B.pts[it.y]), d = __ldg(&B.pts[it.y + 1]);
    cq = __ldg(&Q.point_chain[it.x]);
    cb = __ldg(&B.point_chain[it.y]);
    e1 = {a.x, a.y, b.x, b.y};
    e2 = {c.x, c.y, d.x, d.y};
    found = lsi_intersect(e1, e2);
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
      rjb_xsect r;
      long long xy[2];
#pragma unroll
      for (int axis = 0; axis < 2; axis++) {
        bool def = false;
        xy[axis] = lsi_point_axis<true>(e1, e2, axis, &def);
        if (def) {
          const unsigned slot = atomicAdd(s_nd, 1u);
          if (slot < (unsigned) kResolveDefer) {
            ResolveDefer dd;
            dd.i = pos;
            dd.axis = (uint32_t) axis | 2u;
            dd.st.X0 = (long long) ((unsigned long long) it.x | ((unsigned long long) it.y << 32));
            dd.st.rs = dd.st.aden = 0;
            s_defer[slot] = dd;
          } else {
            xy[axis] = lsi_point_axis_slow(e1, e2, axis);  // list full (never on the bench workload)
          }
        }
      }
      r.x = xy[0];
      r.y = xy[1];
      const uint32_t eq = it.x - cq, eb = it.y - cb;
      r.eid[0] = query_map_id == 0 ? eq : eb;
      r.eid[1] = query_map_id == 0 ? eb : eq;
      r.mid_point_polygon_id = RJB_DONTKNOW;
      r._pad = 0;
      out[pos] = r;
    }
  }
}

template <bool kDirect>
__global__ void __launch_bounds__(kResolveThreads, RJB_RESOLVE_MIN_CTAS)
k_lsi_resolve_w(MapView Q, MapView B, int query_map_id, const uint2* __restrict__ pairs,
                const uint2* __restrict__ leaf_rec, const unsigned int* __restrict__ n_pairs_dev, uint32_t pair_cap,
                rjb_xsect* __restrict__ out, uint32_t cap, unsigned int* counter, unsigned long long* n_cand,
                LsiTail tail) {
  constexpr int kWarps = kResolveThreads / 32;
  __shared__ uint2 s_list[kWarps][kWarpList];
  __shared__ ResolveDefer s_defer[kResolveDefer];
  __shared__ unsigned s_nd;
  if (threadIdx.x == 0) s_nd = 0;
  __syncthreads();
  const uint32_t n = min(load_count(n_pairs_dev), pair_cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2* list = s_list[warp];
  unsigned n_list = 0, cand = 0;  // warp-uniform
  for (uint64_t i0 = ((uint64_t) blockIdx.x * kWarps + warp) * 32; i0 < n; i0 += (uint64_t) gridDim.x * kWarps * 32) {
    const uint64_t i = i0 + lane;
    uint32_t pq = 0, pb0 = 0, cnt = 0;
    Seg e1 = {0, 0, 0, 0};
    longlong2 bp[5];
#pragma unroll
    for (int k = 0; k < 5; k++) bp[k] = make_longlong2(0, 0);
    if (i < n) {
      const uint2 pr = pairs[i];
      if (kDirect) {
        pq = pr.x & ((1u << kDirectShift) - 1);
        cnt = (pr.x >> kDirectShift) + 1;
        pb0 = pr.y;
      } else {
        pq = pr.x;
        const uint2 rec = __ldg(&leaf_rec[pr.y]);
        cnt = rec.y >> 28;
        pb0 = rec.x + (rec.y & 0x0FFFFFFFu);
      }
      const longlong2 a = __ldg(&Q.pts[pq]), b = __ldg(&Q.pts[pq + 1]);
      e1 = {a.x, a.y, b.x, b.y};
#pragma unroll
      for (int k = 0; k < 5; k++)
        if ((uint32_t) k <= cnt) bp[k] = __ldg(&B.pts[pb0 + k]);
    }
    unsigned pass_m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const Seg e2 = {bp[k].x, bp[k].y, bp[k + 1].x, bp[k + 1].y};
      if ((uint32_t) k < cnt && seg_boxes_overlap(e1, e2)) pass_m |= 1u << k;
    }
    if (__any_sync(0xffffffffu, cnt > 4)) {  // leaves of 5..8 edges (lbvh_leaf_size > 4)
      longlong2 p1 = bp[4];
      for (uint32_t k = 4; k < 8; k++) {
        if (k < cnt) {
          const longlong2 p2 = __ldg(&B.pts[pb0 + k + 1]);
          const Seg e2 = {p1.x, p1.y, p2.x, p2.y};
          if (seg_boxes_overlap(e1, e2)) pass_m |= 1u << k;
          p1 = p2;
        }
      }
    }
    const unsigned mine = __popc(pass_m);
    unsigned inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned at = n_list + inc - mine;
    while (pass_m) {
      const uint32_t k = __ffs(pass_m) - 1;
      pass_m &= pass_m - 1;
      list[at++] = make_uint2(pq, pb0 + k);
    }
    n_list += total;
    cand += total;
    __syncwarp();
    while (n_list >= 32) {
      n_list -= 32;
      warp_resolve_drain(Q, B, query_map_id, list + n_list, 32, s_defer, &s_nd, out, cap, counter, lane);
      __syncwarp();
    }
  }
  if (n_list) warp_resolve_drain(Q, B, query_map_id, list, n_list, s_defer, &s_nd, out, cap, counter, lane);
  if (lane == 0 && cand) atomicAdd(n_cand, (unsigned long long) cand);
  __syncthreads();
  resolve_flush(Q, B, s_defer, min(s_nd, (unsigned) kResolveDefer), out);
  lsi_tail(tail);
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
