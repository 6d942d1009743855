// C ABI of librjb200 (see include/rjb200.h).  Host-side orchestration of the
// CUDA kernels; no exceptions leave this file.
#include <math.h>
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "rjb_grid.cuh"
#include "rjb_lbvh.cuh"
#include <sched.h>

#include "rjb_lsi.cuh"
#include "rjb_pip.cuh"
#include "rjb_debug.cuh"

namespace rjb {

static thread_local std::string g_last_error;

struct DeviceMap {
  bool loaded = false;
  uint32_t n_points = 0, n_edges = 0, n_chains = 0;
  DBuf<double2> raw;
  DBuf<longlong2> pts;
  DBuf<uint32_t> edge_chain, point_chain, edge_desc, tile_desc, row_index, last_bits;
  // Morton order of the map's edges as query start points (maps of short chains): computed
  // by the first query that wants it, reused until the map is replaced
  DBuf<uint32_t> edge_order;
  bool edge_order_valid = false;
  DBuf<int32_t> left, right;
  std::vector<int32_t> h_left, h_right;
  DBuf<uint32_t> block_chain;  // chain of every 256th point (k_load_points)
  // pinned staging of the small per-chain arrays {left, right, block_chain}: pageable sources
  // would make every cudaMemcpyAsync of the upload synchronise the stream first
  int32_t* h_stage = nullptr;
  size_t h_stage_cap = 0;
  ~DeviceMap() {
    if (h_stage) cudaFreeHost(h_stage);
  }
  // host copy of the source graph kept for the overlay writer (points as
  // given, chains) -- WriteOutputChain reads ctx.get_planar_graph(im)
  std::vector<double> h_xy;
  std::vector<uint32_t> h_row_index;
  Bvh bvh;
  Grid grid;
  MapView view() const {
    MapView v;
    v.pts = pts.p;
    v.edge_chain = edge_chain.p;
    v.point_chain = point_chain.p;
    v.edge_desc = edge_desc.p;
    v.tile_desc = tile_desc.p;
    v.row_index = row_index.p;
    v.last_bits = last_bits.p;
    v.left = left.p;
    v.right = right.p;
    v.n_points = n_points;
    v.n_edges = n_edges;
    v.n_chains = n_chains;
    return v;
  }
};

// results of the last overlay run (xsect_edges_sorted_[2], closest_eids_[2],
// point_in_polygon_[2] of reference src/app/map_overlay.h:52-55)
struct OverlayState {
  bool done = false;
  uint64_t n_xsects = 0;
  uint32_t n_segs[2] = {0, 0};
  DBuf<rjb_xsect> xsects_sorted[2];
  DBuf<uint32_t> closest_eid[2];
  DBuf<int32_t> point_in_polygon[2];
  DBuf<uint32_t> seg_start[2];
  DBuf<uint64_t> keys_a, keys_b;
  DBuf<uint32_t> vals_a, vals_b, seg_flag, seg_scan;
  DBuf<longlong2> mid_pts;
  SortTemp sort_tmp;
  ScanTemp scan_tmp;
};

}  // namespace rjb

using namespace rjb;

constexpr int kLoadChunksMax = 16;
constexpr int kTimedStages = 4;
constexpr uint32_t kLoadChunkPoints = 1u << 20;  // 16 MB of double2 per chunk

// an LSI query that is on the stream and has not been waited for (rjb_lsi_launch)
struct LsiPending {
  bool active = false;
  int q = 0, mode = 0, attempt = 0;
  double xsect_factor = 0;
  uint32_t cap = 0, ccap = 0, n_slots = 0;
  bool lbvh = false, grid = false, filter = false, cells = false, fused = false;
  unsigned launches = 0;  // kernels this attempt put on the stream
};

struct rjb_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // map upload pipeline: the copy runs on `stream`, the per-chunk load kernel on `aux`
  cudaStream_t aux = nullptr;
  cudaEvent_t chunk_ev[kLoadChunksMax + 1] = {};
  unsigned long long* h_counters = nullptr;  // pinned + mapped: the counts a query reads back
  unsigned long long* d_h_counters = nullptr;  // the same buffer as the device sees it
  DBuf<unsigned int> lsi_ticket;             // k_lsi_resolve: CTAs that have finished (zero between queries)
  bool ctr_clean = false;                    // the device counters are zero (left so by k_lsi_resolve)
  int resolve_ctas_per_sm = 0;               // grid of k_lsi_resolve: 0 = what is resident at once (occupancy API)
  int resolve_resident = 0, resolve_resident_w = 0;
  int resolve_warp = 0;                      // k_lsi_resolve_w (warp-private lists) instead of k_lsi_resolve
  int pdl = 0;                               // programmatic dependent launch of the query's kernels: measured slower
  int fused = 1;                             // LBVH LSI: exact + point pass in one kernel (option lsi_fused)
  bool have_scaling = false;
  rjb_scaling sc;
  DeviceMap maps[2];
  int leaf_size = 4;
  // adaptive leaf grouping (RayJoin's -ag / -ag_iter / -enlarge, src/rt/primitive.h:120-260)
  int ag = 0, ag_iter = 5;
  float ag_enlarge = 5.0f;
  int sort_queries = -1;  // -1 auto: Morton-order query EDGES when the chains are short
  bool filter_useless = false;  // the occupancy filter kept > 50 % last time: skip it
  int stats = 0;  // collect traversal statistics (slower)
  int stage_timing = 1;  // LBVH LSI: an event after every kernel (rjb_last_stage_ms), else after every phase
  bool pip_park = true;  // PIP: park leaves per lane and open them together (rjb_pip.cuh)
  unsigned long long last_stats[8] = {0};
  LsiPending lsi_pending;
  unsigned last_launches = 0;  // kernels launched by the last completed query
  int keep_host_graph = 1;  // overlay writer needs the source coordinates
  // LSI result queue
  DBuf<uint2> pairs;
  DBuf<uint2> cands;      // LBVH traversal output: pairs whose exact boxes overlap
  DBuf<uint32_t> survivors;  // occupancy pre-filter output (query start points)
  int use_filter = -1;    // -1 auto (by occupancy), 0 off, 1 on
  uint32_t last_survivors = 0, last_long = 0;
  // survivors of the last filtered query per QUERY map (0 = not known): the two-level filter pays
  // off only where few edges survive (bench direction: 8 %; the overlay's direction: 23 %, where
  // it measured 80 us against 41 us for the one-level filter)
  uint32_t dir_survivors[2] = {0, 0};
  // the cell directory left long edges to a tree walk of their own last time (per query map): that
  // extra kernel costs more than the directory saves (0.130 vs 0.123 ms), so the walk takes everything
  bool dir_long[2] = {false, false};
  DBuf<uint32_t> long_edges;  // survivors longer than a cell (tree walk)
  uint32_t load_chunk = kLoadChunkPoints;  // points per upload chunk (option load_chunk_points)
  int pip_sort_bits = 24; // grid PIP: key bits the points are ordered by (the high ones: column first)
  int tile_filter = 1;    // LSI: two-level occupancy filter (tiles of 8 edges first, k_lsi_filter_tiles)
  int use_cells = 0;      // LSI: cell directory for the filter's survivors (experimental, off)
  size_t cand_cap = 0;
  size_t grid_work_cap = 0;
  // query window of rjb_lsi: only the query edges starting at points [win_begin, win_end) of the
  // query map (options lsi_window_begin / lsi_window_end; 0, 0 = the whole map).  A multi-GPU
  // driver keeps both maps whole on every rank and gives each rank a window of the query side.
  uint32_t win_begin = 0, win_end = 0;  // grid LSI: capacity of the (query edge, cell) work list
  DBuf<rjb_xsect> xsects;
  DBuf<unsigned long long> counters;  // [0] = queue counter (low 32 bits), [1] = candidates
  // PIP results
  DBuf<uint32_t> pip_eid;
  DBuf<int32_t> pip_face;
  DBuf<longlong2> pip_pts;
  DBuf<uint2> pip_packed;  // {eid, face} per point: the scattered write of ordered queries
  DBuf<double2> pip_raw;
  // query ordering scratch
  DBuf<uint64_t> ord_keys_a, ord_keys_b;
  DBuf<uint32_t> ord_vals_a, ord_vals_b;
  SortTemp ord_sort;
  // overlay state
  OverlayState ov;
  // timing of the kernels of the last query call: ev[k] .. ev[k + 1] brackets stage k
  // (LBVH LSI: filter, traversal, exact pass, point pass; other queries: query kernel,
  // point pass).  Index builds have events of their own (bev).
  cudaEvent_t ev[kTimedStages + 1] = {};
  cudaEvent_t bev[2] = {nullptr, nullptr};
  float last_ms[kTimedStages] = {};
  int timing_pending = 0;  // number of stages recorded by the last query and not read yet
  int timing_layout = 0;   // 0: {query kernel, point pass}; 1: {filter, traversal, exact, points}
};

namespace rjb {

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return RJB_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return RJB_ERR_INVALID;
  }
}

// ---- kernel of the load path -----------------------------------------------
// One thread per point of [p_begin, p_end):
//  * scaling exactly like the reference's device kernel (src/map/map.h:171-180 ->
//    src/map/scaling.h:79-95): fma.rn.f64 then cvt.rzi.s64.f64;
//  * the chain of the point and the edge numbering of src/map/map.h:200-207
//    (eid = p - chain), both directions;
//  * bit p of last_bits <=> p is the last point of its chain (the slot owns no edge);
//  * the occupancy descriptor of the edge ENDING at p (edge_desc[p - 1], what
//    k_lsi_filter streams): the previous vertex comes from the neighbour lane, or is
//    re-scaled from the raw array -- uploaded already, chunks arrive in order.
// p_begin is a multiple of 256: a CTA owns 256 consecutive points, a warp whole last_bits words.
//
// The chain of a point is found WITHOUT a search per point (a 20-step binary search over
// row_index for each of 9 M points made this kernel latency bound at 32 % of the HBM
// roofline): the host hands over the chain of every 256th point (block_chain, one merge pass
// over the chains while it validates them), the CTA stages the next 258 chain starts in
// shared memory, a warp finds the chain of its first point there and marks the chain starts
// among its 32 points in one 32-bit mask (__reduce_or_sync); chain = chain of the first
// point + popcount of the starts at or before the lane.
__global__ void __launch_bounds__(256)
k_load_points(const double2* __restrict__ in, uint32_t p_begin, uint32_t p_end, uint32_t n_points,
              double rx, double ry, double dx, double dy, const uint32_t* __restrict__ row_index,
              uint32_t n_chains, const uint32_t* __restrict__ block_chain, longlong2* __restrict__ out,
              uint32_t* __restrict__ edge_desc, uint32_t* __restrict__ tile_desc,
              uint32_t* __restrict__ point_chain, uint32_t* __restrict__ edge_chain,
              uint32_t* __restrict__ last_bits) {
  __shared__ uint32_t s_ri[258];
  const uint32_t pb0 = p_begin + blockIdx.x * 256;
  const uint32_t p = pb0 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint32_t c_block = __ldg(&block_chain[pb0 >> 8]);
  s_ri[threadIdx.x] = __ldg(&row_index[min(c_block + threadIdx.x, n_chains)]);
  if (threadIdx.x < 2) s_ri[256 + threadIdx.x] = __ldg(&row_index[min(c_block + 256 + threadIdx.x, n_chains)]);
  longlong2 o = make_longlong2(0, 0);
  if (p < p_end) {
    const double2 v = in[p];
    o.x = (long long) fma(v.x, rx, dx);
    o.y = (long long) fma(v.y, ry, dy);
    out[p] = o;
  }
  __syncthreads();
  const uint32_t pw = pb0 + (threadIdx.x & ~31u);  // first point of the warp
  // last staged chain whose first point is <= pw (<= 128 chains start inside a CTA: every
  // chain has at least 2 points); the same search in every lane, no divergence
  uint32_t lo = 0, hi = 130;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (s_ri[mid] <= pw) lo = mid; else hi = mid;
  }
  // chain starts among the warp's points (row_index[n_chains] = n_points acts as one more start)
  const uint32_t off = s_ri[lo + 1 + lane] - pw;
  const uint32_t starts = __reduce_or_sync(0xffffffffu, off < 32u ? 1u << off : 0u);
  const bool start_after = __any_sync(0xffffffffu, off == 32u);
  const uint32_t chain = c_block + lo + __popc(starts & (0xFFFFFFFFu >> (31 - lane)));
  const bool first = lane == 0 ? s_ri[lo] == pw : (starts >> lane) & 1u;
  const bool last = p < p_end && (lane < 31 ? (starts >> (lane + 1)) & 1u : start_after);
  if (p < p_end) {
    point_chain[p] = chain;
    if (!last) edge_chain[p - chain] = chain;
  }
  const unsigned m = __ballot_sync(0xffffffffu, last);
  if (lane == 0 && p < p_end) last_bits[p >> 5] = m;
  longlong2 q;  // vertex p - 1
  q.x = __shfl_up_sync(0xffffffffu, o.x, 1);
  q.y = __shfl_up_sync(0xffffffffu, o.y, 1);
  // cell box of the edge ending at p (empty when there is none), for the tile descriptor
  uint32_t bx0 = 0xFFFFFFFFu, by0 = 0xFFFFFFFFu, bx1 = 0, by1 = 0;
  if (p < p_end && p > 0) {
    uint32_t d = kDescNone << 24;
    if (!first) {
      if (lane == 0) {
        const double2 v = in[p - 1];
        q.x = (long long) fma(v.x, rx, dx);
        q.y = (long long) fma(v.y, ry, dy);
      }
      const uint32_t c1 = occ_code(q.x, q.y), c2 = occ_code(o.x, o.y);
      d = edge_desc_of(c1, c2);
      const uint32_t x1 = c1 & (kOccDim - 1), x2 = c2 & (kOccDim - 1), y1 = c1 >> kOccBits, y2 = c2 >> kOccBits;
      bx0 = min(x1, x2); bx1 = max(x1, x2);
      by0 = min(y1, y2); by1 = max(y1, y2);
    }
    edge_desc[p - 1] = d;
  }
  if (p == n_points - 1) edge_desc[p] = kDescNone << 24;
  // tile descriptors of the warp's 32 points (tile_desc_of): one word per group of kTileT lanes
  {
    const unsigned gmask = (kTileT == 32 ? 0xffffffffu : ((1u << kTileT) - 1u)) << (lane & ~(kTileT - 1));
    const bool any_edge = (__ballot_sync(0xffffffffu, bx0 != 0xFFFFFFFFu) & gmask) != 0;
    bx0 = __reduce_min_sync(gmask, bx0);
    by0 = __reduce_min_sync(gmask, by0);
    bx1 = __reduce_max_sync(gmask, bx1);
    by1 = __reduce_max_sync(gmask, by1);
    if ((lane & (kTileT - 1)) == 0 && p < p_end)
      tile_desc[p / kTileT] = any_edge ? tile_desc_of(bx0, by0, bx1, by1) : (kTileNone << 24);
    if (p == n_points - 1) tile_desc[p / kTileT + 1] = kTileNone << 24;  // the tile after the last one is read too
  }
}

// scaling alone (query points of rjb_pip_host): fma.rn.f64 then cvt.rzi.s64.f64
__global__ void k_scale_points(const double2* __restrict__ in, uint32_t n, double rx, double ry,
                               double dx, double dy, longlong2* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double2 p = in[i];
  longlong2 o;
  o.x = (long long) fma(p.x, rx, dx);
  o.y = (long long) fma(p.y, ry, dy);
  out[i] = o;
}

// Sort keys travel packed with their payload: key (the top 32 Morton bits) in the high half of a
// 64-bit word, payload in the low half (sort_packed, rjb_sort.cuh).
__global__ void k_query_keys_edges(MapView Q, long long imin, uint64_t* __restrict__ packed) {
  uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Q.n_edges) return;
  Seg s = load_seg(Q, e);
  const uint64_t key = morton64(s.x1 + ((s.x2 - s.x1) >> 1), s.y1 + ((s.y2 - s.y1) >> 1), imin);
  // payload: the edge's start point, what the traversal consumes
  packed[e] = (key & 0xFFFFFFFF00000000ull) | (uint64_t) (e + Q.edge_chain[e]);
}

__global__ void k_query_keys_points(const longlong2* __restrict__ pts, uint32_t n, long long imin,
                                    uint64_t* __restrict__ packed) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  longlong2 p = pts[i];
  packed[i] = (morton64(p.x, p.y, imin) & 0xFFFFFFFF00000000ull) | (uint64_t) i;
}

__global__ void k_unpack_low(const uint64_t* __restrict__ packed, uint32_t n, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t) packed[i];
}

static void scaling_init(rjb_scaling& s, double bminx, double bminy, double bmaxx, double bmaxy) {
  // src/map/scaling.h:43-46,56-71 (host arithmetic, no FMA contraction here:
  // every product feeds exactly one rounding because of the volatile temps)
  s.internal_max = INT64_MAX >> 17;
  s.internal_min = INT64_MIN >> 17;
  s.internal_range = s.internal_max - s.internal_min;
  double max_x = bmaxx + 1, min_x = bminx - 1, max_y = bmaxy + 1, min_y = bminy - 1;
  s.rx = (double) s.internal_range / (max_x - min_x);
  s.ry = (double) s.internal_range / (max_y - min_y);
  s.rrx = 1 / s.rx;
  s.rry = 1 / s.ry;
  int64_t isum = s.internal_max + s.internal_min;
  volatile double tx = (max_x + min_x) * s.rx, ty = (max_y + min_y) * s.ry;
  volatile double ux = isum * s.rrx, uy = isum * s.rry;
  s.deltax = 0.5 * (isum - tx);
  s.deltay = 0.5 * (isum - ty);
  s.ddeltax = 0.5 * ((max_x + min_x) - ux);
  s.ddeltay = 0.5 * ((max_y + min_y) - uy);
}

static void check_map_id(int id) { RJB_REQUIRE(id == 0 || id == 1, "map id must be 0 or 1"); }

static const uint32_t* query_order_edges(rjb_ctx* c, DeviceMap& Qm, const MapView& Q) {
  // Consecutive edges of a long chain are spatial neighbours already; maps made of
  // short chains in arbitrary order (polygon soups: ~7 edges per chain) are not, and a
  // warp would walk one cluster after the other.  auto = sort below 32 edges per chain.
  bool want = c->sort_queries > 0 ||
              (c->sort_queries < 0 && Q.n_chains > 0 && Q.n_edges / Q.n_chains < 32);
  if (!want || Q.n_edges == 0) return nullptr;
  if (Qm.edge_order_valid) return Qm.edge_order.p;
  uint32_t n = Q.n_edges;
  uint64_t* ka = c->ord_keys_a.ensure(n);
  uint64_t* kb = c->ord_keys_b.ensure(n);
  k_query_keys_edges<<<div_up(n, 256), 256, 0, c->stream>>>(Q, c->sc.internal_min, ka);
  // only coherence is needed: the top 24 Morton bits (4096 x 4096 cells) = 3 radix passes
  const uint64_t* sorted = sort_packed(ka, kb, n, 8, 32, c->ord_sort, c->stream);
  // the order belongs to the map, not to the query: keep it (the scratch buffers are
  // shared with the point ordering of PIP)
  uint32_t* keep = Qm.edge_order.ensure(n);
  k_unpack_low<<<div_up(n, 256), 256, 0, c->stream>>>(sorted, n, keep);
  Qm.edge_order_valid = true;
  return keep;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait in rjb_lsi.cuh)
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  RJB_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

// Wait for the stream by polling (yielding the core between polls): a blocking cudaStreamSynchronize wakes the host tens of
// microseconds late, which is a tenth of a whole LSI query.  Long waits fall back to it.
static void wait_stream(cudaStream_t s) {
  for (int spins = 0; spins < 20000; spins++) {
    const cudaError_t e = cudaStreamQuery(s);
    if (e == cudaSuccess) return;
    if (e != cudaErrorNotReady) RJB_CUDA(e);
    sched_yield();  // other ranks of the job may share this core
  }
  RJB_CUDA(cudaStreamSynchronize(s));
}

static void ensure_load_pipeline(rjb_ctx* c) {
  if (!c->aux) RJB_CUDA(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
  for (int i = 0; i <= kLoadChunksMax; i++)
    if (!c->chunk_ev[i]) RJB_CUDA(cudaEventCreateWithFlags(&c->chunk_ev[i], cudaEventDisableTiming));
}

static void ensure_events(rjb_ctx* c) {
  if (!c->h_counters) {
    RJB_CUDA(cudaHostAlloc((void**) &c->h_counters, 128 * sizeof(unsigned long long), cudaHostAllocMapped));
    RJB_CUDA(cudaHostGetDevicePointer((void**) &c->d_h_counters, c->h_counters, 0));
    unsigned int* t = c->lsi_ticket.ensure(1);
    RJB_CUDA(cudaMemsetAsync(t, 0, sizeof(unsigned int), c->stream));
    // CTAs of k_lsi_resolve that are resident at once (its grid is one wave)
    int a = 0, b = 0;
    RJB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_lsi_resolve<true>, kResolveThreads, 0));
    RJB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_lsi_resolve<false>, kResolveThreads, 0));
    c->resolve_resident = std::max(1, std::min(a, b));
    RJB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_lsi_resolve_w<true>, kResolveThreads, 0));
    RJB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_lsi_resolve_w<false>, kResolveThreads, 0));
    c->resolve_resident_w = std::max(1, std::min(a, b));
  }
  for (int i = 0; i <= kTimedStages; i++)
    if (!c->ev[i]) RJB_CUDA(cudaEventCreate(&c->ev[i]));
  for (int i = 0; i < 2; i++)
    if (!c->bev[i]) RJB_CUDA(cudaEventCreate(&c->bev[i]));
}

static void do_build_index(rjb_ctx* c, int map_id, int mode, uint32_t grid_size, double* build_ms) {
  check_map_id(map_id);
  DeviceMap& m = c->maps[map_id];
  RJB_REQUIRE(m.loaded, "rjb_build_index: map not loaded");
  ensure_events(c);
  RJB_CUDA(cudaEventRecord(c->bev[0], c->stream));
  if (mode == RJB_MODE_LBVH) {
    build_lbvh(m.bvh, m.view(), c->leaf_size, c->ag ? c->ag_iter : 0, c->ag_enlarge, c->sc.internal_min,
               c->use_cells > 0, c->stream);
  } else if (mode == RJB_MODE_GRID) {
    build_grid(m.grid, m.view(), grid_size, c->sc.internal_min, c->sc.internal_range, c->stream);
  } else if (mode == RJB_MODE_BRUTE) {
    // no index
  } else {
    throw Error(RJB_ERR_INVALID, "rjb_build_index: unknown mode");
  }
  RJB_CUDA(cudaEventRecord(c->bev[1], c->stream));
  RJB_CUDA(cudaEventSynchronize(c->bev[1]));
  float ms = 0;
  RJB_CUDA(cudaEventElapsedTime(&ms, c->bev[0], c->bev[1]));
  if (build_ms) *build_ms = ms;
}

// LSI into c->xsects, in two halves so that a caller can keep the device busy while its host
// thread does something else (rjb_lsi_launch / rjb_lsi_wait): lsi_enqueue puts the whole query
// on the stream -- kernels, and the read-back of the counters into pinned memory -- and returns;
// lsi_finish waits for the stream, and repeats the query if an internal queue was too small
// (the needed size is known exactly after the first attempt).  Between the kernels nothing
// goes through the host: the counts stay on the device.
static void lsi_enqueue(rjb_ctx* c, int q, int mode, double xsect_factor) {
  check_map_id(q);
  DeviceMap& Qm = c->maps[q];
  DeviceMap& Bm = c->maps[1 - q];
  RJB_REQUIRE(Qm.loaded && Bm.loaded, "rjb_lsi: both maps must be loaded");
  RJB_REQUIRE(mode == RJB_MODE_LBVH || mode == RJB_MODE_GRID || mode == RJB_MODE_BRUTE, "rjb_lsi: unknown mode");
  // queue capacity as in src/run_query.cu:226-228 (float arithmetic)
  float total_e = (float) ((size_t) Qm.n_edges + (size_t) Bm.n_edges);
  uint64_t cap64 = (uint64_t) (total_e * (float) xsect_factor);
  RJB_REQUIRE(cap64 < 0xFFFFFFFFull, "rjb_lsi: xsect queue exceeds 2^32 entries");
  uint32_t cap = (uint32_t) cap64;
  rjb_xsect* xs = c->xsects.ensure(cap ? cap : 1);
  // counters: [0] results, [1] candidates (= exact-predicate evaluations),
  // [2..7] traversal statistics
  // [8], [9]: {filter survivors, (query, leaf) pairs, long survivors} as 32-bit counters
  unsigned long long* ctr = c->counters.ensure(128);
  ensure_events(c);
  if (!(mode == RJB_MODE_LBVH && c->fused && !c->stats)) c->ctr_clean = false;  // the other paths leave their counts
  MapView Q = Qm.view(), B = Bm.view();
  LsiPending& P = c->lsi_pending;
  const int attempt = P.active ? P.attempt : 0;  // a retry keeps its attempt number
  P = LsiPending();
  P.q = q;
  P.mode = mode;
  P.xsect_factor = xsect_factor;
  P.cap = cap;
  P.attempt = attempt;
  // query window (start points); the whole map unless a window is set
  uint32_t p_lo = 0, p_hi = Q.n_points;
  if (c->win_end > 0) {
    RJB_REQUIRE(c->win_begin <= c->win_end && c->win_end <= Q.n_points, "rjb_lsi: query window outside the map");
    RJB_REQUIRE(mode != RJB_MODE_BRUTE, "rjb_lsi: the brute-force mode takes no query window");
    p_lo = c->win_begin;
    p_hi = c->win_end;
  }
  const bool windowed = p_lo != 0 || p_hi != Q.n_points;
  const bool nonempty = Q.n_edges > 0 && B.n_edges > 0 && p_hi > p_lo;
  if (mode == RJB_MODE_LBVH && !Bm.bvh.built) throw Error(RJB_ERR_NO_INDEX, "rjb_lsi: no LBVH on the base map");
  if (mode == RJB_MODE_GRID && nonempty && !Bm.grid.built)
    throw Error(RJB_ERR_NO_INDEX, "rjb_lsi: no grid on the base map");
  if (mode == RJB_MODE_LBVH && nonempty) {
    // traversal -> candidate pairs (exact integer boxes overlap) -> dense exact pass.
    // The candidate buffer is internal: it grows and the query is repeated if it was too small.
    if (c->cand_cap < (size_t) cap + 65536) c->cand_cap = 2 * (size_t) cap + 65536;
    const uint32_t* order = windowed ? nullptr : query_order_edges(c, Qm, Q);  // (a window walks in map order)
    // occupancy pre-filter: worthwhile when the base map covers a small part of the
    // plane; it replaces the Morton order (survivors come out in map order)
    const bool filter = !order && (c->use_filter == 1 || (c->use_filter < 0 && Bm.bvh.occ_fraction < 0.25 &&
                                                         !c->filter_useless));
    uint32_t* surv = filter ? c->survivors.ensure(Q.n_points) : nullptr;
    // [0] survivors, [1] (query, leaf) pairs, [2] survivors that are longer than a cell
    unsigned int* surv_n = (unsigned int*) (ctr + 8);
    // cell directory instead of the tree walk for the survivors (option lsi_cells)
    const bool cells = filter && Bm.bvh.have_cells && c->use_cells > 0 && !c->stats && Q.n_points < (1u << kDirectShift) &&
                       !(c->use_cells == 1 && c->dir_long[q]);  // (lsi_cells = 2: the directory regardless)
    uint32_t* long_list = cells ? c->long_edges.ensure(Q.n_points) : nullptr;
    RJB_REQUIRE(c->cand_cap < 0xFFFFFFF0ull, "rjb_lsi: candidate queue exceeds 2^32 entries");
    const uint32_t ccap = (uint32_t) c->cand_cap;
    uint2* cands = c->cands.ensure(ccap);
    // fused exact + point pass: its last CTA hands the counters to the host and zeroes them
    const bool fused = c->fused && !c->stats;
    if (!(fused && c->ctr_clean)) RJB_CUDA(cudaMemsetAsync(ctr, 0, 10 * sizeof(unsigned long long), c->stream));
    c->ctr_clean = false;  // until lsi_finish has seen this query complete
    const bool timed = c->stage_timing >= 0;  // (-1: no events at all: each costs 2-3 us of stream time)
    if (timed) RJB_CUDA(cudaEventRecord(c->ev[0], c->stream));
    // query slots: point indices (edge = slot, slot + 1), or a list of start points
    // (Morton-sorted edges, or the survivors of the occupancy filter, whose count
    // stays on the device)
    uint32_t n_slots = order ? Q.n_edges : p_hi;
    uint32_t slot_lo = order ? 0u : p_lo;  // plain slots: the window's first start point
    const uint32_t* slots = order;
    const unsigned int* n_slots_dev = nullptr;
    if (filter) {
      const bool tiles = c->tile_filter && (c->dir_survivors[q] ? (uint64_t) c->dir_survivors[q] * 8 < Q.n_edges
                                                                 : Bm.bvh.occ_fraction < 0.12);
      if (tiles)
        k_lsi_filter_tiles<<<div_up(p_hi / kTileT - p_lo / kTileT + 1, kTfCtaTiles), kTfWarps * 32, 0, c->stream>>>(
            Q, p_lo, p_hi, Bm.bvh.occ.p, surv, surv_n, long_list, surv_n + 2);
      else
        k_lsi_filter<<<div_up(p_hi - (p_lo & ~31u), kFilterCtaPoints), kFilterThreads, 0, c->stream>>>(
            Q, p_lo, p_hi, Bm.bvh.occ.p, surv, surv_n, long_list, surv_n + 2);
      slots = surv;
      n_slots_dev = surv_n;
      slot_lo = 0;
      // grid for the worst case; warps beyond the survivor count exit at once
      n_slots = c->last_survivors ? min(Q.n_points, c->last_survivors + c->last_survivors / 4 + 4096) : Q.n_points;
    }
    if (c->stage_timing > 0) RJB_CUDA(cudaEventRecord(c->ev[1], c->stream));
    if (cells) {
      // survivors: cell directory; the long ones (usually none) walk the tree
      if (c->pdl)
        launch_pdl(k_lsi_cells, kNumSMs * 48, kLsiWarps * 32, c->stream, Q, Bm.bvh.view(), (const uint32_t*) surv,
                   (const unsigned int*) surv_n, cands, ccap, surv_n + 1);
      else
        k_lsi_cells<<<kNumSMs * 48, kLsiWarps * 32, 0, c->stream>>>(Q, Bm.bvh.view(), surv, surv_n, cands, ccap,
                                                                  surv_n + 1);
      slots = long_list;
      n_slots_dev = surv_n + 2;
      slot_lo = 0;
      n_slots = c->last_long ? min(Q.n_points, c->last_long + c->last_long / 4 + 1024) : 0;
    }
    // the long-edge list comes from all over the map: two queries per warp while it is short
    const uint32_t spw = cells && n_slots <= 16384 ? 2 : 32;
    unsigned tiles = div_up(n_slots, spw) - (slots ? 0u : slot_lo / 32);
    unsigned blocks = div_up(tiles, kLsiWarps);
    if (blocks == 0) {
      // nothing to walk (no long edges last time; a non-empty list triggers the retry in lsi_finish)
    } else if (c->stats)
      k_lsi_bvh<true><<<blocks, kLsiWarps * 32, 0, c->stream>>>(
          Q, B, Bm.bvh.view(), slots, n_slots, slot_lo, n_slots_dev, spw, cands, ccap, surv_n + 1, ctr + 2, cells);
    else if (c->pdl)
      launch_pdl(k_lsi_bvh<false>, blocks, kLsiWarps * 32, c->stream, Q, B, Bm.bvh.view(), slots, n_slots, slot_lo,
                 n_slots_dev, spw, cands, ccap, surv_n + 1, ctr + 2, cells);
    else
      k_lsi_bvh<false><<<blocks, kLsiWarps * 32, 0, c->stream>>>(
          Q, B, Bm.bvh.view(), slots, n_slots, slot_lo, n_slots_dev, spw, cands, ccap, surv_n + 1, ctr + 2, cells);
    if (timed) RJB_CUDA(cudaEventRecord(c->ev[2], c->stream));
    if (fused) {
      const LsiTail tail = {ctr, c->d_h_counters, c->lsi_ticket.p};
      // one resident wave: every CTA ends with a gcd tail, a second wave would pay it twice
      const unsigned resolve_ctas = kNumSMs * (unsigned) (c->resolve_ctas_per_sm ? c->resolve_ctas_per_sm
                                                                                  : (c->resolve_warp ? c->resolve_resident_w : c->resolve_resident));
      if (!c->pdl) {
        if (c->resolve_warp) {  // warp-private lists, no CTA barrier before the end
          if (cells)
            k_lsi_resolve_w<true><<<resolve_ctas, kResolveThreads, 0, c->stream>>>(
                Q, B, q, cands, Bm.bvh.leaf_rec.p, surv_n + 1, ccap, xs, cap, (unsigned int*) ctr, ctr + 1, tail);
          else
            k_lsi_resolve_w<false><<<resolve_ctas, kResolveThreads, 0, c->stream>>>(
                Q, B, q, cands, Bm.bvh.leaf_rec.p, surv_n + 1, ccap, xs, cap, (unsigned int*) ctr, ctr + 1, tail);
        } else if (cells)  // pairs in the direct format of the cell directory
          k_lsi_resolve<true><<<resolve_ctas, kResolveThreads, 0, c->stream>>>(
              Q, B, q, cands, Bm.bvh.leaf_rec.p, surv_n + 1, ccap, xs, cap, (unsigned int*) ctr, ctr + 1, tail);
        else
          k_lsi_resolve<false><<<resolve_ctas, kResolveThreads, 0, c->stream>>>(
              Q, B, q, cands, Bm.bvh.leaf_rec.p, surv_n + 1, ccap, xs, cap, (unsigned int*) ctr, ctr + 1, tail);
      } else if (cells)
        launch_pdl(k_lsi_resolve<true>, resolve_ctas, kResolveThreads, c->stream, Q, B, q, (const uint2*) cands,
                   (const uint2*) Bm.bvh.leaf_rec.p, (const unsigned int*) (surv_n + 1), ccap, xs, cap,
                   (unsigned int*) ctr, ctr + 1, tail);
      else
        launch_pdl(k_lsi_resolve<false>, resolve_ctas, kResolveThreads, c->stream, Q, B, q, (const uint2*) cands,
                   (const uint2*) Bm.bvh.leaf_rec.p, (const unsigned int*) (surv_n + 1), ccap, xs, cap,
                   (unsigned int*) ctr, ctr + 1, tail);
      if (c->stage_timing > 0) RJB_CUDA(cudaEventRecord(c->ev[3], c->stream));
    } else {
      if (cells)  // pairs in the direct format of the cell directory
        k_lsi_exact<true><<<kNumSMs * 4, kExactThreads, 0, c->stream>>>(Q, B, cands, Bm.bvh.leaf_rec.p, surv_n + 1, ccap,
                                                                       xs, cap, (unsigned int*) ctr, ctr + 1);
      else
        k_lsi_exact<false><<<kNumSMs * 4, kExactThreads, 0, c->stream>>>(Q, B, cands, Bm.bvh.leaf_rec.p, surv_n + 1, ccap,
                                                                        xs, cap, (unsigned int*) ctr, ctr + 1);
      if (c->stage_timing > 0) RJB_CUDA(cudaEventRecord(c->ev[3], c->stream));
      k_lsi_points<<<kNumSMs * 3, kPointsThreads, 0, c->stream>>>(Q, B, q, (const unsigned int*) ctr, cap, xs, false);
    }
    if (timed) RJB_CUDA(cudaEventRecord(c->ev[4], c->stream));
    RJB_CUDA(cudaGetLastError());
    P.lbvh = true;
    P.fused = fused;
    P.filter = filter;
    P.cells = cells;
    P.ccap = ccap;
    P.n_slots = n_slots;
    P.launches = (fused ? 2 : 3) + (filter ? 1 : 0) + (cells ? 1 : 0) - (blocks == 0 ? 1 : 0);
    c->timing_pending = timed ? 4 : 0;
    c->timing_layout = c->stage_timing > 0 ? 1 : 2;
    if (!timed)
      for (int k = 0; k < kTimedStages; k++) c->last_ms[k] = 0;
  } else if (mode == RJB_MODE_GRID && nonempty) {
    // cell filter -> (query edge, occupied cell) work items -> dense exact pass -> point pass.
    // Reference semantics of LSIGrid: intersect_test(map-0 edge, map-1 edge) whatever the query
    // side, a pair kept iff the cell of its intersection point holds both edges.
    if (c->grid_work_cap < 65536) c->grid_work_cap = (size_t) Q.n_edges / 4 + 65536;
    RJB_REQUIRE(c->grid_work_cap < 0xFFFFFFF0ull, "rjb_lsi: work list exceeds 2^32 entries");
    const uint32_t wcap = (uint32_t) c->grid_work_cap;
    uint2* work = c->cands.ensure(wcap);
    uint32_t* big = c->long_edges.ensure(Q.n_points);
    unsigned int* wn = (unsigned int*) (ctr + 8);  // [0] work items, [2] long query edges
    const GridView gv = Bm.grid.view();
    RJB_CUDA(cudaMemsetAsync(ctr, 0, 10 * sizeof(unsigned long long), c->stream));
    RJB_CUDA(cudaEventRecord(c->ev[0], c->stream));
    // maps of short chains in arbitrary order (polygon soups) are walked in Morton order, so that
    // neighbouring work items share cells and base vertices (as in the LBVH path)
    const uint32_t* gorder = windowed ? nullptr : query_order_edges(c, Qm, Q);
    if (gorder)
      k_grid_lsi_filter<<<div_up(Q.n_edges, 256), 256, 0, c->stream>>>(Q, gorder, 0, Q.n_edges, gv, work, wcap, wn, big,
                                                                      wn + 2);
    else
      k_grid_lsi_filter<<<div_up(p_hi - (p_lo & ~31u), 256), 256, 0, c->stream>>>(Q, nullptr, p_lo, p_hi, gv, work,
                                                                                 wcap, wn, big, wn + 2);
    k_grid_lsi_big<<<kNumSMs, 256, 0, c->stream>>>(Q, gv, big, wn + 2, work, wcap, wn);
    RJB_CUDA(cudaEventRecord(c->ev[1], c->stream));
    k_grid_lsi_exact<<<kNumSMs * 4, kExactThreads, 0, c->stream>>>(Q, B, gv, q, work, wn, wcap, xs, cap,
                                                                 (unsigned int*) ctr, ctr + 1);
    RJB_CUDA(cudaEventRecord(c->ev[2], c->stream));
    k_lsi_points<<<kNumSMs * 3, kPointsThreads, 0, c->stream>>>(Q, B, q, (const unsigned int*) ctr, cap, xs, q == 1);
    RJB_CUDA(cudaEventRecord(c->ev[3], c->stream));
    RJB_CUDA(cudaGetLastError());
    P.grid = true;
    P.ccap = wcap;
    P.launches = 4;
    c->timing_pending = 3;
    c->timing_layout = 3;
  } else {
    RJB_CUDA(cudaMemsetAsync(ctr, 0, 10 * sizeof(unsigned long long), c->stream));
    RJB_CUDA(cudaEventRecord(c->ev[0], c->stream));
    uint2* pairs = c->pairs.ensure(cap ? cap : 1);
    if (nonempty && mode == RJB_MODE_BRUTE) {
      dim3 g(div_up(Q.n_edges, 256), min(64u, div_up(B.n_edges, 256)));
      k_lsi_brute<<<g, 256, 0, c->stream>>>(Q, B, pairs, cap, (unsigned int*) ctr, ctr + 1);
      P.launches++;
    }
    RJB_CUDA(cudaEventRecord(c->ev[1], c->stream));
    // the point pass reads the count on the device: no host round trip between
    // the two kernels
    if (cap > 0 && nonempty) {
      k_xsect_points_dyn<<<div_up(cap, 128), 128, 0, c->stream>>>(
          Q, B, q, pairs, (const unsigned int*) ctr, cap, xs);
      P.launches++;
    }
    RJB_CUDA(cudaEventRecord(c->ev[2], c->stream));
    RJB_CUDA(cudaGetLastError());
    c->timing_pending = 2;  // event times are fetched when rjb_last_kernel_ms asks for them
    c->timing_layout = 0;
  }
  // one read-back into pinned memory: the only host round trip of the query (the fused LBVH
  // pipeline writes it from its last kernel)
  if (!P.fused)
    RJB_CUDA(cudaMemcpyAsync(c->h_counters, ctr, 10 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             c->stream));
  P.active = true;
}

// Returns the number of intersections found.
static uint64_t lsi_finish(rjb_ctx* c, uint64_t* n_candidates) {
  LsiPending& P = c->lsi_pending;
  RJB_REQUIRE(P.active, "rjb_lsi_wait: no query was launched");
  unsigned long long h[8];
  unsigned launches = 0;
  while (true) {
    try {
      wait_stream(c->stream);
    } catch (...) {
      P.active = false;
      throw;
    }
    launches += P.launches;
    if (P.fused) c->ctr_clean = true;  // zeroed by the last CTA of k_lsi_resolve
    memcpy(h, c->h_counters, sizeof(h));
    unsigned int hs[4];
    memcpy(hs, c->h_counters + 8, sizeof(hs));
    if (P.grid) {
      h[2] = hs[0];  // (query edge, occupied cell) work items
      h[6] = hs[2];  // query edges longer than kGridSmallCells cells
      if (hs[0] <= P.ccap) break;
      if (P.attempt >= 2) {
        P.active = false;
        throw Error(RJB_ERR_INVALID, "rjb_lsi: internal queues overflowed repeatedly");
      }
      c->grid_work_cap = (size_t) hs[0] + hs[0] / 8 + 65536;
      P.attempt++;
      const int q = P.q, mode = P.mode;
      const double xf = P.xsect_factor;
      try {
        lsi_enqueue(c, q, mode, xf);
      } catch (...) {
        P.active = false;
        throw;
      }
      continue;
    }
    if (!P.lbvh) break;
    // launches sized from the last query
    const bool grid_too_small = P.cells ? hs[2] > P.n_slots : (P.filter && hs[0] > P.n_slots);
    if (P.cells) {
      c->last_long = hs[2];
      if (hs[2]) c->dir_long[P.q] = true;
    }
    if (P.filter) {
      c->last_survivors = hs[0];
      c->dir_survivors[P.q] = hs[0] ? hs[0] : 1;
      // adaptive: not worth a pass over S
      c->filter_useless = hs[0] > c->maps[P.q].n_edges / 2;
    }
    if (!c->stats) h[2] = hs[1];  // (query, leaf) pairs the traversal handed to the exact pass
    if (P.filter) h[7] = c->last_survivors;
    if (P.cells) {  // which path produced the candidates (no traversal statistics in this mode)
      h[5] = 1;
      h[6] = c->last_long;
    }
    if (hs[1] <= P.ccap && !grid_too_small) break;
    if (P.attempt >= 2) {
      P.active = false;
      throw Error(RJB_ERR_INVALID, "rjb_lsi: internal queues overflowed repeatedly");
    }
    if (hs[1] > P.ccap) c->cand_cap = (size_t) hs[1] + hs[1] / 8 + 65536;
    P.attempt++;
    const int q = P.q, mode = P.mode;
    const double xf = P.xsect_factor;
    try {
      lsi_enqueue(c, q, mode, xf);
    } catch (...) {
      P.active = false;
      throw;
    }
  }
  const uint32_t cap = P.cap;
  P.active = false;
  P.attempt = 0;
  c->last_launches = launches;
  memcpy(c->last_stats, h, sizeof(h));
  uint64_t n = (uint32_t) h[0];
  if (n_candidates) *n_candidates = h[1];
  if (n > cap) {
    throw Error(RJB_ERR_QUEUE_OVERFLOW,
                "rjb_lsi: " + std::to_string(n) + " intersections exceed the queue capacity " +
                    std::to_string(cap) + " (raise -xsect_factor)");
  }
  return n;
}

static uint64_t do_lsi(rjb_ctx* c, int q, int mode, double xsect_factor, uint64_t* n_candidates) {
  c->lsi_pending.active = false;  // an abandoned launch is superseded
  lsi_enqueue(c, q, mode, xsect_factor);
  return lsi_finish(c, n_candidates);
}

// order of the query points for coherent warps: key = the 4096 x 4096 cell of the point in
// Morton order (LBVH) or its grid cell in the grid's own column-major numbering (grid: a warp
// then walks one column together).  Only coherence is needed, not a total order.
__global__ void k_query_keys_points_grid(const longlong2* __restrict__ pts, uint32_t n, GridView g,
                                         uint64_t* __restrict__ packed) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const longlong2 p = pts[i];
  const uint64_t key = (uint64_t) grid_cx(g, p.x) * g.gs + (uint64_t) grid_cy(g, p.y - 1);  // < 2^32
  packed[i] = (key << 32) | (uint64_t) i;
}

// `user_points`: the caller handed its own points (not the vertices of a map, which are
// coherent along their chains): with sort_queries = -1 (auto) those are ordered when there are
// enough of them to pay for the sort.
static void do_pip(rjb_ctx* c, int q, int mode, const longlong2* d_pts, uint32_t n,
                   uint64_t* n_candidates, bool user_points = false) {
  check_map_id(q);
  DeviceMap& Bm = c->maps[1 - q];
  RJB_REQUIRE(Bm.loaded, "rjb_pip: base map not loaded");
  RJB_REQUIRE(mode == RJB_MODE_LBVH || mode == RJB_MODE_GRID || mode == RJB_MODE_BRUTE, "rjb_pip: unknown mode");
  uint32_t* eid = c->pip_eid.ensure(n ? n : 1);
  int32_t* face = c->pip_face.ensure(n ? n : 1);
  // [0..7] statistics, [16 .. 16 + kCtrSlots) partial candidate counts
  constexpr int kPipCtrs = 16 + kCtrSlots;
  unsigned long long* ctr = c->counters.ensure(kPipCtrs);
  ensure_events(c);
  c->ctr_clean = false;
  RJB_CUDA(cudaMemsetAsync(ctr, 0, kPipCtrs * sizeof(unsigned long long), c->stream));
  MapView B = Bm.view();
  if (mode == RJB_MODE_LBVH && !Bm.bvh.built) throw Error(RJB_ERR_NO_INDEX, "rjb_pip: no LBVH on the base map");
  if (mode == RJB_MODE_GRID && !Bm.grid.built) throw Error(RJB_ERR_NO_INDEX, "rjb_pip: no grid on the base map");
  unsigned launches = 0;
  RJB_CUDA(cudaEventRecord(c->ev[0], c->stream));
  const uint64_t* order = nullptr;  // packed: the point index is the low half of each word
  const bool sort = n > 0 && mode != RJB_MODE_BRUTE && B.n_edges > 0 &&
                    (c->sort_queries > 0 || (c->sort_queries < 0 && user_points && n >= 65536));
  if (sort) {
    uint64_t* ka = c->ord_keys_a.ensure(n);
    uint64_t* kb = c->ord_keys_b.ensure(n);
    if (mode == RJB_MODE_GRID) {
      const GridView gv = Bm.grid.view();
      RJB_REQUIRE((uint64_t) gv.gx * gv.gs <= (1ull << 32), "grid too large to order the points by cell");
      k_query_keys_points_grid<<<div_up(n, 256), 256, 0, c->stream>>>(d_pts, n, gv, ka);
      int bits = 1;
      while (bits < 32 && (((uint64_t) gv.gx * gv.gs) >> bits)) bits++;
      // the low (row) bits beyond pip_sort_bits (default 24) key bits buy nothing: three radix
      // passes at most
      const int lo = bits > c->pip_sort_bits ? bits - c->pip_sort_bits : 0;
      order = sort_packed(ka, kb, n, lo, bits, c->ord_sort, c->stream);
      launches += 3 + (bits - lo + 7) / 8;
    } else {
      k_query_keys_points<<<div_up(n, 256), 256, 0, c->stream>>>(d_pts, n, c->sc.internal_min, ka);
      // the top 24 Morton bits (4096 x 4096 cells) = 3 radix passes instead of 8
      order = sort_packed(ka, kb, n, 8, 32, c->ord_sort, c->stream);
      launches += 3 + 3;
    }
  }
  RJB_CUDA(cudaEventRecord(c->ev[1], c->stream));
  // ordered queries write their results through ONE scattered 8-byte store per point
  uint2* packed = order ? c->pip_packed.ensure(n) : nullptr;
  if (n > 0) {
    if (mode == RJB_MODE_LBVH) {
      const unsigned pb = div_up(n, kLsiWarps * 32), pt = kLsiWarps * 32;
      if (c->stats)
        k_pip_bvh<true, false><<<pb, pt, 0, c->stream>>>(d_pts, n, order, B, Bm.bvh.view(), q, eid, face, packed, ctr);
      else if (c->pip_park)
        k_pip_bvh<false, true><<<pb, pt, 0, c->stream>>>(d_pts, n, order, B, Bm.bvh.view(), q, eid, face, packed, ctr);
      else
        k_pip_bvh<false, false><<<pb, pt, 0, c->stream>>>(d_pts, n, order, B, Bm.bvh.view(), q, eid, face, packed, ctr);
    } else if (mode == RJB_MODE_GRID) {
      const GridView gv = Bm.grid.view();
      if (packed)
        k_pip_grid<true><<<div_up(n, 256), 256, 0, c->stream>>>(d_pts, n, order, B, gv, q, eid, face, packed, ctr + 16);
      else
        k_pip_grid<false><<<div_up(n, 256), 256, 0, c->stream>>>(d_pts, n, order, B, gv, q, eid, face, nullptr, ctr + 16);
    } else {
      k_pip_brute<<<div_up(n, 256), 256, 0, c->stream>>>(d_pts, n, B, q, eid, face);
    }
    launches++;
    if (packed) {
      k_pip_split<<<div_up(n, 256), 256, 0, c->stream>>>(packed, n, eid, face);
      launches++;
    }
  }
  RJB_CUDA(cudaEventRecord(c->ev[2], c->stream));
  RJB_CUDA(cudaGetLastError());
  RJB_CUDA(cudaMemcpyAsync(c->h_counters, ctr, kPipCtrs * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           c->stream));
  RJB_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(c->last_stats, c->h_counters, 8 * sizeof(unsigned long long));
  for (int k = 0; k < kCtrSlots; k++) c->last_stats[1] += c->h_counters[16 + k];
  c->last_launches = launches;
  c->timing_pending = 2;
  c->timing_layout = 4;  // {ordering of the points, query kernel (+ result split)}
  if (n_candidates) *n_candidates = c->last_stats[1];
}

}  // namespace rjb

#include "rjb_overlay.cuh"

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* rjb_last_error(void) { return g_last_error.c_str(); }
void rjb__set_error(const char* msg) { g_last_error = msg ? msg : ""; }
const char* rjb_version(void) { return "rjb200 0.1 (sm_100a)"; }

#ifdef RJB_TRACE
// development builds only (-DRJB_TRACE): the phase trace of the last k_lsi_resolve launch
int rjb_debug_trace(rjb_ctx* c, unsigned long long* out, int n_words) {
  return guarded([&] {
    RJB_CUDA(cudaStreamSynchronize(c->stream));
    if (n_words > 4096 * 16)  // the per-warp trace
      RJB_CUDA(cudaMemcpyFromSymbol(out, rjb::g_trace_w, (size_t) n_words * sizeof(unsigned long long)));
    else
      RJB_CUDA(cudaMemcpyFromSymbol(out, rjb::g_trace, (size_t) n_words * sizeof(unsigned long long)));
  });
}
#endif

int rjb_create(int device, rjb_ctx** out) {
  return guarded([&] {
    RJB_REQUIRE(out != nullptr, "rjb_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      throw Error(RJB_ERR_CUDA, std::string("rjb_create: no CUDA device (") +
                                    cudaGetErrorString(e) + "); there is no CPU fallback");
    RJB_REQUIRE(device >= 0 && device < n, "rjb_create: bad device ordinal");
    RJB_CUDA(cudaSetDevice(device));
    std::unique_ptr<rjb_ctx> c(new rjb_ctx());
    c->device = device;
    RJB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    *out = c.release();
  });
}

void rjb_destroy(rjb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i <= kTimedStages; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 2; i++)
    if (c->bev[i]) cudaEventDestroy(c->bev[i]);
  for (int i = 0; i <= kLoadChunksMax; i++)
    if (c->chunk_ev[i]) cudaEventDestroy(c->chunk_ev[i]);
  if (c->aux) cudaStreamDestroy(c->aux);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int rjb_set_stream(rjb_ctx* c, void* s) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t) s : c->own_stream;
  });
}

int rjb_set_bounding_box(rjb_ctx* c, double min_x, double min_y, double max_x, double max_y) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_REQUIRE(min_x <= max_x && min_y <= max_y, "rjb_set_bounding_box: empty box");
    scaling_init(c->sc, min_x, min_y, max_x, max_y);
    c->have_scaling = true;
    // scaled coordinates of loaded maps are stale now, and so is every result derived from them
    c->ov.done = false;
    c->ov.n_xsects = 0;
    for (auto& m : c->maps) {
      m.loaded = false;
      m.edge_order_valid = false;
      m.bvh.built = false;
      m.grid.built = false;
    }
  });
}

int rjb_get_scaling(const rjb_ctx* c, rjb_scaling* out) {
  return guarded([&] {
    RJB_REQUIRE(c && out, "NULL argument");
    RJB_REQUIRE(c->have_scaling, "rjb_get_scaling: call rjb_set_bounding_box first");
    *out = c->sc;
  });
}

int rjb_set_option(rjb_ctx* c, const char* name, int64_t value) {
  return guarded([&] {
    RJB_REQUIRE(c && name, "NULL argument");
    std::string n(name);
    if (n == "lbvh_leaf_size") {
      RJB_REQUIRE(value >= 1 && value <= 8, "lbvh_leaf_size must be in 1..8");
      c->leaf_size = (int) value;
    } else if (n == "lsi_window_begin") {
      RJB_REQUIRE(value >= 0 && value < 0xFFFFFFF0ll, "lsi_window_begin out of range");
      c->win_begin = (uint32_t) value;
    } else if (n == "lsi_window_end") {
      RJB_REQUIRE(value >= 0 && value < 0xFFFFFFF0ll, "lsi_window_end out of range");
      c->win_end = (uint32_t) value;
    } else if (n == "lbvh_ag") {
      c->ag = value != 0;
    } else if (n == "lbvh_ag_iter") {
      RJB_REQUIRE(value >= 1 && value <= 64, "lbvh_ag_iter must be in 1..64");
      c->ag_iter = (int) value;
    } else if (n == "lbvh_enlarge_x1000") {
      RJB_REQUIRE(value >= 1000 && value <= 1000000000ll, "lbvh_enlarge_x1000 must be >= 1000");
      c->ag_enlarge = (float) value / 1000.0f;
    } else if (n == "sort_queries") {
      c->sort_queries = (int) value;
    } else if (n == "lsi_filter") {
      c->use_filter = (int) value;
    } else if (n == "load_chunk_points") {
      RJB_REQUIRE(value >= 1024 && value <= (1ll << 30) && value % 1024 == 0,
                  "load_chunk_points must be a multiple of 1024");
      c->load_chunk = (uint32_t) value;
    } else if (n == "lsi_cells") {
      c->use_cells = (int) value;
    } else if (n == "lsi_tile_filter") {
      c->tile_filter = value != 0;
    } else if (n == "pip_sort_bits") {
      RJB_REQUIRE(value >= 1 && value <= 32, "pip_sort_bits must be in 1..32");
      c->pip_sort_bits = (int) value;
    } else if (n == "lsi_resolve_ctas") {
      RJB_REQUIRE(value >= 0 && value <= 64, "lsi_resolve_ctas must be in 0..64");
      c->resolve_ctas_per_sm = (int) value;
    } else if (n == "lsi_resolve_warp") {
      c->resolve_warp = value != 0;
    } else if (n == "lsi_pdl") {
#ifdef RJB_PDL
      c->pdl = value != 0;
#else
      RJB_REQUIRE(value == 0, "lsi_pdl: this build has no programmatic dependent launch (-DRJB_PDL)");
#endif
    } else if (n == "lsi_fused") {
      c->fused = value != 0;
    } else if (n == "pip_park") {
      c->pip_park = value != 0;
    } else if (n == "stage_timing") {
      RJB_REQUIRE(value >= -1 && value <= 1, "stage_timing must be -1, 0 or 1");
      c->stage_timing = (int) value;
    } else if (n == "stats") {
      c->stats = value != 0;
    } else if (n == "keep_host_graph") {
      c->keep_host_graph = value != 0;
    } else {
      throw Error(RJB_ERR_INVALID, "rjb_set_option: unknown option " + n);
    }
  });
}

int rjb_set_map(rjb_ctx* c, int map_id, const double* xy, uint64_t n_points,
                const uint32_t* row_index, const int64_t* left, const int64_t* right,
                uint64_t n_chains) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    check_map_id(map_id);
    RJB_REQUIRE(c->have_scaling, "rjb_set_map: call rjb_set_bounding_box first");
    RJB_REQUIRE(n_points < 0xFFFFFFF0ull, "rjb_set_map: index_t is 32-bit (src/config.h:12)");
    RJB_REQUIRE(n_chains == 0 || (xy && row_index && left && right), "rjb_set_map: NULL array");
    RJB_CUDA(cudaSetDevice(c->device));
    DeviceMap& m = c->maps[map_id];
    m.loaded = false;
    m.edge_order_valid = false;
    m.bvh.built = false;
    m.grid.built = false;
    c->ov.done = false;  // overlay results index the OLD map's edges and points
    c->ov.n_xsects = 0;
    c->filter_useless = false;  // new data: let the occupancy filter prove itself again
    c->dir_survivors[0] = c->dir_survivors[1] = 0;
    c->dir_long[0] = c->dir_long[1] = false;
    if (n_chains > 0) {
      RJB_REQUIRE(row_index[0] == 0 && row_index[n_chains] == n_points,
                  "rjb_set_map: row_index must span [0, n_points]");
    } else {
      RJB_REQUIRE(n_points == 0, "rjb_set_map: points without chains");
    }
    m.n_points = (uint32_t) n_points;
    m.n_chains = (uint32_t) n_chains;
    m.n_edges = (uint32_t) (n_points - n_chains);
    m.h_left.resize(n_chains);
    m.h_right.resize(n_chains);
    // chain of every 256th point, for the load kernel (one merge pass over the chains)
    const uint64_t n_blocks = (n_points + 255) / 256;
    const size_t stage_n = 2 * n_chains + n_blocks + 1;
    if (m.h_stage_cap < stage_n) {
      if (m.h_stage) cudaFreeHost(m.h_stage);
      m.h_stage = nullptr;
      m.h_stage_cap = 0;
      RJB_CUDA(cudaHostAlloc((void**) &m.h_stage, (stage_n + stage_n / 8) * sizeof(int32_t), cudaHostAllocDefault));
      m.h_stage_cap = stage_n + stage_n / 8;
    }
    int32_t* h_l = m.h_stage;
    int32_t* h_r = m.h_stage + n_chains;
    uint32_t* h_bc = (uint32_t*) (m.h_stage + 2 * n_chains);
    uint64_t b = 0;
    for (uint64_t i = 0; i < n_chains; i++) {
      RJB_REQUIRE(row_index[i + 1] >= row_index[i] + 2, "rjb_set_map: chain with < 2 points");
      m.h_left[i] = h_l[i] = (int32_t) left[i];
      m.h_right[i] = h_r[i] = (int32_t) right[i];
      for (; b < n_blocks && 256 * b < row_index[i + 1]; b++) h_bc[b] = (uint32_t) i;
    }
    if (c->keep_host_graph) {
      m.h_xy.assign(xy, xy + 2 * n_points);
      m.h_row_index.assign(row_index, row_index + (n_chains ? n_chains + 1 : 0));
    } else {
      m.h_xy.clear();
      m.h_row_index.clear();
    }
    double2* raw = m.raw.ensure(n_points ? n_points : 1);
    longlong2* pts = m.pts.ensure(n_points ? n_points + 1 : 1);
    uint32_t* ri = m.row_index.ensure(n_chains + 1);
    int32_t* l = m.left.ensure(n_chains ? n_chains : 1);
    int32_t* r = m.right.ensure(n_chains ? n_chains : 1);
    uint32_t* ec = m.edge_chain.ensure(m.n_edges ? m.n_edges : 1);
    uint32_t* pc = m.point_chain.ensure(n_points ? n_points : 1);
    // padded with "no edge" (0xFF bytes: class bits = kDescNone): the filter reads 16 per thread
    uint32_t* cc = m.edge_desc.ensure(n_points + 16);
    uint32_t* td = m.tile_desc.ensure(n_points / kTileT + 2);
    uint32_t n_words = (uint32_t) (n_points / 32 + 2);
    uint32_t* lb = m.last_bits.ensure(n_words);
    uint32_t* bc = m.block_chain.ensure(n_blocks ? n_blocks : 1);
    cudaStream_t st = c->stream;
    if (n_points) {
      // Small arrays first, then the vertices in chunks: the copy engine streams chunk
      // k+1 on `stream` while the load kernel works on chunk k on `aux`, so scaling and
      // edge numbering hide behind the PCIe transfer (which bounds the whole upload).
      RJB_CUDA(cudaMemcpyAsync(ri, row_index, (n_chains + 1) * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, st));
      RJB_CUDA(cudaMemcpyAsync(l, h_l, n_chains * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      RJB_CUDA(cudaMemcpyAsync(r, h_r, n_chains * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      RJB_CUDA(cudaMemcpyAsync(bc, h_bc, n_blocks * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
      RJB_CUDA(cudaMemsetAsync(lb + (n_words - 2), 0, 2 * sizeof(uint32_t), st));
      RJB_CUDA(cudaMemsetAsync(cc + n_points, 0xFF, 16 * sizeof(uint32_t), st));
      ensure_load_pipeline(c);
      uint32_t chunk = c->load_chunk;
      if (div_up(n_points, chunk) > (unsigned) kLoadChunksMax)
        chunk = (div_up(n_points, kLoadChunksMax) + 1023u) & ~1023u;
      int k = 0;
      for (uint64_t p0 = 0; p0 < n_points; p0 += chunk, k++) {
        const uint32_t p1 = (uint32_t) std::min<uint64_t>(n_points, p0 + chunk);
        RJB_CUDA(cudaMemcpyAsync(raw + p0, xy + 2 * p0, (size_t) (p1 - p0) * sizeof(double2),
                                 cudaMemcpyHostToDevice, st));
        RJB_CUDA(cudaEventRecord(c->chunk_ev[k], st));
        RJB_CUDA(cudaStreamWaitEvent(c->aux, c->chunk_ev[k], 0));
        k_load_points<<<div_up(p1 - p0, 256), 256, 0, c->aux>>>(
            raw, (uint32_t) p0, p1, m.n_points, c->sc.rx, c->sc.ry, c->sc.deltax, c->sc.deltay, ri,
            m.n_chains, bc, pts, cc, td, pc, ec, lb);
      }
      RJB_CUDA(cudaGetLastError());
      RJB_CUDA(cudaEventRecord(c->chunk_ev[kLoadChunksMax], c->aux));
      RJB_CUDA(cudaStreamWaitEvent(st, c->chunk_ev[kLoadChunksMax], 0));
    }
    wait_stream(st);
    m.loaded = true;
  });
}

int rjb_map_info(const rjb_ctx* c, int map_id, uint64_t out[3]) {
  return guarded([&] {
    RJB_REQUIRE(c && out, "NULL argument");
    check_map_id(map_id);
    const DeviceMap& m = c->maps[map_id];
    RJB_REQUIRE(m.loaded, "rjb_map_info: map not loaded");
    out[0] = m.n_points;
    out[1] = m.n_edges;
    out[2] = m.n_chains;
  });
}

int rjb_map_device_views(const rjb_ctx* c, int map_id, const int64_t** d_points_xy,
                         const uint32_t** d_edge_chain) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    check_map_id(map_id);
    const DeviceMap& m = c->maps[map_id];
    RJB_REQUIRE(m.loaded, "rjb_map_device_views: map not loaded");
    if (d_points_xy) *d_points_xy = (const int64_t*) m.pts.p;
    if (d_edge_chain) *d_edge_chain = m.edge_chain.p;
  });
}

int rjb_build_index(rjb_ctx* c, int map_id, int mode, uint32_t grid_size, double* build_ms) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_CUDA(cudaSetDevice(c->device));
    do_build_index(c, map_id, mode, grid_size, build_ms);
  });
}

int rjb_lsi(rjb_ctx* c, int query_map_id, int mode, double xsect_factor,
            const rjb_xsect** d_xsects, uint64_t* n_xsects, uint64_t* n_candidates) {
  if (n_xsects) *n_xsects = 0;
  if (d_xsects) *d_xsects = nullptr;
  uint64_t needed = 0;
  int rc = guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_CUDA(cudaSetDevice(c->device));
    try {
      needed = do_lsi(c, query_map_id, mode, xsect_factor, n_candidates);
    } catch (const Error& e) {
      if (e.code == RJB_ERR_QUEUE_OVERFLOW) {
        // report how many were needed
        needed = (uint32_t) c->last_stats[0];
      }
      throw;
    }
    if (d_xsects) *d_xsects = c->xsects.p;
  });
  if (n_xsects) *n_xsects = needed;
  return rc;
}

int rjb_lsi_launch(rjb_ctx* c, int query_map_id, int mode, double xsect_factor) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_CUDA(cudaSetDevice(c->device));
    c->lsi_pending.active = false;
    lsi_enqueue(c, query_map_id, mode, xsect_factor);
  });
}

int rjb_lsi_wait(rjb_ctx* c, const rjb_xsect** d_xsects, uint64_t* n_xsects, uint64_t* n_candidates) {
  if (n_xsects) *n_xsects = 0;
  if (d_xsects) *d_xsects = nullptr;
  uint64_t needed = 0;
  int rc = guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_CUDA(cudaSetDevice(c->device));
    try {
      needed = lsi_finish(c, n_candidates);
    } catch (const Error& e) {
      if (e.code == RJB_ERR_QUEUE_OVERFLOW) needed = (uint32_t) c->last_stats[0];
      throw;
    }
    if (d_xsects) *d_xsects = c->xsects.p;
  });
  if (n_xsects) *n_xsects = needed;
  return rc;
}

int rjb_last_launches(const rjb_ctx* c, uint32_t* out) {
  return guarded([&] {
    RJB_REQUIRE(c && out, "NULL argument");
    *out = c->last_launches;
  });
}

int rjb_pip(rjb_ctx* c, int query_map_id, int mode, const int64_t* d_points_xy, uint64_t n_points,
            const uint32_t** d_closest_eid, const int32_t** d_face_id, uint64_t* n_candidates) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_CUDA(cudaSetDevice(c->device));
    check_map_id(query_map_id);
    const longlong2* pts = (const longlong2*) d_points_xy;
    const bool user_points = pts != nullptr;
    if (!pts) {
      const DeviceMap& Qm = c->maps[query_map_id];
      RJB_REQUIRE(Qm.loaded, "rjb_pip: query map not loaded and no points given");
      pts = Qm.pts.p;
      n_points = Qm.n_points;
    }
    RJB_REQUIRE(n_points < 0xFFFFFFF0ull, "rjb_pip: too many points");
    do_pip(c, query_map_id, mode, pts, (uint32_t) n_points, n_candidates, user_points);
    if (d_closest_eid) *d_closest_eid = c->pip_eid.p;
    if (d_face_id) *d_face_id = c->pip_face.p;
  });
}

int rjb_pip_host(rjb_ctx* c, int query_map_id, int mode, const double* h_xy, uint64_t n_points,
                 uint32_t* h_closest_eid, int32_t* h_face_id) {
  return guarded([&] {
    RJB_REQUIRE(c && (h_xy || n_points == 0), "NULL argument");
    RJB_REQUIRE(c->have_scaling, "rjb_pip_host: call rjb_set_bounding_box first");
    RJB_REQUIRE(n_points < 0xFFFFFFF0ull, "rjb_pip_host: too many points");
    RJB_CUDA(cudaSetDevice(c->device));
    uint32_t n = (uint32_t) n_points;
    double2* raw = c->pip_raw.ensure(n ? n : 1);
    longlong2* pts = c->pip_pts.ensure(n ? n : 1);
    if (n) {
      RJB_CUDA(cudaMemcpyAsync(raw, h_xy, (size_t) n * sizeof(double2), cudaMemcpyHostToDevice,
                               c->stream));
      k_scale_points<<<div_up(n, 256), 256, 0, c->stream>>>(raw, n, c->sc.rx, c->sc.ry,
                                                            c->sc.deltax, c->sc.deltay, pts);
    }
    do_pip(c, query_map_id, mode, pts, n, nullptr, true);
    if (n && h_closest_eid)
      RJB_CUDA(cudaMemcpyAsync(h_closest_eid, c->pip_eid.p, (size_t) n * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, c->stream));
    if (n && h_face_id)
      RJB_CUDA(cudaMemcpyAsync(h_face_id, c->pip_face.p, (size_t) n * sizeof(int32_t),
                               cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int rjb_pip_host_scaled(rjb_ctx* c, int query_map_id, int mode, const int64_t* h_points_xy,
                        uint64_t n_points, uint32_t* h_closest_eid, int32_t* h_face_id) {
  return guarded([&] {
    RJB_REQUIRE(c && (h_points_xy || n_points == 0), "NULL argument");
    RJB_REQUIRE(n_points < 0xFFFFFFF0ull, "rjb_pip_host_scaled: too many points");
    RJB_CUDA(cudaSetDevice(c->device));
    uint32_t n = (uint32_t) n_points;
    longlong2* pts = c->pip_pts.ensure(n ? n : 1);
    if (n)
      RJB_CUDA(cudaMemcpyAsync(pts, h_points_xy, (size_t) n * sizeof(longlong2),
                               cudaMemcpyHostToDevice, c->stream));
    do_pip(c, query_map_id, mode, pts, n, nullptr, true);
    if (n && h_closest_eid)
      RJB_CUDA(cudaMemcpyAsync(h_closest_eid, c->pip_eid.p, (size_t) n * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, c->stream));
    if (n && h_face_id)
      RJB_CUDA(cudaMemcpyAsync(h_face_id, c->pip_face.p, (size_t) n * sizeof(int32_t),
                               cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

// the events of the last query have completed (every query ends with a stream wait)
static void resolve_timing(rjb_ctx* m) {
  if (!m->timing_pending) return;
  for (int k = 0; k < kTimedStages; k++) m->last_ms[k] = 0;
  if (m->timing_layout == 2) {
    // LBVH LSI without the per-kernel events: {filter + traversal, 0, exact + points, 0}
    RJB_CUDA(cudaEventElapsedTime(&m->last_ms[0], m->ev[0], m->ev[2]));
    RJB_CUDA(cudaEventElapsedTime(&m->last_ms[2], m->ev[2], m->ev[4]));
    m->timing_layout = 1;
  } else {
    for (int k = 0; k < m->timing_pending; k++)
      RJB_CUDA(cudaEventElapsedTime(&m->last_ms[k], m->ev[k], m->ev[k + 1]));
  }
  m->timing_pending = 0;
}

/* device times (ms) of the kernels of the last rjb_lsi / rjb_pip call:
 * out[0] = traversal / cell kernel (LBVH LSI: occupancy filter + tree walk),
 * out[1] = exact pass (LBVH LSI: exact pass + point pass; grid / brute: point pass) */
int rjb_last_kernel_ms(const rjb_ctx* c, double out[2]) {
  return guarded([&] {
    RJB_REQUIRE(c && out, "NULL argument");
    rjb_ctx* m = const_cast<rjb_ctx*>(c);
    resolve_timing(m);
    if (c->timing_layout == 1) {
      out[0] = (double) c->last_ms[0] + c->last_ms[1];
      out[1] = (double) c->last_ms[2] + c->last_ms[3];
    } else if (c->timing_layout == 3) {  // grid LSI: {cell filter, exact pass, point pass}
      out[0] = (double) c->last_ms[0] + c->last_ms[1];
      out[1] = c->last_ms[2];
    } else if (c->timing_layout == 4) {  // PIP: {ordering, query kernel}
      out[0] = c->last_ms[1];
      out[1] = 0;
    } else {
      out[0] = c->last_ms[0];
      out[1] = c->last_ms[1];
    }
  });
}

/* per-kernel device times (ms) of the last query.  LBVH LSI: out = {k_lsi_filter,
 * k_lsi_bvh (+ k_lsi_cells), k_lsi_exact, k_lsi_points}; any other query: {query kernel,
 * point pass, 0, 0}.  *layout (optional) says which (1 = the four LBVH LSI stages). */
int rjb_last_stage_ms(const rjb_ctx* c, double out[4], int* layout) {
  return guarded([&] {
    RJB_REQUIRE(c && out, "NULL argument");
    rjb_ctx* m = const_cast<rjb_ctx*>(c);
    resolve_timing(m);
    for (int k = 0; k < kTimedStages; k++) out[k] = c->last_ms[k];
    if (layout) *layout = c->timing_layout;
  });
}

/* raw device counters of the last query: [0] results, [1] candidates, and with
 * option "stats" = 1: [2] node visits, [3] leaf visits, [4] single-child prefix
 * visits, [5] lane-level leaf tests, [6] warps that reached a leaf, [7] max stack */
int rjb_last_stats(const rjb_ctx* c, uint64_t out[8]) {
  return guarded([&] {
    RJB_REQUIRE(c && out, "NULL argument");
    for (int i = 0; i < 8; i++) out[i] = c->last_stats[i];
  });
}

/* bytes of the index of map_id (for the roofline's algorithmic bytes) */
int rjb_index_info(const rjb_ctx* c, int map_id, int mode, uint64_t out[4]) {
  return guarded([&] {
    RJB_REQUIRE(c && out, "NULL argument");
    check_map_id(map_id);
    const DeviceMap& m = c->maps[map_id];
    memset(out, 0, 4 * sizeof(uint64_t));
    if (mode == RJB_MODE_LBVH && m.bvh.built) {
      out[0] = m.bvh.n_leaves;
      out[1] = m.bvh.index_bytes();
      out[2] = m.bvh.leaf_size;
      out[3] = m.bvh.cell_directory_bytes();  // part of out[1]
    } else if (mode == RJB_MODE_GRID && m.grid.built) {
      out[0] = m.grid.n_items;
      out[1] = m.grid.index_bytes();
      out[2] = m.grid.gsize;
    }
  });
}

/* debug hook for the test-suite: sorts host (key, value) pairs on the device with
 * the engine's radix sort over key bits [begin_bit, end_bit) (stable) */
int rjb_debug_sort_pairs(rjb_ctx* c, uint64_t* h_keys, uint32_t* h_vals, uint64_t n, int begin_bit,
                         int end_bit) {
  return guarded([&] {
    RJB_REQUIRE(c && (n == 0 || (h_keys && h_vals)), "NULL argument");
    RJB_REQUIRE(n < (1u << 30), "too many pairs");
    RJB_CUDA(cudaSetDevice(c->device));
    if (n == 0) return;
    uint64_t* ka = c->ord_keys_a.ensure(n);
    uint64_t* kb = c->ord_keys_b.ensure(n);
    uint32_t* va = c->ord_vals_a.ensure(n);
    uint32_t* vb = c->ord_vals_b.ensure(n);
    RJB_CUDA(cudaMemcpyAsync(ka, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    RJB_CUDA(cudaMemcpyAsync(va, h_vals, n * 4, cudaMemcpyHostToDevice, c->stream));
    sort_pairs_u64_u32(ka, kb, va, vb, (uint32_t) n, begin_bit, end_bit, c->ord_sort, c->stream);
    RJB_CUDA(cudaMemcpyAsync(h_keys, kb, n * 8, cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaMemcpyAsync(h_vals, vb, n * 4, cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

/* debug hooks: the exact arithmetic on the device over caller-supplied cases (rjb_debug.cuh) */
int rjb_debug_intersect_batch(rjb_ctx* c, const int64_t* h_pts, uint64_t n, int mode, uint8_t* h_flags,
                              int64_t* h_x, int64_t* h_y) {
  return guarded([&] {
    RJB_REQUIRE(c && (n == 0 || (h_pts && h_flags && h_x && h_y)), "NULL argument");
    RJB_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
    RJB_CUDA(cudaSetDevice(c->device));
    if (n == 0) return;
    DBuf<long long> pts, x, y;
    DBuf<unsigned char> fl;
    pts.ensure(8 * n); x.ensure(n); y.ensure(n); fl.ensure(n);
    RJB_CUDA(cudaMemcpyAsync(pts.p, h_pts, 8 * n * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    k_debug_intersect<<<div_up(n, 128), 128, 0, c->stream>>>(pts.p, n, mode, fl.p, x.p, y.p);
    RJB_CUDA(cudaGetLastError());
    RJB_CUDA(cudaMemcpyAsync(h_flags, fl.p, n, cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaMemcpyAsync(h_x, x.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaMemcpyAsync(h_y, y.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int rjb_debug_i128_batch(rjb_ctx* c, const uint64_t* h_v, const uint64_t* h_d, uint64_t n, double* h_cvt,
                         double* h_div, int64_t* h_trunc) {
  return guarded([&] {
    RJB_REQUIRE(c && (n == 0 || (h_v && h_cvt)), "NULL argument");
    RJB_REQUIRE(!h_d || (h_div && h_trunc), "NULL argument");
    RJB_CUDA(cudaSetDevice(c->device));
    if (n == 0) return;
    DBuf<unsigned long long> v, d;
    DBuf<double> cvt, dv;
    DBuf<long long> tr;
    v.ensure(2 * n); cvt.ensure(n);
    RJB_CUDA(cudaMemcpyAsync(v.p, h_v, 16 * n, cudaMemcpyHostToDevice, c->stream));
    if (h_d) {
      d.ensure(2 * n); dv.ensure(n); tr.ensure(n);
      RJB_CUDA(cudaMemcpyAsync(d.p, h_d, 16 * n, cudaMemcpyHostToDevice, c->stream));
    }
    k_debug_i128<<<div_up(n, 256), 256, 0, c->stream>>>(v.p, h_d ? d.p : nullptr, n, cvt.p, dv.p, tr.p);
    RJB_CUDA(cudaGetLastError());
    RJB_CUDA(cudaMemcpyAsync(h_cvt, cvt.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_d) {
      RJB_CUDA(cudaMemcpyAsync(h_div, dv.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
      RJB_CUDA(cudaMemcpyAsync(h_trunc, tr.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    RJB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int rjb_debug_pip_batch(rjb_ctx* c, const int64_t* h_edges, uint64_t n_edges, const int64_t* h_pts,
                        uint64_t n, int query_map_id, uint32_t* h_out) {
  return guarded([&] {
    RJB_REQUIRE(c && (n == 0 || (h_pts && h_out)) && (n_edges == 0 || h_edges), "NULL argument");
    RJB_REQUIRE(n_edges < 0xFFFFFFF0ull, "too many edges");
    check_map_id(query_map_id);
    RJB_CUDA(cudaSetDevice(c->device));
    if (n == 0) return;
    DBuf<long long> e, p;
    DBuf<uint32_t> o;
    e.ensure(4 * n_edges + 1); p.ensure(2 * n); o.ensure(n);
    if (n_edges) RJB_CUDA(cudaMemcpyAsync(e.p, h_edges, 32 * n_edges, cudaMemcpyHostToDevice, c->stream));
    RJB_CUDA(cudaMemcpyAsync(p.p, h_pts, 16 * n, cudaMemcpyHostToDevice, c->stream));
    k_debug_pip<<<div_up(n, 128), 128, 0, c->stream>>>(e.p, (uint32_t) n_edges, p.p, n, query_map_id, o.p);
    RJB_CUDA(cudaGetLastError());
    RJB_CUDA(cudaMemcpyAsync(h_out, o.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

/* the packed-pair sort every path of the engine uses: words = (key << 32 | payload), sorted in
 * place (stable) by key bits [begin_bit, end_bit) of the 32-bit key */
int rjb_debug_sort_packed(rjb_ctx* c, uint64_t* h_words, uint64_t n, int begin_bit, int end_bit) {
  return guarded([&] {
    RJB_REQUIRE(c && (n == 0 || h_words), "NULL argument");
    RJB_REQUIRE(n < (1u << 30), "too many pairs");
    RJB_CUDA(cudaSetDevice(c->device));
    if (n == 0) return;
    uint64_t* ka = c->ord_keys_a.ensure(n);
    uint64_t* kb = c->ord_keys_b.ensure(n);
    RJB_CUDA(cudaMemcpyAsync(ka, h_words, n * 8, cudaMemcpyHostToDevice, c->stream));
    const uint64_t* sorted = sort_packed(ka, kb, (uint32_t) n, begin_bit, end_bit, c->ord_sort, c->stream);
    RJB_CUDA(cudaMemcpyAsync(h_words, sorted, n * 8, cudaMemcpyDeviceToHost, c->stream));
    RJB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int rjb_overlay_run(rjb_ctx* c, int mode, uint32_t grid_size, double xsect_factor,
                    double* phase_ms) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_CUDA(cudaSetDevice(c->device));
    overlay_run(c, mode, grid_size, xsect_factor, phase_ms);
  });
}

int rjb_overlay_finish(rjb_ctx* c, int mode, uint32_t grid_size, const rjb_xsect* h_xsects,
                       uint64_t n_xsects, const uint32_t* h_closest_eid0,
                       const int32_t* h_point_in_polygon0, const uint32_t* h_closest_eid1,
                       const int32_t* h_point_in_polygon1, double* phase_ms) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_REQUIRE(n_xsects == 0 || h_xsects, "rjb_overlay_finish: NULL xsects");
    RJB_REQUIRE(h_closest_eid0 && h_point_in_polygon0 && h_closest_eid1 && h_point_in_polygon1,
                "rjb_overlay_finish: NULL vertex-location arrays");
    RJB_CUDA(cudaSetDevice(c->device));
    OverlayImport imp;
    imp.h_xsects = h_xsects;
    imp.n_xsects = n_xsects;
    imp.h_closest_eid[0] = h_closest_eid0;
    imp.h_closest_eid[1] = h_closest_eid1;
    imp.h_point_in_polygon[0] = h_point_in_polygon0;
    imp.h_point_in_polygon[1] = h_point_in_polygon1;
    overlay_run(c, mode, grid_size, 0.0, phase_ms, &imp);
  });
}

int rjb_overlay_finish_device(rjb_ctx* c, int mode, uint32_t grid_size, const rjb_xsect* d_xsects,
                              uint64_t n_xsects, const uint32_t* d_closest_eid0,
                              const int32_t* d_point_in_polygon0, const uint32_t* d_closest_eid1,
                              const int32_t* d_point_in_polygon1, double* phase_ms) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_REQUIRE(n_xsects == 0 || d_xsects, "rjb_overlay_finish_device: NULL xsects");
    RJB_REQUIRE(d_closest_eid0 && d_point_in_polygon0 && d_closest_eid1 && d_point_in_polygon1,
                "rjb_overlay_finish_device: NULL vertex-location arrays");
    RJB_CUDA(cudaSetDevice(c->device));
    OverlayImport imp;
    imp.device = true;
    imp.h_xsects = d_xsects;
    imp.n_xsects = n_xsects;
    imp.h_closest_eid[0] = d_closest_eid0;
    imp.h_closest_eid[1] = d_closest_eid1;
    imp.h_point_in_polygon[0] = d_point_in_polygon0;
    imp.h_point_in_polygon[1] = d_point_in_polygon1;
    overlay_run(c, mode, grid_size, 0.0, phase_ms, &imp);
  });
}

int rjb_overlay_results(const rjb_ctx* c, int im, const rjb_xsect** d_xsects, uint64_t* n_xsects,
                        const uint32_t** d_closest_eid, const int32_t** d_point_in_polygon) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    check_map_id(im);
    RJB_REQUIRE(c->ov.done, "rjb_overlay_results: run rjb_overlay_run first");
    if (d_xsects) *d_xsects = c->ov.xsects_sorted[im].p;
    if (n_xsects) *n_xsects = c->ov.n_xsects;
    if (d_closest_eid) *d_closest_eid = c->ov.closest_eid[im].p;
    if (d_point_in_polygon) *d_point_in_polygon = c->ov.point_in_polygon[im].p;
  });
}

int rjb_overlay_write(rjb_ctx* c, const char* path) {
  return guarded([&] {
    RJB_REQUIRE(c && path, "NULL argument");
    RJB_CUDA(cudaSetDevice(c->device));
    overlay_write(c, path);
  });
}

int rjb_copy_to_host(rjb_ctx* c, const void* d_src, void* h_dst, uint64_t bytes) {
  return guarded([&] {
    RJB_REQUIRE(c && (bytes == 0 || (d_src && h_dst)), "NULL argument");
    if (bytes == 0) return;
    RJB_CUDA(cudaSetDevice(c->device));
    RJB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    wait_stream(c->stream);
  });
}

int rjb_sync(rjb_ctx* c) {
  return guarded([&] {
    RJB_REQUIRE(c, "ctx is NULL");
    RJB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

}  // extern "C"
