/*
 * rjb200.h -- C ABI of the B200-native spatial-join engine (LSI / PIP / overlay).
 *
 * This is the drop-in boundary for RayJoin's operator interfaces.  RayJoin has
 * no FFI of its own; its seam is three C++ abstract templates plus two CLIs
 * (paths relative to the RayJoin tree):
 *
 *   Context<coord_t, coefficient_t>    src/context.h:16-127
 *   LSI<CTX>::Init/Query/get_xsects    src/app/lsi.h:7-43
 *   PIP<CTX>::Init/Query/get_closest_eids   src/app/pip.h:8-38
 *   MapOverlay<CTX> 6-step protocol    src/app/map_overlay.h:9-56
 *   load_from / read_pgraph / .bin     src/map/planar_graph.h:41-252
 *   WriteOutputChain                   src/app/output_chain.h:41-205
 *
 * Every entry point below names the reference interface it replaces.  Plain
 * pointers and sizes only; no C++/torch types cross the boundary; no
 * exceptions cross it either: every call returns an int status (0 = ok) and
 * rjb_last_error() gives the message for the calling thread.
 *
 * Conventions
 *   - map ids are 0 / 1 exactly as in RayJoin (query_exec: base map R = 0,
 *     query map S = 1; polyover_exec: IntersectEdge(0) queries map 0 against
 *     the index of map 1).
 *   - edge ids (eid), point ids and chain numbering are RayJoin's:
 *     edge eid of chain c joins points (eid + c, eid + c + 1)
 *     (src/map/map.h:200-207).
 *   - "d_" pointers are device pointers owned by the context, valid until the
 *     next call of the same kind on that context (mirrors get_xsects()).
 *   - all calls are synchronous on the context's stream unless noted.
 */
#ifndef RJB200_H
#define RJB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rjb_ctx rjb_ctx;

/* status codes */
enum {
  RJB_OK = 0,
  RJB_ERR_INVALID = 1,        /* bad argument / call order                 */
  RJB_ERR_CUDA = 2,           /* CUDA runtime error (message has details)  */
  RJB_ERR_QUEUE_OVERFLOW = 3, /* xsect queue too small (see rjb_lsi)       */
  RJB_ERR_IO = 4,             /* file could not be read / written / parsed */
  RJB_ERR_NO_INDEX = 5        /* query before rjb_build_index              */
};

/* index / execution modes: RayJoin's -mode flag (src/flags.cc:8).  "rt" has
 * no equivalent on B200 (no RT cores); BRUTE is an all-pairs GPU mode that
 * exists for testing the exact arithmetic without any index.              */
enum { RJB_MODE_GRID = 0, RJB_MODE_LBVH = 1, RJB_MODE_BRUTE = 2 };

#define RJB_NO_HIT 0xFFFFFFFFu /* closest_eid of a point with no edge above */
#define RJB_EXTERIOR_FACE 0    /* src/config.h:8 */
#define RJB_DONTKNOW (-1)      /* src/config.h:3 */

/* One intersection.  Replaces dev::Intersection<int64_t> (src/algo/lsi.h:9-25,
 * 48 bytes: x.num,x.den,y.num,y.den,eid[2],mid_point_polygon_id); the
 * reference always stores denominators == 1 (truncating conversion,
 * src/util/rational.h:190-192), so they are dropped.                      */
typedef struct {
  int64_t x, y;                 /* scaled (internal) coordinates            */
  uint32_t eid[2];              /* eid[m] = edge of map m                   */
  int32_t mid_point_polygon_id; /* RJB_DONTKNOW until overlay fills it      */
  int32_t _pad;
} rjb_xsect;

/* Scaling<double> (src/map/scaling.h:32-136) as plain data */
typedef struct {
  double rx, ry, rrx, rry;
  double deltax, deltay, ddeltax, ddeltay;
  int64_t internal_min, internal_max, internal_range;
} rjb_scaling;

/* ---- lifetime ----------------------------------------------------------- */
const char* rjb_last_error(void);
const char* rjb_version(void);

/* Context ctor (src/context.h:31-74): owns one stream and both maps.       */
int rjb_create(int device, rjb_ctx** out);
void rjb_destroy(rjb_ctx* ctx);

/* Run all work of this context on an externally owned cudaStream_t
 * (e.g. torch's current stream) instead of the internal non-blocking one
 * (src/util/stream.h:13-27).  NULL restores the internal stream.           */
int rjb_set_stream(rjb_ctx* ctx, void* cuda_stream);

/* ---- maps ---------------------------------------------------------------
 * PlanarGraph<double> (src/map/planar_graph.h:32-39) handed over as host
 * SoA: xy = n_points x {x,y} doubles, row_index = n_chains+1 CSR offsets into
 * points, left/right = face ids per chain.  Host buffers are only read during
 * the call.  Replaces Context::LoadToDevice -> Map::LoadFrom
 * (src/context.h:76-88, src/map/map.h:161-233): uploads, scales with
 * fma + truncation exactly like the device kernel at map.h:171-180, and
 * derives RayJoin's edge numbering.
 *
 * The scaling must be fixed first (rjb_set_bounding_box) because RayJoin
 * derives it from the union bounding box of both maps (context.h:37-47).   */
int rjb_set_bounding_box(rjb_ctx* ctx, double min_x, double min_y, double max_x,
                         double max_y);
int rjb_get_scaling(const rjb_ctx* ctx, rjb_scaling* out);
int rjb_set_map(rjb_ctx* ctx, int map_id, const double* xy, uint64_t n_points,
                const uint32_t* row_index, const int64_t* left,
                const int64_t* right, uint64_t n_chains);
/* counts: out[0] = points, out[1] = edges, out[2] = chains */
int rjb_map_info(const rjb_ctx* ctx, int map_id, uint64_t out[3]);
/* device views of the loaded map (scaled points as int64 x,y pairs; per-edge
 * chain id so that p1 = eid + chain) for callers that stay on the GPU       */
int rjb_map_device_views(const rjb_ctx* ctx, int map_id,
                         const int64_t** d_points_xy,
                         const uint32_t** d_edge_chain);

/* ---- index build ---------------------------------------------------------
 * Replaces UniformGrid::AddMapToGrid (src/grid/uniform_grid.h:131-358) for
 * RJB_MODE_GRID and FillPrimitivesLBVH + lbvh::bvh::construct
 * (src/tree/primtive.h:33-57, deps/lbvh/lbvh/bvh.cuh:277-481) for
 * RJB_MODE_LBVH.  grid_size is RayJoin's -grid_size (ignored for LBVH);
 * build_ms (optional) receives the device time of the build.               */
int rjb_build_index(rjb_ctx* ctx, int map_id, int mode, uint32_t grid_size,
                    double* build_ms);

/* tuning knobs that have no RayJoin flag (defaults are fine):
 *   "lbvh_leaf_size"  edges per LBVH leaf, 1..8 (default 4)
 *   "lsi_window_begin", "lsi_window_end"  rjb_lsi queries only the edges that START at points
 *                     [begin, end) of the query map (0, 0 = the whole map, the default): a
 *                     multi-GPU driver keeps both maps whole on every rank and gives each rank
 *                     a window (whole chains) of the query side; edge ids stay global
 *   "lbvh_ag"         1 = adaptive leaf grouping: leaves are runs of consecutive chain edges
 *                     merged by RayJoin's Adaptive Grouping rule (-ag, src/rt/primitive.h:120-260:
 *                     neighbours merge while area(merged) / max(area) < enlarge), 0 (default) =
 *                     fixed runs of lbvh_leaf_size edges.  Set it before rjb_build_index.
 *   "lbvh_ag_iter"    merge rounds (-ag_iter, default 5; leaves hold <= 8 edges: 3 take effect)
 *   "lbvh_enlarge_x1000"  the area limit times 1000 (-enlarge, default 5.0 -> 5000)
 *   "sort_queries"    visit queries in Morton order: 1 on, 0 off, -1 auto (default:
 *                     LSI query edges are ordered when the query map averages
 *                     < 32 edges per chain; points only on request)
 *   "lsi_filter"      LBVH LSI occupancy pre-filter: -1 auto (default: on when the
 *                     base map occupies < 25 % of a 4096^2 bitmap), 0 off, 1 on
 *   "lsi_cells"       LBVH LSI: 1 = the filter's survivors look their candidate leaves up in a
 *                     directory of the occupied cells, built with the index (+0.15 ms and
 *                     +48 MB on a 4 M-edge map; edges longer than 3 x 3 cells still walk the
 *                     tree) and the exact pass reads no leaf records; 0 (default) = every
 *                     survivor walks the tree; 2 = like 1, and keep the directory even when a query
 *                     found edges too long for it (1 falls back to the walk for that query map:
 *                     their separate walk costs more than the directory saves).  Set it before
 *                     rjb_build_index.
 *   "lsi_fused"       LBVH LSI: 1 (default) = exact pass and point pass are one kernel whose last
 *                     CTA hands the counters to the host (no memset / memcpy around a query);
 *                     0 = two kernels (k_lsi_exact, k_lsi_points)
 *   "lsi_tile_filter" LBVH LSI: 1 (default) = two-level occupancy filter (tiles of 8 edges decided
 *                     first, 25 us instead of 35 us on the bench workload); 0 = one level
 *   "lsi_resolve_ctas" CTAs per SM of the fused kernel; 0 (default) = one resident wave
 *   "lsi_resolve_warp" LBVH LSI: 1 = the fused kernel with warp-private lists (no CTA barrier);
 *                     measured no faster; 0 (default)
 *   "lsi_pdl"         LBVH LSI: 1 = the kernels of a query are launched as programmatic dependents
 *                     (griddepcontrol); measured slower (early CTAs of the next kernel take the
 *                     registers the running one needs); 0 (default; 1 needs a -DRJB_PDL build)
 *   "pip_sort_bits"   grid PIP with ordered points: how many (high) bits of the cell key the
 *                     points are ordered by, default 24 = three radix passes
 *   "load_chunk_points" points per upload chunk of rjb_set_map (multiple of 1024, default 2^20):
 *                     the load kernel of chunk k runs while chunk k+1 is copied
 *   "pip_park"        LBVH PIP: 1 (default) = lanes park the leaf their ray meets and the
 *                     warp opens the parked leaves together; 0 = open a leaf when reached
 *   "stage_timing"    LBVH LSI: 1 (default) = a CUDA event after every kernel
 *                     (rjb_last_stage_ms), 0 = after every phase only, -1 = none (the times of
 *                     rjb_last_kernel_ms / rjb_last_stage_ms are 0 then; every event record
 *                     costs 2-3 us of stream time, ~8 % of a 0.13 ms query)
 *   "stats"           1 = collect traversal statistics (rjb_last_stats; slower)
 *   "keep_host_graph" 0 = rjb_set_map keeps no host copy of the source graph
 *                     (saves a memcpy; rjb_overlay_write then refuses)     */
int rjb_set_option(rjb_ctx* ctx, const char* name, int64_t value);

/* ---- LSI ------------------------------------------------------------------
 * LSI<CTX>::Init + Query + get_xsects (src/app/lsi.h:21-37,
 * src/app/lsi_lbvh.h:27-98, src/app/lsi_grid.h:97-159).  Queries every edge of
 * map `query_map_id` against the index of the other map with
 * e1 = query-side edge, e2 = base-side edge (lsi_lbvh.h:71).  The result
 * queue holds (|E0|+|E1|) * xsect_factor entries like
 * src/run_query.cu:226-228; unlike the reference (assert only,
 * src/util/queue.h:23-39) an overflow is detected: the call returns
 * RJB_ERR_QUEUE_OVERFLOW, *n_xsects is the number that was needed and no
 * partial result is exposed.  Result order is unspecified (atomic queue).
 * n_candidates (optional) = exact-predicate evaluations ("Total tests").    */
int rjb_lsi(rjb_ctx* ctx, int query_map_id, int mode, double xsect_factor,
            const rjb_xsect** d_xsects, uint64_t* n_xsects,
            uint64_t* n_candidates);

/* The same query in two halves, for callers that overlap host work (or several GPUs driven by
 * one thread) with the query: rjb_lsi_launch puts the kernels and the read-back of the counts on
 * the context's stream and returns without waiting; rjb_lsi_wait completes it with the result
 * contract of rjb_lsi (and repeats the query internally if an internal queue was too small).
 * Nothing else may be called on the context in between.  rjb_lsi == launch + wait.          */
int rjb_lsi_launch(rjb_ctx* ctx, int query_map_id, int mode, double xsect_factor);
int rjb_lsi_wait(rjb_ctx* ctx, const rjb_xsect** d_xsects, uint64_t* n_xsects,
                 uint64_t* n_candidates);

/* ---- PIP ------------------------------------------------------------------
 * PIP<CTX>::Query + get_closest_eids (src/app/pip.h:26-36,
 * src/app/pip_lbvh.h:25-142, src/app/pip_grid.h:21-77): for each query point
 * (scaled int64 x,y; or all vertices of map query_map_id when d_points_xy is
 * NULL, src/run_query.cu:346) the closest edge of the other map above it,
 * RJB_NO_HIT if none; face id = get_face_id (src/map/map.h:79-87) or
 * RJB_EXTERIOR_FACE.  d_points_xy is a DEVICE pointer.                      */
int rjb_pip(rjb_ctx* ctx, int query_map_id, int mode, const int64_t* d_points_xy,
            uint64_t n_points, const uint32_t** d_closest_eid,
            const int32_t** d_face_id, uint64_t* n_candidates);

/* host-buffer convenience used by the end-to-end path: unscaled double points
 * on the host are uploaded, scaled on the device (fma semantics) and queried;
 * results are copied back into caller buffers (may be NULL).               */
int rjb_pip_host(rjb_ctx* ctx, int query_map_id, int mode, const double* h_xy,
                 uint64_t n_points, uint32_t* h_closest_eid, int32_t* h_face_id);

/* same, for points the caller already scaled on the HOST (GeneratePIPQueries,
 * src/run_query.cu:146-167, scales -gen_n random points without FMA)        */
int rjb_pip_host_scaled(rjb_ctx* ctx, int query_map_id, int mode,
                        const int64_t* h_points_xy, uint64_t n_points,
                        uint32_t* h_closest_eid, int32_t* h_face_id);

/* ---- overlay ----------------------------------------------------------------
 * MapOverlay<CTX> protocol (src/app/map_overlay.h:20-30, call order of
 * src/run_overlay.cu:196-226).  rjb_overlay_run performs Init, BuildIndex (both
 * maps), IntersectEdge(0), LocateVerticesInOtherMap(0/1) and
 * ComputeOutputPolygons; phase_ms[6] (optional) receives
 * {build, lsi, pip0, pip1, polygons, total}.  rjb_overlay_write is
 * WriteOutputChain (src/app/output_chain.h:41-205), same text format.       */
int rjb_overlay_run(rjb_ctx* ctx, int mode, uint32_t grid_size,
                    double xsect_factor, double* phase_ms);
/* Multi-GPU overlay, final step on one rank: IntersectEdge and
 * LocateVerticesInOtherMap were run on shards (rjb_lsi / rjb_pip per rank) and
 * gathered; this call builds the indexes, imports the gathered host arrays
 * (edge / point ids of the UNSHARDED maps) and runs ComputeOutputPolygons, after
 * which rjb_overlay_results / rjb_overlay_write behave as after rjb_overlay_run. */
int rjb_overlay_finish(rjb_ctx* ctx, int mode, uint32_t grid_size,
                       const rjb_xsect* h_xsects, uint64_t n_xsects,
                       const uint32_t* h_closest_eid0, const int32_t* h_point_in_polygon0,
                       const uint32_t* h_closest_eid1, const int32_t* h_point_in_polygon1,
                       double* phase_ms);
/* The same with the gathered arrays in DEVICE memory (what a grouped ncclSend / ncclRecv
 * leaves on the finishing rank: no host staging).  When the context already holds the indexes of
 * this mode -- the finishing rank took part in the sharded phases -- they are reused.          */
int rjb_overlay_finish_device(rjb_ctx* ctx, int mode, uint32_t grid_size,
                              const rjb_xsect* d_xsects, uint64_t n_xsects,
                              const uint32_t* d_closest_eid0, const int32_t* d_point_in_polygon0,
                              const uint32_t* d_closest_eid1, const int32_t* d_point_in_polygon1,
                              double* phase_ms);
/* results of the last rjb_overlay_run, device pointers:
 *  xsects sorted by eid[im] and along the edge, with mid_point_polygon_id
 *  (xsect_edges_sorted_[im]); closest_eid / point_in_polygon per vertex of
 *  map im (get_closet_eids / get_point_in_polygon)                          */
int rjb_overlay_results(const rjb_ctx* ctx, int im, const rjb_xsect** d_xsects,
                        uint64_t* n_xsects, const uint32_t** d_closest_eid,
                        const int32_t** d_point_in_polygon);
int rjb_overlay_write(rjb_ctx* ctx, const char* path);

/* ---- introspection for measurement ------------------------------------------
 * device times (ms, CUDA events on the context's stream) of the kernels of the
 * last rjb_lsi / rjb_pip call: out[0] = traversal (or grid-cell) kernel,
 * out[1] = intersection-point pass.  Replaces the reference's Stopwatch /
 * -profile sub-stage timers (src/util/stopwatch.h).                            */
int rjb_last_kernel_ms(const rjb_ctx* ctx, double out[2]);
/* per-kernel device times (ms) of the last query, by *layout:
 *   1  LBVH LSI : {k_lsi_filter, k_lsi_bvh (+ k_lsi_cells), k_lsi_exact, k_lsi_points};
 *                 with lsi_fused: {filter, bvh / cells, k_lsi_resolve, ~0}
 *                 (option "stage_timing" = 0: one event per phase only,
 *                  {filter + traversal, 0, exact + points, 0})
 *   3  grid LSI : {k_grid_lsi_filter (+ big), k_grid_lsi_exact, k_lsi_points, 0}
 *   4  PIP      : {ordering of the points (0 when not ordered), query kernel (+ split), 0, 0}
 *   0  brute LSI: {all-pairs kernel, point pass, 0, 0}                                     */
int rjb_last_stage_ms(const rjb_ctx* ctx, double out[4], int* layout);
/* raw counters of the last query: [0] results, [1] candidates; with option
 * "stats" = 1 also traversal statistics ([2] binary node visits, [3] leaf visits,
 * [4] top-tree steps, [5] lane-level leaf tests, [6] warps that
 * reached a leaf, [7] deepest stack).  Without "stats", an LBVH LSI query reports
 * [2] = (query edge, leaf) pairs handed from the traversal to the exact pass,
 * [7] = survivors of the occupancy filter (0 = filter off), [5] = 1 when they went
 * through the cell directory, [6] = survivors that walked the tree instead (longer
 * than a cell).  Replaces the reference's Debug-build
 * "Total tests" / "Visited nodes" counters (src/app/lsi_lbvh.h:37-42,83-96).   */
int rjb_last_stats(const rjb_ctx* ctx, uint64_t out[8]);
/* number of kernels the last completed rjb_lsi / rjb_pip put on the stream (all attempts) */
int rjb_last_launches(const rjb_ctx* ctx, uint32_t* out);
/* index of map_id: out[0] = leaves (LBVH) / edge-cell incidences (grid),
 * out[1] = bytes of the index, out[2] = leaf size / grid size,
 * out[3] = bytes of the LBVH's cell directory (part of out[1]; 0 if none)     */
int rjb_index_info(const rjb_ctx* ctx, int map_id, int mode, uint64_t out[4]);

/* test hook: the engine's onesweep radix sort on host (key, value) pairs, key
 * bits [begin_bit, end_bit), stable; arrays are sorted in place                 */
int rjb_debug_sort_pairs(rjb_ctx* ctx, uint64_t* h_keys, uint32_t* h_vals,
                         uint64_t n, int begin_bit, int end_bit);

/* the packed-pair onesweep every path of the engine sorts with (rjb_sort.cuh): 64-bit words =
 * 32-bit key << 32 | 32-bit payload, sorted in place, stably, by bits [begin_bit, end_bit) of the
 * KEY (0 <= begin_bit < end_bit <= 32)                                                       */
int rjb_debug_sort_packed(rjb_ctx* ctx, uint64_t* h_words, uint64_t n, int begin_bit, int end_bit);

/* test hooks: the exact arithmetic of the query kernels (the same device functions) over
 * caller-supplied HOST arrays, so that the golden vectors of the reference's
 * src/algo/lsi.h:27-143 + src/util/rational.h:190-203 and known-answer tests of the
 * (double)(__int128) conversions go through the sm_100a compile.
 *  intersect: pts = n x 8 int64 {e1.p1, e1.p2, e2.p1, e2.p2} (e1 = query side); mode 0 = always
 *    through the gcd, 1 = the deferring path of the point kernel; flags bit 0 = intersects,
 *    bits 1 / 2 = x / y went through the deferred pass; x, y = the stored intersection point.
 *  i128: v, d = n x {lo, hi} words; cvt = (double) v, div = (double) v / (double) d,
 *    trunc = (int64) div  (d, div, trunc may be NULL together).
 *  pip: the update rule of src/algo/pip.h:27-96 scanned over the edges (n_edges x 4 int64) in
 *    order for every point (n x 2 int64); out = index of the chosen edge or RJB_NO_HIT.    */
int rjb_debug_intersect_batch(rjb_ctx* ctx, const int64_t* h_pts, uint64_t n, int mode,
                              uint8_t* h_flags, int64_t* h_x, int64_t* h_y);
int rjb_debug_i128_batch(rjb_ctx* ctx, const uint64_t* h_v, const uint64_t* h_d, uint64_t n,
                         double* h_cvt, double* h_div, int64_t* h_trunc);
int rjb_debug_pip_batch(rjb_ctx* ctx, const int64_t* h_edges, uint64_t n_edges,
                        const int64_t* h_pts, uint64_t n, int query_map_id, uint32_t* h_out);

/* ---- transfers --------------------------------------------------------------- */
int rjb_copy_to_host(rjb_ctx* ctx, const void* d_src, void* h_dst,
                     uint64_t bytes);
int rjb_sync(rjb_ctx* ctx);

/* ---- CDB files (host side) ---------------------------------------------------
 * read_pgraph / serialize_pgraph / deserialize_pgraph / load_from
 * (src/map/planar_graph.h:41-252).  The returned graph is owned by the
 * library; arrays stay valid until rjb_graph_free.                           */
typedef struct {
  uint64_t n_chains, n_points;
  const int64_t* chain_id;    /* n_chains */
  const int64_t* first_point; /* n_chains (as written in the file; unused) */
  const int64_t* last_point;  /* n_chains */
  const int64_t* left;        /* n_chains */
  const int64_t* right;       /* n_chains */
  const uint32_t* row_index;  /* n_chains + 1 (0 entries for an empty graph) */
  const double* xy;           /* n_points x 2 */
  double min_x, min_y, max_x, max_y;
  void* _owner;
} rjb_graph;

int rjb_graph_load(const char* path, const char* serialize_prefix,
                   rjb_graph* out);
int rjb_graph_read_text(const char* path, rjb_graph* out);
int rjb_graph_read_bin(const char* path, rjb_graph* out);
int rjb_graph_write_bin(const rjb_graph* g, const char* path);
void rjb_graph_free(rjb_graph* g);

#ifdef __cplusplus
}
#endif
#endif /* RJB200_H */
