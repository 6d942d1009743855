/*
 * oracle.c -- CPU restatement of RayJoin's LSI / PIP arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rayjoin_b200/ (the product) may
 * include, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity pin: the LSI predicate and the intersection point are checked
 * against the reference's own src/algo/lsi.h compiled on the host
 * (oracle/ref_lsi_pin.cc -> oracle/_ref/libref_lsi.so, tests/test_oracle_pin.py)
 * and against the committed vectors in tests/golden/ that were produced by
 * that library (tools/make_golden.py).  The PIP rule has no host-compilable
 * reference (it is embedded in device lambdas), so PIP parity is pinned on
 * the GPU box against the shim-built reference binary (oracle/_ref/ref_exec).
 *
 * Build: gcc -O2 -fopenmp -fwrapv -ffp-contract=off -shared -fPIC
 *   -fwrapv          : the reference's __int128 products can wrap for very
 *                      long edges (numx up to 2^142); nvcc wraps, so do we.
 *   -ffp-contract=off: host code of the reference is built without FMA
 *                      contraction; device scaling uses an explicit fma().
 *
 * Every function cites the reference file:line it restates
 * (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef __int128 i128;
typedef unsigned __int128 u128;

#define ORC_NO_HIT 0xFFFFFFFFu

/* ------------------------------------------------------------------ */
/* Scaling  (src/map/scaling.h:32-136)                                 */
/* ------------------------------------------------------------------ */
typedef struct {
  double rx, ry, rrx, rry;
  double deltax, deltay, ddeltax, ddeltay;
  int64_t imin, imax, irange;
} orc_scaling;

/* src/map/scaling.h:43-46 (shift 17 for double) and :56-71 */
void orc_scaling_init(orc_scaling* s, double bb_min_x, double bb_min_y,
                      double bb_max_x, double bb_max_y) {
  s->imax = INT64_MAX >> 17;
  s->imin = INT64_MIN >> 17;
  s->irange = s->imax - s->imin;
  /* SCALING_BOUNDING_BOX_MARGIN == 1 (src/config.h:4) */
  double max_x = bb_max_x + 1, min_x = bb_min_x - 1;
  double max_y = bb_max_y + 1, min_y = bb_min_y - 1;
  s->rx = (double) s->irange / (max_x - min_x);
  s->ry = (double) s->irange / (max_y - min_y);
  s->rrx = 1 / s->rx;
  s->rry = 1 / s->ry;
  /* (internal_max_ + internal_min_) is evaluated in int64 (= -1) first */
  int64_t isum = s->imax + s->imin;
  s->deltax = 0.5 * (isum - (max_x + min_x) * s->rx);
  s->deltay = 0.5 * (isum - (max_y + min_y) * s->ry);
  s->ddeltax = 0.5 * ((max_x + min_x) - isum * s->rrx);
  s->ddeltay = 0.5 * ((max_y + min_y) - isum * s->rry);
}

/* Device semantics: src/map/map.h:171-180 calls ScaleX/ScaleY
 * (scaling.h:79-95) inside a kernel; nvcc contracts x*rx+delta into
 * fma.rn.f64 and the int64 conversion is cvt.rzi (truncate).            */
void orc_scale_points_dev(const orc_scaling* s, const double* xy, uint64_t n,
                          int64_t* out) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; i++) {
    out[2 * i] = (int64_t) fma(xy[2 * i], s->rx, s->deltax);
    out[2 * i + 1] = (int64_t) fma(xy[2 * i + 1], s->ry, s->deltay);
  }
}

/* Host semantics (no FMA): GeneratePIPQueries, src/run_query.cu:146-167 */
void orc_scale_points_host(const orc_scaling* s, const double* xy, uint64_t n,
                           int64_t* out) {
  for (uint64_t i = 0; i < n; i++) {
    double tx = xy[2 * i] * s->rx;
    double ty = xy[2 * i + 1] * s->ry;
    out[2 * i] = (int64_t) (tx + s->deltax);
    out[2 * i + 1] = (int64_t) (ty + s->deltay);
  }
}

/* Host unscale, scaling.h:100-106 as used by output_chain.h:33-37 (no FMA) */
void orc_unscale_points_host(const orc_scaling* s, const int64_t* xy,
                             uint64_t n, double* out) {
  for (uint64_t i = 0; i < n; i++) {
    double tx = (double) xy[2 * i] * s->rrx;
    double ty = (double) xy[2 * i + 1] * s->rry;
    out[2 * i] = tx + s->ddeltax;
    out[2 * i + 1] = ty + s->ddeltay;
  }
}

/* ------------------------------------------------------------------ */
/* Edge numbering + equation  (src/map/map.h:187-230, :19-39)          */
/* ------------------------------------------------------------------ */
/* edge eid of chain c starts at point p = eid + c (n points -> n-1 edges) */
void orc_build_edges(const uint32_t* row_index, uint64_t n_chains,
                     uint32_t* edge_p1, uint32_t* edge_chain) {
  for (uint64_t c = 0; c < n_chains; c++) {
    for (uint32_t p = row_index[c]; p + 1 < row_index[c + 1]; p++) {
      uint32_t eid = p - (uint32_t) c;
      edge_p1[eid] = p;
      if (edge_chain)
        edge_chain[eid] = (uint32_t) c;
    }
  }
}

typedef struct {
  i128 a, b, c;
} edge_eq;

static inline edge_eq make_eq(int64_t x1, int64_t y1, int64_t x2, int64_t y2) {
  edge_eq e;
  e.a = (i128) y1 - y2;
  e.b = (i128) x2 - x1;
  e.c = -(i128) x1 * e.a - (i128) y1 * e.b;
  if (e.b < 0) {
    e.a = -e.a;
    e.b = -e.b;
    e.c = -e.c;
  }
  return e;
}

/* ------------------------------------------------------------------ */
/* LSI predicate  (src/algo/lsi.h:27-103)                              */
/* ------------------------------------------------------------------ */
static inline int sgn128(i128 v) { return (v > 0) - (v < 0); }

static inline int intersect_test_eq(const edge_eq* e1, int64_t e1p1x,
                                    int64_t e1p1y, int64_t e1p2x,
                                    int64_t e1p2y, const edge_eq* e2,
                                    int64_t e2p1x, int64_t e2p1y,
                                    int64_t e2p2x, int64_t e2p2y) {
#define SUBEDGE(px, py, e) ((i128) (px) * (e)->a + (i128) (py) * (e)->b + (e)->c)
  i128 e2_p1_agst_e1 = SUBEDGE(e2p1x, e2p1y, e1);
  i128 e2_p2_agst_e1 = SUBEDGE(e2p2x, e2p2y, e1);
  i128 e1_p1_agst_e2 = SUBEDGE(e1p1x, e1p1y, e2);
  i128 e1_p2_agst_e2 = SUBEDGE(e1p2x, e1p2y, e2);
#undef SUBEDGE
  /* lsi.h:42-60 : endpoints of e1 on the line of e2 -> perturb by -e2.a, -e2.b */
  if (e1_p1_agst_e2 == 0) e1_p1_agst_e2 = -e2->a;
  if (e1_p1_agst_e2 == 0) e1_p1_agst_e2 = -e2->b;
  if (e1_p1_agst_e2 == 0) return 0;
  if (e1_p2_agst_e2 == 0) e1_p2_agst_e2 = -e2->a;
  if (e1_p2_agst_e2 == 0) e1_p2_agst_e2 = -e2->b;
  if (e1_p2_agst_e2 == 0) return 0;
  /* lsi.h:64-67 */
  if ((e1_p1_agst_e2 > 0 && e1_p2_agst_e2 > 0) ||
      (e1_p1_agst_e2 < 0 && e1_p2_agst_e2 < 0))
    return 0;
  /* lsi.h:70-87 : endpoints of e2 on the line of e1 -> perturb by +e1.a, +e1.b */
  if (e2_p1_agst_e1 == 0) e2_p1_agst_e1 = e1->a;
  if (e2_p1_agst_e1 == 0) e2_p1_agst_e1 = e1->b;
  if (e2_p1_agst_e1 == 0) return 0;
  if (e2_p2_agst_e1 == 0) e2_p2_agst_e1 = e1->a;
  if (e2_p2_agst_e1 == 0) e2_p2_agst_e1 = e1->b;
  if (e2_p2_agst_e1 == 0) return 0;
  /* lsi.h:88-91 */
  if ((e2_p1_agst_e1 > 0 && e2_p2_agst_e1 > 0) ||
      (e2_p1_agst_e1 < 0 && e2_p2_agst_e1 < 0))
    return 0;
  /* lsi.h:97-100 : identical edges never intersect */
  if ((e1p1x == e2p1x && e1p1y == e2p1y && e1p2x == e2p2x && e1p2y == e2p2y) ||
      (e1p1x == e2p2x && e1p1y == e2p2y && e1p2x == e2p1x && e1p2y == e2p1y))
    return 0;
  return 1;
}

/* tcb::rational<__int128>(num, den) constructor -> simplify()
 * (src/util/rational.h:36-43, :88-91, :198-203)                         */
static inline void rat_make(i128 num, i128 den, i128* onum, i128* oden) {
  i128 a = num, b = den;
  while (b != 0) {
    i128 t = b;
    b = a % b;
    a = t;
  }
  i128 g = a < 0 ? -a : a;
  i128 sign_den = den < 0 ? -1 : 1;
  *onum = sign_den * num / g;
  *oden = (den < 0 ? -den : den) / g;
}

static inline int64_t min4(int64_t a, int64_t b, int64_t c, int64_t d) {
  int64_t m = a < b ? a : b, n = c < d ? c : d;
  return m < n ? m : n;
}
static inline int64_t max4(int64_t a, int64_t b, int64_t c, int64_t d) {
  int64_t m = a > b ? a : b, n = c > d ? c : d;
  return m > n ? m : n;
}

/* lsi.h:105-143 + the rational<int64> = rational<int128> conversion that
 * goes through operator double() and truncates (rational.h:84-85,190-192;
 * SURVEY section 0).  Result: integer x,y with denominator 1.              */
static inline void xsect_point_eq(const edge_eq* e1, int64_t e1p1x,
                                  int64_t e1p1y, int64_t e1p2x, int64_t e1p2y,
                                  const edge_eq* e2, int64_t e2p1x,
                                  int64_t e2p1y, int64_t e2p2x, int64_t e2p2y,
                                  int64_t* ox, int64_t* oy) {
  i128 denom = e1->a * e2->b - e2->a * e1->b;
  i128 numx = e2->c * e1->b - e1->c * e2->b;
  i128 numy = e2->a * e1->c - e1->a * e2->c;
  i128 xn, xd, yn, yd;
  rat_make(numx, denom, &xn, &xd);
  rat_make(numy, denom, &yn, &yd);
  /* operator<(rational, integer): num*1 < t*den  (rational.h:329-333) */
  int64_t t = min4(e1p1x, e1p2x, e2p1x, e2p2x);
  if (xn < (i128) t * xd) { xn = t; xd = 1; }
  t = max4(e1p1x, e1p2x, e2p1x, e2p2x);
  if ((i128) t * xd < xn) { xn = t; xd = 1; }
  t = min4(e1p1y, e1p2y, e2p1y, e2p2y);
  if (yn < (i128) t * yd) { yn = t; yd = 1; }
  t = max4(e1p1y, e1p2y, e2p1y, e2p2y);
  if ((i128) t * yd < yn) { yn = t; yd = 1; }
  *ox = (int64_t) ((double) xn / (double) xd);
  *oy = (int64_t) ((double) yn / (double) yd);
}

/* single-pair entry points (for KATs / golden vectors).
 * pts = {e1p1x,e1p1y,e1p2x,e1p2y,e2p1x,e2p1y,e2p2x,e2p2y}                */
int orc_intersect_test(const int64_t* p) {
  edge_eq e1 = make_eq(p[0], p[1], p[2], p[3]);
  edge_eq e2 = make_eq(p[4], p[5], p[6], p[7]);
  return intersect_test_eq(&e1, p[0], p[1], p[2], p[3], &e2, p[4], p[5], p[6],
                           p[7]);
}

int orc_intersect_point(const int64_t* p, int64_t* ox, int64_t* oy) {
  edge_eq e1 = make_eq(p[0], p[1], p[2], p[3]);
  edge_eq e2 = make_eq(p[4], p[5], p[6], p[7]);
  if (!intersect_test_eq(&e1, p[0], p[1], p[2], p[3], &e2, p[4], p[5], p[6],
                         p[7]))
    return 0;
  xsect_point_eq(&e1, p[0], p[1], p[2], p[3], &e2, p[4], p[5], p[6], p[7], ox,
                 oy);
  return 1;
}

/* batch version: n pairs, 8 int64 each; hit[i] in {0,1}; x,y valid if hit */
void orc_intersect_batch(const int64_t* pts, uint64_t n, uint8_t* hit,
                         int64_t* x, int64_t* y) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; i++) {
    int64_t ox = 0, oy = 0;
    hit[i] = (uint8_t) orc_intersect_point(pts + 8 * i, &ox, &oy);
    x[i] = ox;
    y[i] = oy;
  }
}

/* ------------------------------------------------------------------ */
/* LSI over two maps                                                   */
/*   e1 = query-side edge, e2 = base-side edge, as in                  */
/*   src/app/lsi_lbvh.h:44-81 (callback at :71) and                    */
/*   src/algo/rt_lsi_custom.cu:36-38.                                  */
/* ------------------------------------------------------------------ */
typedef struct {
  uint32_t eq, eb;
  int64_t x, y;
} pair_rec;

static int pair_cmp(const void* a, const void* b) {
  const pair_rec* p = (const pair_rec*) a;
  const pair_rec* q = (const pair_rec*) b;
  if (p->eq != q->eq) return p->eq < q->eq ? -1 : 1;
  if (p->eb != q->eb) return p->eb < q->eb ? -1 : 1;
  return 0;
}

typedef struct {
  pair_rec* v;
  uint64_t n, cap;
} pair_vec;

static void pv_push(pair_vec* pv, pair_rec r) {
  if (pv->n == pv->cap) {
    pv->cap = pv->cap ? pv->cap * 2 : 1024;
    pv->v = (pair_rec*) realloc(pv->v, pv->cap * sizeof(pair_rec));
  }
  pv->v[pv->n++] = r;
}

static inline int test_pair(const int64_t* xyq, uint32_t pq, const int64_t* xyb,
                            uint32_t pb, int64_t* ox, int64_t* oy) {
  const int64_t* q = xyq + 2 * (uint64_t) pq;
  const int64_t* b = xyb + 2 * (uint64_t) pb;
  edge_eq e1 = make_eq(q[0], q[1], q[2], q[3]);
  edge_eq e2 = make_eq(b[0], b[1], b[2], b[3]);
  if (!intersect_test_eq(&e1, q[0], q[1], q[2], q[3], &e2, b[0], b[1], b[2],
                         b[3]))
    return 0;
  xsect_point_eq(&e1, q[0], q[1], q[2], q[3], &e2, b[0], b[1], b[2], b[3], ox,
                 oy);
  return 1;
}

static uint64_t emit_sorted(pair_vec* pvs, int nt, uint32_t* out_eq,
                            uint32_t* out_eb, int64_t* out_x, int64_t* out_y,
                            uint64_t cap) {
  uint64_t total = 0;
  for (int t = 0; t < nt; t++) total += pvs[t].n;
  pair_rec* all = (pair_rec*) malloc((total ? total : 1) * sizeof(pair_rec));
  uint64_t o = 0;
  for (int t = 0; t < nt; t++) {
    memcpy(all + o, pvs[t].v, pvs[t].n * sizeof(pair_rec));
    o += pvs[t].n;
    free(pvs[t].v);
  }
  qsort(all, total, sizeof(pair_rec), pair_cmp);
  uint64_t w = total < cap ? total : cap;
  for (uint64_t i = 0; i < w; i++) {
    out_eq[i] = all[i].eq;
    out_eb[i] = all[i].eb;
    if (out_x) out_x[i] = all[i].x;
    if (out_y) out_y[i] = all[i].y;
  }
  free(all);
  return total;
}

/* All |Q|x|B| pairs, no filter: the definition of the pair set. Returns the
 * total number of intersecting pairs (may exceed cap; only cap are written),
 * sorted by (eid_query, eid_base).                                       */
uint64_t orc_lsi_brute(const int64_t* xyq, const uint32_t* q_p1, uint64_t nq,
                       const int64_t* xyb, const uint32_t* b_p1, uint64_t nb,
                       uint32_t* out_eq, uint32_t* out_eb, int64_t* out_x,
                       int64_t* out_y, uint64_t cap) {
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  pair_vec* pvs = (pair_vec*) calloc(nt, sizeof(pair_vec));
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
#pragma omp for schedule(dynamic, 64)
    for (uint64_t i = 0; i < nq; i++) {
      for (uint64_t j = 0; j < nb; j++) {
        int64_t x, y;
        if (test_pair(xyq, q_p1[i], xyb, b_p1[j], &x, &y)) {
          pair_rec r = {(uint32_t) i, (uint32_t) j, x, y};
          pv_push(&pvs[tid], r);
        }
      }
    }
  }
  uint64_t total = emit_sorted(pvs, nt, out_eq, out_eb, out_x, out_y, cap);
  free(pvs);
  return total;
}

/* -- host uniform grid used only as a candidate filter for the oracle -- */
typedef struct {
  int gsize, shift;
  int64_t imin;
  uint64_t* cell_begin; /* gsize*gsize + 1 */
  uint32_t* items;
} host_grid;

static inline int cell_of(const host_grid* g, int64_t v) {
  int64_t c = (v - g->imin) >> g->shift;
  if (c < 0) c = 0;
  if (c >= g->gsize) c = g->gsize - 1;
  return (int) c;
}

static void grid_build(host_grid* g, const int64_t* xy, const uint32_t* p1,
                       uint64_t ne, int64_t imin, int64_t irange) {
  /* about one edge per cell, gsize a power of two in [16, 8192] */
  int gs = 16;
  while ((uint64_t) gs * gs < ne && gs < 8192) gs <<= 1;
  int shift = 0;
  while ((((u128) 1 << shift) * (u128) gs) <= (u128) irange) shift++;
  g->gsize = gs;
  g->shift = shift;
  g->imin = imin;
  uint64_t ncell = (uint64_t) gs * gs;
  g->cell_begin = (uint64_t*) calloc(ncell + 2, sizeof(uint64_t));
  for (int pass = 0; pass < 2; pass++) {
    for (uint64_t e = 0; e < ne; e++) {
      const int64_t* p = xy + 2 * (uint64_t) p1[e];
      int x0 = cell_of(g, p[0] < p[2] ? p[0] : p[2]);
      int x1 = cell_of(g, p[0] < p[2] ? p[2] : p[0]);
      int y0 = cell_of(g, p[1] < p[3] ? p[1] : p[3]);
      int y1 = cell_of(g, p[1] < p[3] ? p[3] : p[1]);
      for (int cy = y0; cy <= y1; cy++)
        for (int cx = x0; cx <= x1; cx++) {
          uint64_t c = (uint64_t) cy * gs + cx;
          if (pass == 0)
            g->cell_begin[c + 2]++;
          else
            g->items[g->cell_begin[c + 1]++] = (uint32_t) e;
        }
    }
    if (pass == 0) {
      /* cell_begin[c+2] holds counts; prefix so cell_begin[c+1] = start(c) */
      for (uint64_t c = 0; c < ncell; c++)
        g->cell_begin[c + 2] += g->cell_begin[c + 1];
      g->items = (uint32_t*) malloc(
          (g->cell_begin[ncell + 1] ? g->cell_begin[ncell + 1] : 1) *
          sizeof(uint32_t));
    }
  }
  /* after fill, cell_begin[c+1] = end(c) = start(c+1); start(0) = 0 */
}

static void grid_free(host_grid* g) {
  free(g->cell_begin);
  free(g->items);
}

/* Grid-filtered LSI: same pair set as orc_lsi_brute (tests assert that),
 * usable at millions of edges.  n_candidates = exact-predicate evaluations. */
uint64_t orc_lsi_grid(const int64_t* xyq, const uint32_t* q_p1, uint64_t nq,
                      const int64_t* xyb, const uint32_t* b_p1, uint64_t nb,
                      int64_t imin, int64_t irange, uint32_t* out_eq,
                      uint32_t* out_eb, int64_t* out_x, int64_t* out_y,
                      uint64_t cap, uint64_t* n_candidates) {
  host_grid g;
  grid_build(&g, xyb, b_p1, nb, imin, irange);
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  pair_vec* pvs = (pair_vec*) calloc(nt, sizeof(pair_vec));
  uint64_t ncand = 0;
#pragma omp parallel reduction(+ : ncand)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
#pragma omp for schedule(dynamic, 256)
    for (uint64_t i = 0; i < nq; i++) {
      const int64_t* q = xyq + 2 * (uint64_t) q_p1[i];
      int64_t qx0 = q[0] < q[2] ? q[0] : q[2], qx1 = q[0] < q[2] ? q[2] : q[0];
      int64_t qy0 = q[1] < q[3] ? q[1] : q[3], qy1 = q[1] < q[3] ? q[3] : q[1];
      int cx0 = cell_of(&g, qx0), cx1 = cell_of(&g, qx1);
      int cy0 = cell_of(&g, qy0), cy1 = cell_of(&g, qy1);
      for (int cy = cy0; cy <= cy1; cy++)
        for (int cx = cx0; cx <= cx1; cx++) {
          uint64_t c = (uint64_t) cy * g.gsize + cx;
          for (uint64_t k = g.cell_begin[c]; k < g.cell_begin[c + 1]; k++) {
            uint32_t j = g.items[k];
            const int64_t* b = xyb + 2 * (uint64_t) b_p1[j];
            int64_t bx0 = b[0] < b[2] ? b[0] : b[2],
                    bx1 = b[0] < b[2] ? b[2] : b[0];
            int64_t by0 = b[1] < b[3] ? b[1] : b[3],
                    by1 = b[1] < b[3] ? b[3] : b[1];
            /* predicate true => closed boxes overlap (exact integers) */
            if (bx1 < qx0 || qx1 < bx0 || by1 < qy0 || qy1 < by0) continue;
            /* visit each pair once: in the cell of the box-intersection's
             * lower-left corner, which both edges are registered in        */
            int64_t lx = qx0 > bx0 ? qx0 : bx0, ly = qy0 > by0 ? qy0 : by0;
            if (cell_of(&g, lx) != cx || cell_of(&g, ly) != cy) continue;
            ncand++;
            int64_t x, y;
            if (test_pair(xyq, q_p1[i], xyb, b_p1[j], &x, &y)) {
              pair_rec r = {(uint32_t) i, j, x, y};
              pv_push(&pvs[tid], r);
            }
          }
        }
    }
  }
  uint64_t total = emit_sorted(pvs, nt, out_eq, out_eb, out_x, out_y, cap);
  free(pvs);
  grid_free(&g);
  if (n_candidates) *n_candidates = ncand;
  return total;
}

/* ------------------------------------------------------------------ */
/* LSI with the semantics of the reference's GRID backend               */
/*   src/app/lsi_grid.h:19-78: for every cell, every (map-0 edge,       */
/*   map-1 edge) pair registered in it: intersect_test(e1 = map 0,      */
/*   e2 = map 1) -- the query map id is ignored (:96-104) -- and the    */
/*   pair is kept only if calculate_cell(xsect) is this very cell       */
/*   (:62-67).  Cells of an edge: the cell-box of its end points        */
/*   (src/grid/uniform_grid.h:44-86).  So a pair is reported iff the    */
/*   predicate holds and the cell of the intersection point lies in the */
/*   cell-boxes of both edges.                                          */
/* ------------------------------------------------------------------ */
/* calculate_cell for an integer coordinate (src/grid/cell.h:15-22) */
static inline int ref_cell_int(int64_t v, int64_t imin, double cell_scale) {
  return (int) ((double) (v - imin) * cell_scale);
}

/* calculate_cell for the rational intersection coordinate: `val - internal_min` is
 * rational{num*1 - imin*den, den*1} simplified (rational.h:399-407), `* cell_scale` has no
 * rational overload, so the value converts through operator double (rational.h:94-97) */
static inline int ref_cell_rat(i128 num, i128 den, int64_t imin, double cell_scale) {
  i128 n1 = num - (i128) imin * den, n2, d2;
  rat_make(n1, den, &n2, &d2);
  return (int) (((double) n2 / (double) d2) * cell_scale);
}

/* lsi.h:105-143 without the final truncation: the clamped rationals */
static inline void xsect_rational_eq(const edge_eq* e1, int64_t e1p1x, int64_t e1p1y, int64_t e1p2x,
                                     int64_t e1p2y, const edge_eq* e2, int64_t e2p1x, int64_t e2p1y,
                                     int64_t e2p2x, int64_t e2p2y, i128* xn, i128* xd, i128* yn,
                                     i128* yd) {
  i128 denom = e1->a * e2->b - e2->a * e1->b;
  i128 numx = e2->c * e1->b - e1->c * e2->b;
  i128 numy = e2->a * e1->c - e1->a * e2->c;
  rat_make(numx, denom, xn, xd);
  rat_make(numy, denom, yn, yd);
  int64_t t = min4(e1p1x, e1p2x, e2p1x, e2p2x);
  if (*xn < (i128) t * *xd) { *xn = t; *xd = 1; }
  t = max4(e1p1x, e1p2x, e2p1x, e2p2x);
  if ((i128) t * *xd < *xn) { *xn = t; *xd = 1; }
  t = min4(e1p1y, e1p2y, e2p1y, e2p2y);
  if (*yn < (i128) t * *yd) { *yn = t; *yd = 1; }
  t = max4(e1p1y, e1p2y, e2p1y, e2p2y);
  if ((i128) t * *yd < *yn) { *yn = t; *yd = 1; }
}

/* a = map-0 edge (4 coords), b = map-1 edge; returns 1 and the stored point when the
 * reference's grid backend reports the pair */
static inline int refgrid_pair(const int64_t* a, const int64_t* b, int gsize, int64_t imin,
                               double cell_scale, int64_t* ox, int64_t* oy) {
  edge_eq e1 = make_eq(a[0], a[1], a[2], a[3]);
  edge_eq e2 = make_eq(b[0], b[1], b[2], b[3]);
  if (!intersect_test_eq(&e1, a[0], a[1], a[2], a[3], &e2, b[0], b[1], b[2], b[3])) return 0;
  i128 xn, xd, yn, yd;
  xsect_rational_eq(&e1, a[0], a[1], a[2], a[3], &e2, b[0], b[1], b[2], b[3], &xn, &xd, &yn, &yd);
  int cx = ref_cell_rat(xn, xd, imin, cell_scale), cy = ref_cell_rat(yn, yd, imin, cell_scale);
  (void) gsize;
#define CMIN(u, v) ref_cell_int((u) < (v) ? (u) : (v), imin, cell_scale)
#define CMAX(u, v) ref_cell_int((u) < (v) ? (v) : (u), imin, cell_scale)
  int ok = cx >= CMIN(a[0], a[2]) && cx <= CMAX(a[0], a[2]) && cx >= CMIN(b[0], b[2]) && cx <= CMAX(b[0], b[2]) &&
           cy >= CMIN(a[1], a[3]) && cy <= CMAX(a[1], a[3]) && cy >= CMIN(b[1], b[3]) && cy <= CMAX(b[1], b[3]);
#undef CMIN
#undef CMAX
  if (!ok) return 0;
  *ox = (int64_t) ((double) xn / (double) xd);
  *oy = (int64_t) ((double) yn / (double) yd);
  return 1;
}

/* batch form for known-answer tests: pts = n x 8 {map-0 edge, map-1 edge}; cx, cy = cell of the
 * intersection point as the reference computes it (valid where hit), owned = pair reported */
void orc_refgrid_cells(const int64_t* pts, uint64_t n, int gsize, int64_t imin, int64_t irange,
                       uint8_t* hit, int32_t* cx, int32_t* cy, uint8_t* owned) {
  double cell_scale = (double) gsize / irange * 0.999;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; i++) {
    const int64_t* a = pts + 8 * i;
    const int64_t* b = a + 4;
    edge_eq e1 = make_eq(a[0], a[1], a[2], a[3]);
    edge_eq e2 = make_eq(b[0], b[1], b[2], b[3]);
    hit[i] = (uint8_t) intersect_test_eq(&e1, a[0], a[1], a[2], a[3], &e2, b[0], b[1], b[2], b[3]);
    cx[i] = cy[i] = 0;
    owned[i] = 0;
    if (!hit[i]) continue;
    i128 xn, xd, yn, yd;
    xsect_rational_eq(&e1, a[0], a[1], a[2], a[3], &e2, b[0], b[1], b[2], b[3], &xn, &xd, &yn, &yd);
    cx[i] = ref_cell_rat(xn, xd, imin, cell_scale);
    cy[i] = ref_cell_rat(yn, yd, imin, cell_scale);
    int64_t x, y;
    owned[i] = (uint8_t) refgrid_pair(a, b, gsize, imin, cell_scale, &x, &y);
  }
}

/* xy0/p1_0 = map 0, xy1/p1_1 = map 1.  Output sorted by (first, second) where first is the
 * edge of map `sort_map`: out_first = eids of that map, out_second = eids of the other.
 * The candidate filter (a host grid over map 0, all-pairs when brute != 0) loses nothing:
 * the predicate implies overlapping boxes. */
uint64_t orc_lsi_refgrid(const int64_t* xy0, const uint32_t* p1_0, uint64_t n0, const int64_t* xy1,
                         const uint32_t* p1_1, uint64_t n1, int64_t imin, int64_t irange, int gsize,
                         int sort_map, int brute, uint32_t* out_first, uint32_t* out_second,
                         int64_t* out_x, int64_t* out_y, uint64_t cap) {
  double cell_scale = (double) gsize / irange * 0.999; /* cell.h:19 */
  host_grid g;
  if (!brute) grid_build(&g, xy0, p1_0, n0, imin, irange);
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  pair_vec* pvs = (pair_vec*) calloc(nt, sizeof(pair_vec));
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
#pragma omp for schedule(dynamic, 256)
    for (uint64_t i = 0; i < n1; i++) {
      const int64_t* q = xy1 + 2 * (uint64_t) p1_1[i];
      if (brute) {
        for (uint64_t j = 0; j < n0; j++) {
          int64_t x, y;
          if (refgrid_pair(xy0 + 2 * (uint64_t) p1_0[j], q, gsize, imin, cell_scale, &x, &y)) {
            pair_rec r = {sort_map == 1 ? (uint32_t) i : (uint32_t) j, sort_map == 1 ? (uint32_t) j : (uint32_t) i, x, y};
            pv_push(&pvs[tid], r);
          }
        }
        continue;
      }
      int64_t qx0 = q[0] < q[2] ? q[0] : q[2], qx1 = q[0] < q[2] ? q[2] : q[0];
      int64_t qy0 = q[1] < q[3] ? q[1] : q[3], qy1 = q[1] < q[3] ? q[3] : q[1];
      int cx0 = cell_of(&g, qx0), cx1 = cell_of(&g, qx1);
      int cy0 = cell_of(&g, qy0), cy1 = cell_of(&g, qy1);
      for (int cy = cy0; cy <= cy1; cy++)
        for (int cx = cx0; cx <= cx1; cx++) {
          uint64_t c = (uint64_t) cy * g.gsize + cx;
          for (uint64_t k = g.cell_begin[c]; k < g.cell_begin[c + 1]; k++) {
            uint32_t j = g.items[k];
            const int64_t* b = xy0 + 2 * (uint64_t) p1_0[j];
            int64_t bx0 = b[0] < b[2] ? b[0] : b[2], bx1 = b[0] < b[2] ? b[2] : b[0];
            int64_t by0 = b[1] < b[3] ? b[1] : b[3], by1 = b[1] < b[3] ? b[3] : b[1];
            if (bx1 < qx0 || qx1 < bx0 || by1 < qy0 || qy1 < by0) continue;
            int64_t lx = qx0 > bx0 ? qx0 : bx0, ly = qy0 > by0 ? qy0 : by0;
            if (cell_of(&g, lx) != cx || cell_of(&g, ly) != cy) continue;
            int64_t x, y;
            if (refgrid_pair(b, q, gsize, imin, cell_scale, &x, &y)) {
              pair_rec r = {sort_map == 1 ? (uint32_t) i : j, sort_map == 1 ? j : (uint32_t) i, x, y};
              pv_push(&pvs[tid], r);
            }
          }
        }
    }
  }
  uint64_t total = emit_sorted(pvs, nt, out_first, out_second, out_x, out_y, cap);
  free(pvs);
  if (!brute) grid_free(&g);
  return total;
}

/* ------------------------------------------------------------------ */
/* PIP: closest edge above the point                                   */
/*   rule text identical in src/algo/pip.h:27-96,                      */
/*   src/app/pip_lbvh.h:57-123, src/algo/rt_pip_custom.cu:44-106       */
/* ------------------------------------------------------------------ */
typedef struct {
  double best_y;
  i128 best_a, best_b;
  uint32_t best_eid;
} pip_state;

/* Applies the update rule for one candidate edge.  Full ties (same y* and
 * same slope: coincident edges) are resolved the way a scan in increasing
 * eid order resolves them: q==1 keeps the first (smallest eid), q==0 keeps
 * the last (largest eid); the reference leaves this traversal-order
 * dependent (src/run_query.cu:52-53).                                   */
static inline void pip_update(pip_state* st, int query_map_id, int64_t px,
                              int64_t py, const int64_t* e /*x1,y1,x2,y2*/,
                              uint32_t eid) {
  int64_t x_min = e[0] < e[2] ? e[0] : e[2];
  int64_t x_max = e[0] < e[2] ? e[2] : e[0];
  if (px < x_min || px > x_max || px == (query_map_id == 0 ? x_min : x_max))
    return;
  edge_eq eq = make_eq(e[0], e[1], e[2], e[3]);
  double xsect_y = (double) (-eq.a * px - eq.c) / (double) eq.b;
  double diff_y = (double) py - xsect_y;
  if (diff_y == 0) diff_y = (double) (query_map_id == 0 ? -eq.a : eq.a);
  if (diff_y == 0) diff_y = (double) (query_map_id == 0 ? -eq.b : eq.b);
  if (diff_y > 0) return;
  if (xsect_y > st->best_y) return;
  if (xsect_y == st->best_y) {
    double cur = (double) eq.a / (double) eq.b;
    double best = (double) st->best_a / (double) st->best_b;
    if (cur == best) {
      /* coincident edges: deterministic eid rule (see above) */
      if (query_map_id ? (eid > st->best_eid) : (eid < st->best_eid)) return;
    } else {
      int flag = cur > best;
      if ((query_map_id && !flag) || (flag && !query_map_id)) return;
    }
  }
  st->best_y = xsect_y;
  st->best_a = eq.a;
  st->best_b = eq.b;
  st->best_eid = eid;
}

void orc_pip_brute(const int64_t* xyb, const uint32_t* b_p1, uint64_t nb,
                   const int64_t* pts, uint64_t n, int query_map_id,
                   uint32_t* out_eid) {
#pragma omp parallel for schedule(dynamic, 64)
  for (uint64_t i = 0; i < n; i++) {
    pip_state st;
    st.best_y = INFINITY;
    st.best_eid = ORC_NO_HIT;
    st.best_a = 0;
    st.best_b = 1;
    for (uint64_t j = 0; j < nb; j++)
      pip_update(&st, query_map_id, pts[2 * i], pts[2 * i + 1],
                 xyb + 2 * (uint64_t) b_p1[j], (uint32_t) j);
    out_eid[i] = st.best_eid;
  }
}

/* Grid-walk PIP (same answers as orc_pip_brute; tests assert that).  Walks
 * the point's column upward like src/app/pip_grid.h:37-70 but stops only
 * when the best hit is provably below the top of the visited cell (the
 * double y* can differ from the exact crossing by < 1 unit).              */
void orc_pip_grid(const int64_t* xyb, const uint32_t* b_p1, uint64_t nb,
                  int64_t imin, int64_t irange, const int64_t* pts, uint64_t n,
                  int query_map_id, uint32_t* out_eid, uint64_t* n_candidates) {
  host_grid g;
  grid_build(&g, xyb, b_p1, nb, imin, irange);
  uint64_t ncand = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : ncand)
  for (uint64_t i = 0; i < n; i++) {
    int64_t px = pts[2 * i], py = pts[2 * i + 1];
    pip_state st;
    st.best_y = INFINITY;
    st.best_eid = ORC_NO_HIT;
    st.best_a = 0;
    st.best_b = 1;
    int cx = cell_of(&g, px);
    for (int cy = cell_of(&g, py - 1); cy < g.gsize; cy++) {
      uint64_t c = (uint64_t) cy * g.gsize + cx;
      for (uint64_t k = g.cell_begin[c]; k < g.cell_begin[c + 1]; k++) {
        uint32_t j = g.items[k];
        ncand++;
        pip_update(&st, query_map_id, px, py, xyb + 2 * (uint64_t) b_p1[j], j);
      }
      if (st.best_eid != ORC_NO_HIT) {
        double cell_top =
            (double) (g.imin + (((int64_t) cy + 1) << g.shift)) - 1.0;
        if (st.best_y < cell_top) break;
      }
    }
    out_eid[i] = st.best_eid;
  }
  grid_free(&g);
  if (n_candidates) *n_candidates = ncand;
}

/* face id seen from below the edge: src/map/map.h:79-87; EXTERIOR_FACE_ID
 * (0) when no edge is above: src/app/map_overlay_lbvh.h:96-104            */
void orc_face_ids(const int64_t* xyb, const uint32_t* b_p1,
                  const uint32_t* b_chain, const int64_t* left,
                  const int64_t* right, const uint32_t* eids, uint64_t n,
                  int32_t* out_face) {
  for (uint64_t i = 0; i < n; i++) {
    uint32_t e = eids[i];
    if (e == ORC_NO_HIT) {
      out_face[i] = 0;
      continue;
    }
    const int64_t* p = xyb + 2 * (uint64_t) b_p1[e];
    uint32_t c = b_chain[e];
    out_face[i] = (int32_t) (p[0] < p[2] ? right[c] : left[c]);
  }
}

/* bench.py pins torch to one thread per rank, which also lowers the OpenMP default of
 * this library when both share one libgomp: the CPU baseline asks for all cores again */
/* (double)(__int128) as this host compiles it (libgcc __floattidf, round to nearest even):
 * what tcb::rational's operator double (src/util/rational.h:94-97) and the PIP y*
 * (src/algo/pip.h:56-58) do per operand.  v = n x {lo, hi} words. */
void orc_i128_to_double(const uint64_t* v, uint64_t n, double* out) {
  for (uint64_t i = 0; i < n; i++) {
    i128 a = (i128) (((unsigned __int128) v[2 * i + 1] << 64) | v[2 * i]);
    out[i] = (double) a;
  }
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void) n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
