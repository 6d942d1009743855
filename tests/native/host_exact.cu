// Test-only: compiles the product's exact arithmetic (rayjoin_b200/csrc/rjb_exact.cuh,
// the SAME source the kernels use) for the HOST, so the CPU test-suite can pin it
// against the oracle on millions of cases without a GPU.  Not part of librjb200.
#include "rjb_exact.cuh"

using namespace rjb;

extern "C" {

// pts: n x 8 int64 {e1.x1, e1.y1, e1.x2, e1.y2, e2.x1, e2.y1, e2.x2, e2.y2}
// mode 0: always-exact path; 1: kDefer path, deferred items re-run like the kernel does
void hx_intersect_batch(const long long* pts, unsigned long long n, int mode, unsigned char* hit,
                        long long* ox, long long* oy, unsigned long long* n_deferred) {
  unsigned long long nd = 0;
  for (unsigned long long i = 0; i < n; i++) {
    const long long* p = pts + 8 * i;
    const Seg e1 = {p[0], p[1], p[2], p[3]}, e2 = {p[4], p[5], p[6], p[7]};
    hit[i] = lsi_intersect(e1, e2) ? 1 : 0;
    ox[i] = oy[i] = 0;
    if (!hit[i]) continue;
    if (mode == 2) {  // the path of k_lsi_resolve: both axes at once, parked states finished afterwards
      long long out[2];
      int code[2];
      PointState st[2];
      lsi_point_both(e1, e2, out, code, st);
      for (int axis = 0; axis < 2; axis++) {
        if (code[axis] != kPointDone) nd++;
        if (code[axis] == kPointGcd) out[axis] = lsi_point_finish(st[axis]);
        if (code[axis] == kPointRedo) out[axis] = lsi_point_axis<false>(e1, e2, axis, nullptr);
      }
      ox[i] = out[0];
      oy[i] = out[1];
      continue;
    }
    for (int axis = 0; axis < 2; axis++) {
      long long v;
      if (mode == 0) {
        v = lsi_point_axis(e1, e2, axis);
      } else {
        bool def = false;
        v = lsi_point_axis<true>(e1, e2, axis, &def);
        if (def) {
          nd++;
          v = lsi_point_axis<false>(e1, e2, axis, nullptr);
        }
      }
      (axis == 0 ? ox : oy)[i] = v;
    }
  }
  if (n_deferred) *n_deferred = nd;
}

// cell of the intersection point the way the grid LSI kernel computes it (lsi_xsect_ref_cell:
// gcd-free short-cut with the reference's literal sequence as the fall-back)
void hx_xsect_ref_cells(const long long* pts, unsigned long long n, long long imin, double cell_scale,
                        int* cx, int* cy) {
  for (unsigned long long i = 0; i < n; i++) {
    const long long* p = pts + 8 * i;
    const Seg e1 = {p[0], p[1], p[2], p[3]}, e2 = {p[4], p[5], p[6], p[7]};
    cx[i] = cy[i] = 0;
    if (!lsi_intersect(e1, e2)) continue;
    cx[i] = lsi_xsect_ref_cell(e1, e2, 0, imin, cell_scale);
    cy[i] = lsi_xsect_ref_cell(e1, e2, 1, imin, cell_scale);
  }
}

// occupancy cell code of a vertex and descriptor of an edge (what k_load_points writes and
// k_lsi_filter reads)
unsigned hx_occ_code(long long x, long long y) { return occ_code(x, y); }
unsigned hx_edge_desc(long long x1, long long y1, long long x2, long long y2) {
  return edge_desc_of(occ_code(x1, y1), occ_code(x2, y2));
}
int hx_occ_cell_of_quant(long long v) { return occ_cell((int) (v >> kQuantShift)); }
unsigned hx_tile_desc(unsigned x0, unsigned y0, unsigned x1, unsigned y1) { return tile_desc_of(x0, y0, x1, y1); }

// Cell-directory items (k_cell_lists / k_lsi_cells): for n pairs of quantised boxes {x0, y0, x1, y1}
// (leaf box, query box) counts in how many cells of the query's cell box, among the cells the
// leaf is registered in, cell_item_hit fires (out_hits) and records the last such cell.
void hx_cell_item_batch(const int* leaf, const int* query, unsigned long long n, unsigned* out_hits,
                        unsigned* out_cell, unsigned* out_first, unsigned* out_count) {
  for (unsigned long long i = 0; i < n; i++) {
    const int4 lb = make_int4(leaf[4 * i], leaf[4 * i + 1], leaf[4 * i + 2], leaf[4 * i + 3]);
    const int4 qb = make_int4(query[4 * i], query[4 * i + 1], query[4 * i + 2], query[4 * i + 3]);
    unsigned hits = 0, cell = 0, first = 0, count = 0;
    for (int cy = occ_cell(qb.y); cy <= occ_cell(qb.w); cy++)
      for (int cx = occ_cell(qb.x); cx <= occ_cell(qb.z); cx++) {
        if (cx < occ_cell(lb.x) || cx > occ_cell(lb.z) || cy < occ_cell(lb.y) || cy > occ_cell(lb.w)) continue;
        const uint4 it = cell_item_of(lb, cx, cy, 1000u + (unsigned) i, 1u + (unsigned) (i % 8));
        if (cell_item_hit(it, cell_clip(qb, cx, cy))) {
          hits++;
          cell = (unsigned) cy * kOccDim + (unsigned) cx;
          first = it.w;
          count = (it.z >> 14) + 1;
        }
      }
    out_hits[i] = hits;
    out_cell[i] = cell;
    out_first[i] = first;
    out_count[i] = count;
  }
}

// PIP: the update rule scanned over all edges in eid order for a batch of points
// (edges: nb x 4 int64 {x1, y1, x2, y2}; out: chosen edge index or 0xFFFFFFFF)
void hx_pip_batch(const long long* edges, unsigned long long nb, const long long* pts,
                  unsigned long long n, int q, unsigned* out) {
  for (unsigned long long i = 0; i < n; i++) {
    PipBest st;
    pip_init(st);
    for (unsigned long long j = 0; j < nb; j++) {
      const Seg e = {edges[4 * j], edges[4 * j + 1], edges[4 * j + 2], edges[4 * j + 3]};
      pip_update(st, q, pts[2 * i], pts[2 * i + 1], e, (uint32_t) j);
    }
    out[i] = st.eid;
  }
}

// PIP update rule over a list of edges for one point (returns the chosen index or -1)
int hx_pip_scan(const long long* edges, unsigned long long n, int q, long long px, long long py) {
  PipBest st;
  pip_init(st);
  for (unsigned long long i = 0; i < n; i++) {
    const Seg e = {edges[4 * i], edges[4 * i + 1], edges[4 * i + 2], edges[4 * i + 3]};
    pip_update(st, q, px, py, e, (uint32_t) i);
  }
  return st.eid == RJB_NO_HIT ? -1 : (int) st.eid;
}
}
