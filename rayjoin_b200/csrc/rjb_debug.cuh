// Test hooks: the exact arithmetic of the query kernels, run on the DEVICE over caller-supplied
// cases.  The kernels below call the very functions k_lsi_exact / k_lsi_points / k_pip_bvh call
// (rjb_exact.cuh), so the golden vectors of the reference's lsi.h / rational.h and
// known-answer tests for (double)(__int128) go through the sm_100a compile, not only through
// the host compile of the same header (tests/native/host_exact.cu).
#pragma once
#include "rjb_exact.cuh"

namespace rjb {

// pts: n x 8 int64 {e1.x1, e1.y1, e1.x2, e1.y2, e2.x1, e2.y1, e2.x2, e2.y2}; e1 = query side.
// mode 0: lsi_intersect + lsi_point_axis<false> (always through the gcd);
// mode 1: the path of k_lsi_points: lsi_point_axis<true>, deferred coordinates re-run with <false>;
// mode 2: the path of k_lsi_resolve: lsi_point_both, parked states through lsi_point_finish.
// flags[i]: bit 0 = intersects, bit 1 / 2 = x / y was deferred.
__global__ void k_debug_intersect(const long long* __restrict__ pts, uint64_t n, int mode,
                                  unsigned char* __restrict__ flags, long long* __restrict__ ox,
                                  long long* __restrict__ oy) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long* p = pts + 8 * i;
  const Seg e1 = {p[0], p[1], p[2], p[3]}, e2 = {p[4], p[5], p[6], p[7]};
  unsigned char f = lsi_intersect(e1, e2) ? 1 : 0;
  long long x = 0, y = 0;
  if (f) {
    if (mode == 0) {
      x = lsi_point_axis<false>(e1, e2, 0, nullptr);
      y = lsi_point_axis<false>(e1, e2, 1, nullptr);
    } else if (mode == 2) {
      long long out[2];
      int code[2];
      PointState st[2];
      lsi_point_both(e1, e2, out, code, st);
      for (int axis = 0; axis < 2; axis++) {
        if (code[axis] != kPointDone) f |= 2 << axis;
        if (code[axis] == kPointGcd) out[axis] = lsi_point_finish(st[axis]);
        if (code[axis] == kPointRedo) out[axis] = lsi_point_axis<false>(e1, e2, axis, nullptr);
      }
      x = out[0];
      y = out[1];
    } else {
      bool dx = false, dy = false;
      x = lsi_point_axis<true>(e1, e2, 0, &dx);
      y = lsi_point_axis<true>(e1, e2, 1, &dy);
      if (dx) { x = lsi_point_axis<false>(e1, e2, 0, nullptr); f |= 2; }
      if (dy) { y = lsi_point_axis<false>(e1, e2, 1, nullptr); f |= 4; }
    }
  }
  flags[i] = f;
  ox[i] = x;
  oy[i] = y;
}

// v[i] = {lo, hi} words of a signed 128-bit integer, d[i] likewise (d may be null).
// out_cvt[i] = (double) v; out_div[i] = (double) v / (double) d; out_trunc[i] = (int64) of that
// quotient when it is in range -- the conversions PIP's y* and the rational -> int64 store use.
__global__ void k_debug_i128(const unsigned long long* __restrict__ v, const unsigned long long* __restrict__ d,
                             uint64_t n, double* __restrict__ out_cvt, double* __restrict__ out_div,
                             long long* __restrict__ out_trunc) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const i128 a = (i128) (((u128) v[2 * i + 1] << 64) | v[2 * i]);
  const double da = (double) a;
  out_cvt[i] = da;
  if (d) {
    const i128 b = (i128) (((u128) d[2 * i + 1] << 64) | d[2 * i]);
    const double q = da / (double) b;
    out_div[i] = q;
    out_trunc[i] = (q > -9.2e18 && q < 9.2e18) ? (long long) q : 0;
  }
}

// PIP update rule scanned over nb edges in eid order, one thread per point
// (edges: nb x 4 int64 {x1, y1, x2, y2})
__global__ void k_debug_pip(const long long* __restrict__ edges, uint32_t nb, const long long* __restrict__ pts,
                            uint64_t n, int q, uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PipBest st;
  pip_init(st);
  const long long px = pts[2 * i], py = pts[2 * i + 1];
  for (uint32_t j = 0; j < nb; j++) {
    const Seg e = {edges[4 * j], edges[4 * j + 1], edges[4 * j + 2], edges[4 * j + 3]};
    pip_update(st, q, px, py, e, j);
  }
  out[i] = st.eid;
}

}  // namespace rjb
