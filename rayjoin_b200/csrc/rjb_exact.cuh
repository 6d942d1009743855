// Exact fixed-point predicates (device).  These reproduce RayJoin's results
// bit for bit; they are written from the arithmetic definition, not from the
// reference's AoS Edge{a,b,c} objects:
//
//   S(p, e) = a*p.x + b*p.y + c  with  c = -x1*a - y1*b   (src/map/map.h:216-226)
//           = a*(p.x - x1) + b*(p.y - y1)                 (exact integers)
//
// so the edge equation never has to be stored: a, b come from the endpoints
// (sign-normalised so that b >= 0, or a unchanged when b == 0) and the 2^95
// constant term disappears.  All products are < 2^96 -> int128 is exact.
#pragma once
#include "rjb_common.cuh"

namespace rjb {

typedef __int128 i128;

struct Seg {
  long long x1, y1, x2, y2;
};

// (a, b) of the edge equation, normalised like src/map/map.h:222-226
static RJB_HD void edge_ab(const Seg& e, long long& a,
                                               long long& b) {
  a = e.y1 - e.y2;
  b = e.x2 - e.x1;
  if (b < 0) {
    a = -a;
    b = -b;
  }
}

// sign of a*dx + b*dy (|a|,|b|,|dx|,|dy| < 2^48): -1, 0, +1
static RJB_HD int side_sign(long long a, long long b,
                                                long long dx, long long dy) {
  // 32-bit fast path: one IMAD.WIDE per product, the sum cannot overflow
  const long long lim = 1ll << 31;
  if ((unsigned long long) (a + lim) < (1ull << 32) &&
      (unsigned long long) (b + lim) < (1ull << 32) &&
      (unsigned long long) (dx + lim) < (1ull << 32) &&
      (unsigned long long) (dy + lim) < (1ull << 32)) {
    long long s = (long long) (int) a * (int) dx + (long long) (int) b * (int) dy;
    return (s > 0) - (s < 0);
  }
  i128 s = (i128) a * dx + (i128) b * dy;
  return (s > 0) - (s < 0);
}

static RJB_HD int sgn(long long v) { return (v > 0) - (v < 0); }

// intersect_test(e1, e2), src/algo/lsi.h:27-103.  e1 is the QUERY-side edge,
// e2 the BASE-side edge (src/app/lsi_lbvh.h:71); the simulation-of-simplicity
// perturbations are not symmetric in (e1, e2).
static RJB_HD bool lsi_intersect(const Seg& e1, const Seg& e2) {
  long long a1, b1, a2, b2;
  edge_ab(e1, a1, b1);
  edge_ab(e2, a2, b2);
  // endpoints of e1 against the line of e2 (lsi.h:42-67)
  int s11 = side_sign(a2, b2, e1.x1 - e2.x1, e1.y1 - e2.y1);
  if (s11 == 0) s11 = -sgn(a2);
  if (s11 == 0) s11 = -sgn(b2);
  if (s11 == 0) return false;
  int s12 = side_sign(a2, b2, e1.x2 - e2.x1, e1.y2 - e2.y1);
  if (s12 == 0) s12 = -sgn(a2);
  if (s12 == 0) s12 = -sgn(b2);
  if (s12 == 0) return false;
  if (s11 == s12) return false;
  // endpoints of e2 against the line of e1 (lsi.h:70-91)
  int s21 = side_sign(a1, b1, e2.x1 - e1.x1, e2.y1 - e1.y1);
  if (s21 == 0) s21 = sgn(a1);
  if (s21 == 0) s21 = sgn(b1);
  if (s21 == 0) return false;
  int s22 = side_sign(a1, b1, e2.x2 - e1.x1, e2.y2 - e1.y1);
  if (s22 == 0) s22 = sgn(a1);
  if (s22 == 0) s22 = sgn(b1);
  if (s22 == 0) return false;
  if (s21 == s22) return false;
  // identical edges (lsi.h:97-100)
  if ((e1.x1 == e2.x1 && e1.y1 == e2.y1 && e1.x2 == e2.x2 && e1.y2 == e2.y2) ||
      (e1.x1 == e2.x2 && e1.y1 == e2.y2 && e1.x2 == e2.x1 && e1.y2 == e2.y1))
    return false;
  return true;
}

// closed integer boxes of two segments overlap.  intersect_test() == true
// implies this (the crossing point lies on both closed segments), so it is a
// loss-free prefilter in exact arithmetic.
static RJB_HD bool seg_boxes_overlap(const Seg& p, const Seg& q) {
  return min(p.x1, p.x2) <= max(q.x1, q.x2) && min(q.x1, q.x2) <= max(p.x1, p.x2) &&
         min(p.y1, p.y2) <= max(q.y1, q.y2) && min(q.y1, q.y2) <= max(p.y1, p.y2);
}

// ---- intersection point ----------------------------------------------------
// src/algo/lsi.h:117-141 + tcb::rational (src/util/rational.h): the point is
// rational(num, den) gcd-reduced in int128 (wrapping like the device code of
// the reference does for very long edges), clamped to the bbox of the four
// endpoints, then converted with (int64)((double)num / (double)den).
static RJB_HD i128 iabs128(i128 v) { return v < 0 ? -v : v; }

typedef unsigned __int128 u128;

static RJB_HD int ctz64(unsigned long long v) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long) v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}

// binary gcd of two 64-bit magnitudes (no hardware divide on the GPU: Euclid's
// `%` would cost ~100 instructions per step)
static RJB_HD unsigned long long gcd64(unsigned long long a,
                                                           unsigned long long b) {
  if (a == 0) return b;
  if (b == 0) return a;
  int sh = ctz64(a | b);
  a >>= ctz64(a);
  do {
    b >>= ctz64(b);
    if (a > b) {
      unsigned long long t = a;
      a = b;
      b = t;
    }
    b -= a;
  } while (b != 0);
  return a << sh;
}

static RJB_HD int ctz128(u128 v) {
  unsigned long long lo = (unsigned long long) v;
  return lo ? ctz64(lo) : 64 + ctz64((unsigned long long) (v >> 64));
}

// magnitude of gcd(a, b); any algorithm gives the same value as the reference's
// Euclid loop (src/util/rational.h:36-43).  Binary gcd in 128 bits until both
// operands fit 64 bits (typically after a few dozen steps: |den| < 2^64).
static __host__ __device__ i128 gcd128(i128 a, i128 b) {
  u128 x = (u128) iabs128(a), y = (u128) iabs128(b);
  if (x == 0) return (i128) y;
  if (y == 0) return (i128) x;
  int sh = ctz128(x | y);
  x >>= ctz128(x);
  while (true) {
    y >>= ctz128(y);
    if ((x >> 64) == 0 && (y >> 64) == 0)
      return (i128) ((u128) gcd64((unsigned long long) x, (unsigned long long) y) << sh);
    if (x > y) {
      u128 t = x;
      x = y;
      y = t;
    }
    y -= x;
    if (y == 0) return (i128) (x << sh);
  }
}

static RJB_HD void rat_make(i128 num, i128 den, i128& on,
                                                i128& od) {
  i128 g = gcd128(num, den);  // den != 0 for a true intersection
  i128 sn = den < 0 ? -num : num;
  i128 ad = iabs128(den);
  if (g == 1) {  // the common case: skip two software 128-bit divisions
    on = sn;
    od = ad;
  } else {
    on = sn / g;
    od = ad / g;
  }
}

static RJB_HD long long min4ll(long long a, long long b,
                                                   long long c, long long d) {
  return min(min(a, b), min(c, d));
}
static RJB_HD long long max4ll(long long a, long long b,
                                                   long long c, long long d) {
  return max(max(a, b), max(c, d));
}

// One coordinate (axis 0 = x, 1 = y) of the intersection point.
//
// kDefer == false: always returns the reference's value.
// kDefer == true : returns it when the gcd-free path decides (*deferred = false);
//                  otherwise sets *deferred and returns 0 -- the caller re-runs the
//                  item with kDefer == false in a dense pass of its own.
//
// Why most points need no gcd.  The reference reduces num/den by their gcd g, clamps to
// the bbox [lo, hi] of the four endpoints and stores trunc(fl(fl(num/g) / fl(den/g)))
// (rational.h:190-203, lsi.h:124-141).  g only changes WHICH doubles are divided, never
// the rational: three roundings of <= 2^-53 relative error each on a value of magnitude
// <= 2^46 move the quotient by < 0.0235.  So when the exact value x = X0 + rs/D
// (X0 integer, 0 <= rs < D) has its fraction in [1/32, 31/32], every g gives
// trunc(x); when rs == 0 the reduced rational is X0/1 and the division is exact; the
// clamp compares exact rationals with integers and needs no g either.  Only fractions
// within 1/32 of an integer (6 % of the points) go through the gcd.
template <bool kDefer>
static __host__ __device__ long long lsi_point_axis(const Seg& e1, const Seg& e2, int axis,
                                                    bool* deferred) {
  long long a1l, b1l, a2l, b2l;
  edge_ab(e1, a1l, b1l);
  edge_ab(e2, a2l, b2l);
  const long long lo = axis == 0 ? min4ll(e1.x1, e1.x2, e2.x1, e2.x2) : min4ll(e1.y1, e1.y2, e2.y1, e2.y2);
  const long long hi = axis == 0 ? max4ll(e1.x1, e1.x2, e2.x1, e2.x2) : max4ll(e1.y1, e1.y2, e2.y1, e2.y2);
  const i128 a1 = a1l, b1 = b1l, a2 = a2l, b2 = b2l;
  // unsigned products: two's-complement wrap-around, no UB
  const i128 denom = (i128) ((u128) a1 * (u128) b2 - (u128) a2 * (u128) b1);
  i128 rn, rd;
  const long long lim = 1ll << 38;
  if (a1l > -lim && a1l < lim && b1l < lim && a2l > -lim && a2l < lim && b2l < lim) {
    // Edges spanning < 2^38 internal units (1/512 of the coordinate range): nothing
    // wraps.  With t = e1.p1 the shifted numerators are (x - x1)*den = c2'*b1 and
    // (y - y1)*den = -a1*c2', where c2' is e2's constant term in coordinates relative
    // to e1.p1.  (x - x1) and (y - y1) are bounded by e1's span (< 2^38), so one double
    // division yields the quotient to +-1.
    const i128 aden = iabs128(denom);
    const i128 c2p = -((i128) (e2.x1 - e1.x1) * a2l + (i128) (e2.y1 - e1.y1) * b2l);
    const i128 np = axis == 0 ? c2p * b1l : -c2p * a1l;
    const long long q = (long long) rint((double) np / (double) denom);
    const i128 r = np - (i128) q * denom;  // |r| <= (1/2 + eps)*|den|
    // exact value = X0 + rs/aden with 0 <= rs < aden
    long long X0 = (axis == 0 ? e1.x1 : e1.y1) + q;
    i128 rs = denom < 0 ? -r : r;
    while (rs < 0) { rs += aden; X0--; }
    while (rs >= aden) { rs -= aden; X0++; }
    if (X0 < lo) return lo;                          // x < lo  <=>  floor(x) < lo
    if (X0 > hi || (X0 == hi && rs != 0)) return hi; // x > hi
    if (rs == 0) return X0;                          // the reduced rational is X0 / 1
    // (the error bound above assumes |x| <= 2^46, the range Scaling produces)
    if (32 * rs >= aden && 32 * (aden - rs) >= aden && lo >= -(1ll << 46) && hi <= (1ll << 46))
      return X0 >= 0 ? X0 : X0 + 1;  // trunc(x)
    if (kDefer) {
      *deferred = true;
      return 0;
    }
    // gcd(num, den) = gcd(num - t*den, den) for any integer t: the binary gcd starts from
    // ~2k-bit operands instead of a ~110-bit numerator.  The reduced rational follows from
    // x = X0 + rs/aden without dividing that numerator either: num/g = X0 * (aden/g) + rs/g
    // (exact: g divides rs and aden), two 64-bit divisions when aden fits 64 bits -- it does
    // unless both edges span more than 2^31 units -- instead of two software 128-bit ones.
    const i128 g = gcd128(rs, aden);
    if (g == 1) {
      rd = aden;
      rn = (i128) X0 * aden + rs;
    } else if ((aden >> 64) == 0) {
      const unsigned long long g64 = (unsigned long long) g;
      const unsigned long long rd64 = (unsigned long long) aden / g64;
      rd = (i128) (u128) rd64;
      rn = (i128) X0 * (i128) (u128) rd64 + (i128) (u128) ((unsigned long long) rs / g64);
    } else {
      rd = aden / g;
      rn = (i128) X0 * rd + rs / g;
    }
  } else {
    if (kDefer) {
      *deferred = true;
      return 0;
    }
    // c with the SAME normalisation sign as (a, b): c = -x1*a - y1*b
    const i128 c1 = -(i128) e1.x1 * a1 - (i128) e1.y1 * b1;
    const i128 c2 = -(i128) e2.x1 * a2 - (i128) e2.y1 * b2;
    const i128 num = axis == 0 ? (i128) ((u128) c2 * (u128) b1 - (u128) c1 * (u128) b2)
                               : (i128) ((u128) a2 * (u128) c1 - (u128) a1 * (u128) c2);
    rat_make(num, denom, rn, rd);
  }
  if (rn < (i128) ((u128) (i128) lo * (u128) rd)) { rn = lo; rd = 1; }
  if ((i128) ((u128) (i128) hi * (u128) rd) < rn) { rn = hi; rd = 1; }
  return (long long) ((double) rn / (double) rd);
}

// ---- both coordinates at once, deferred coordinates with their state ------------------------
// The fused kernel (k_lsi_resolve) computes x and y of a hit together: edge equations, the
// denominator and e2's constant term are shared between the axes (a third of the
// instructions of two lsi_point_axis calls), and a coordinate that needs the gcd is parked
// WITH what has been computed -- x = X0 + rs / aden -- so that the dense pass starts at the gcd
// instead of re-loading four vertices and re-deriving everything.  Same values as
// lsi_point_axis (the CPU and GPU tests run both against the oracle).
struct PointState {
  long long X0;
  unsigned long long rs, aden;  // 0 < rs < aden < 2^64
};

constexpr int kPointDone = 0, kPointGcd = 1, kPointRedo = 2;

// the gcd path from a parked state: reduce rs / aden, assemble num/g = X0 * (aden/g) + rs/g
static RJB_HD long long lsi_point_finish(const PointState& st) {
  const unsigned long long g = gcd64(st.rs, st.aden);
  const unsigned long long rd = g == 1 ? st.aden : st.aden / g;
  const unsigned long long rr = g == 1 ? st.rs : st.rs / g;
  const i128 rn = (i128) st.X0 * (i128) (u128) rd + (i128) (u128) rr;
  return (long long) ((double) rn / (double) (i128) (u128) rd);
}

// out[axis] = the coordinate when code[axis] == kPointDone; kPointGcd: st[axis] is filled, finish
// with lsi_point_finish; kPointRedo: long edges (or a 65+ bit denominator), run lsi_point_axis<false>
static __host__ __device__ void lsi_point_both(const Seg& e1, const Seg& e2, long long out[2], int code[2],
                                               PointState st[2]) {
  long long a1l, b1l, a2l, b2l;
  edge_ab(e1, a1l, b1l);
  edge_ab(e2, a2l, b2l);
  const long long lim = 1ll << 38;
  if (!(a1l > -lim && a1l < lim && b1l < lim && a2l > -lim && a2l < lim && b2l < lim)) {
    code[0] = code[1] = kPointRedo;
    out[0] = out[1] = 0;
    return;
  }
  const i128 denom = (i128) a1l * b2l - (i128) a2l * b1l;  // < 2^77: nothing wraps here
  const i128 aden = iabs128(denom);
  const i128 c2p = -((i128) (e2.x1 - e1.x1) * a2l + (i128) (e2.y1 - e1.y1) * b2l);
  const double dden = (double) denom;
#pragma unroll
  for (int axis = 0; axis < 2; axis++) {
    const long long lo = axis == 0 ? min4ll(e1.x1, e1.x2, e2.x1, e2.x2) : min4ll(e1.y1, e1.y2, e2.y1, e2.y2);
    const long long hi = axis == 0 ? max4ll(e1.x1, e1.x2, e2.x1, e2.x2) : max4ll(e1.y1, e1.y2, e2.y1, e2.y2);
    const i128 np = axis == 0 ? c2p * b1l : -c2p * a1l;
    const long long q = (long long) rint((double) np / dden);
    const i128 r = np - (i128) q * denom;
    long long X0 = (axis == 0 ? e1.x1 : e1.y1) + q;
    i128 rs = denom < 0 ? -r : r;
    while (rs < 0) { rs += aden; X0--; }
    while (rs >= aden) { rs -= aden; X0++; }
    code[axis] = kPointDone;
    if (X0 < lo) out[axis] = lo;
    else if (X0 > hi || (X0 == hi && rs != 0)) out[axis] = hi;
    else if (rs == 0) out[axis] = X0;
    else if (32 * rs >= aden && 32 * (aden - rs) >= aden && lo >= -(1ll << 46) && hi <= (1ll << 46))
      out[axis] = X0 >= 0 ? X0 : X0 + 1;
    else if ((aden >> 64) == 0) {
      out[axis] = 0;
      code[axis] = kPointGcd;
      st[axis].X0 = X0;
      st[axis].rs = (unsigned long long) rs;
      st[axis].aden = (unsigned long long) aden;
    } else {
      out[axis] = 0;
      code[axis] = kPointRedo;
    }
  }
}

static __host__ __device__ long long lsi_point_axis(const Seg& e1, const Seg& e2, int axis) {
  return lsi_point_axis<false>(e1, e2, axis, nullptr);
}

// out of line: the rare "redo" path of the fused kernel (long edges) must not bloat its hot code
static __host__ __device__ __noinline__ long long lsi_point_axis_slow(const Seg& e1, const Seg& e2, int axis) {
  return lsi_point_axis<false>(e1, e2, axis, nullptr);
}

static __host__ __device__ void lsi_point(const Seg& e1, const Seg& e2, long long& ox, long long& oy) {
  ox = lsi_point_axis(e1, e2, 0);
  oy = lsi_point_axis(e1, e2, 1);
}

// ---- cell of the intersection point, as the reference's grid LSI computes it ----------
// LSIGrid keeps a pair only in the cell of its intersection point (src/app/lsi_grid.h:62-66):
// calculate_cell(gsize, scaling, xsect_x) with xsect_x the gcd-reduced, clamped
// tcb::rational<__int128> (src/grid/cell.h:15-22).  `val - internal_min` is again a rational
// (n' = num - imin * den over den), `* cell_scale` has no rational overload, so the rational
// converts through operator double: cell = (int) (fl(fl(n') / fl(d')) * cell_scale).
//
// For an integer coordinate v the same function gives (int) ((double) (v - imin) * cell_scale).
static RJB_HD int ref_cell(long long v, long long imin, double cell_scale) {
  return (int) ((double) (v - imin) * cell_scale);
}

// Exact value V = X0 + rs / aden of the (clamped) point is known without the gcd (see
// lsi_point_axis); the reference's three roundings move V * cell_scale by < 2^-36 cells and
// the short-cut below by < 2^-37, so unless the product lies within 10^-6 of a cell boundary
// every evaluation order truncates to the same cell.  Otherwise -- and for map-spanning edges,
// whose 128-bit products wrap -- the reference's sequence is replayed literally.
static __host__ __device__ int lsi_xsect_ref_cell(const Seg& e1, const Seg& e2, int axis, long long imin,
                                                  double cell_scale) {
  long long a1l, b1l, a2l, b2l;
  edge_ab(e1, a1l, b1l);
  edge_ab(e2, a2l, b2l);
  const long long lo = axis == 0 ? min4ll(e1.x1, e1.x2, e2.x1, e2.x2) : min4ll(e1.y1, e1.y2, e2.y1, e2.y2);
  const long long hi = axis == 0 ? max4ll(e1.x1, e1.x2, e2.x1, e2.x2) : max4ll(e1.y1, e1.y2, e2.y1, e2.y2);
  const i128 a1 = a1l, b1 = b1l, a2 = a2l, b2 = b2l;
  const i128 denom = (i128) ((u128) a1 * (u128) b2 - (u128) a2 * (u128) b1);
  const long long lim = 1ll << 38;
  if (a1l > -lim && a1l < lim && b1l < lim && a2l > -lim && a2l < lim && b2l < lim &&
      lo >= -(1ll << 46) && hi <= (1ll << 46)) {
    const i128 aden = iabs128(denom);
    const i128 c2p = -((i128) (e2.x1 - e1.x1) * a2l + (i128) (e2.y1 - e1.y1) * b2l);
    const i128 np = axis == 0 ? c2p * b1l : -c2p * a1l;
    const long long q = (long long) rint((double) np / (double) denom);
    const i128 r = np - (i128) q * denom;
    long long X0 = (axis == 0 ? e1.x1 : e1.y1) + q;
    i128 rs = denom < 0 ? -r : r;
    while (rs < 0) { rs += aden; X0--; }
    while (rs >= aden) { rs -= aden; X0++; }
    if (X0 < lo) return ref_cell(lo, imin, cell_scale);
    if (X0 > hi || (X0 == hi && rs != 0)) return ref_cell(hi, imin, cell_scale);
    if (rs == 0) return ref_cell(X0, imin, cell_scale);
    const double approx = ((double) (X0 - imin) + (double) rs / (double) aden) * cell_scale;
    const double fr = approx - floor(approx);
    if (fr > 1e-6 && fr < 1.0 - 1e-6) return (int) approx;
  }
  // the reference's own sequence
  const i128 c1 = -(i128) e1.x1 * a1 - (i128) e1.y1 * b1;
  const i128 c2 = -(i128) e2.x1 * a2 - (i128) e2.y1 * b2;
  const i128 num = axis == 0 ? (i128) ((u128) c2 * (u128) b1 - (u128) c1 * (u128) b2)
                             : (i128) ((u128) a2 * (u128) c1 - (u128) a1 * (u128) c2);
  i128 rn, rd;
  rat_make(num, denom, rn, rd);
  if (rn < (i128) ((u128) (i128) lo * (u128) rd)) { rn = lo; rd = 1; }
  if ((i128) ((u128) (i128) hi * (u128) rd) < rn) { rn = hi; rd = 1; }
  // val - internal_min: rational{num * 1 - imin * den, den * 1}, simplified again
  const i128 n1 = (i128) ((u128) rn - (u128) (i128) imin * (u128) rd);
  i128 n2, d2;
  rat_make(n1, rd, n2, d2);
  return (int) (((double) n2 / (double) d2) * cell_scale);
}

// ---- PIP: closest edge above -----------------------------------------------
// Update rule of src/algo/pip.h:27-96 == src/app/pip_lbvh.h:57-123.  Full ties
// (same y*, same slope = coincident edges) are resolved like a scan in
// increasing eid order: q == 1 keeps the smallest eid, q == 0 the largest.
struct PipBest {
  double y;        // best_y
  long long a, b;  // equation of the best edge (|.| < 2^48)
  uint32_t eid;
};

static RJB_HD void pip_init(PipBest& st) {
  st.y = __builtin_huge_val();  // +inf
  st.a = 0;
  st.b = 1;
  st.eid = RJB_NO_HIT;
}

// returns true when st changed
static RJB_HD bool pip_update(PipBest& st, int q, long long px,
                                                  long long py, const Seg& e,
                                                  uint32_t eid) {
  long long x_min = min(e.x1, e.x2), x_max = max(e.x1, e.x2);
  if (px < x_min || px > x_max || px == (q == 0 ? x_min : x_max)) return false;
  long long a, b;
  edge_ab(e, a, b);
  // -a*px - c = a*(x1 - px) + b*y1   (exact, < 2^97)
  i128 num = (i128) a * (e.x1 - px) + (i128) b * e.y1;
  double ys = (double) num / (double) b;
  double diff = (double) py - ys;
  if (diff == 0) diff = (double) (q == 0 ? -a : a);
  if (diff == 0) diff = (double) (q == 0 ? -b : b);
  if (diff > 0) return false;
  if (ys > st.y) return false;
  if (ys == st.y) {
    double cur = (double) a / (double) b;
    double best = (double) st.a / (double) st.b;
    if (cur == best) {
      if (q ? (eid > st.eid) : (eid < st.eid)) return false;
    } else {
      bool flag = cur > best;
      if ((q && !flag) || (flag && !q)) return false;
    }
  }
  st.y = ys;
  st.a = a;
  st.b = b;
  st.eid = eid;
  return true;
}

}  // namespace rjb
