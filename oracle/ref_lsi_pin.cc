/*
 * ref_lsi_pin.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the REFERENCE's own LSI predicate and rational intersection point
 * (src/algo/lsi.h, src/util/rational.h, included from the read-only
 * /root/reference tree -- never copied) as plain host C++ and exports a
 * batch entry point, so the oracle restatement (oracle/oracle.c) can be
 * pinned against the real thing without a GPU.
 *
 * Built by oracle/Makefile into oracle/_ref/libref_lsi.so (git-ignored).
 * The call sequence mirrors the LBVH callback, src/app/lsi_lbvh.h:63-79:
 * intersect_test(e1,...,xsect_x,xsect_y) then `xsect.x = xsect_x`
 * into a dev::Intersection<int64_t>.
 */
#include <cstdint>
#include <cstdio>
#include <limits>

#include "config.h"
#include "algo/lsi.h"
#include "map/scaling.h"

namespace {
struct Pt {
  int64_t x, y;
  bool operator==(const Pt& o) const { return x == o.x && y == o.y; }
};
struct Eq {
  __int128 a, b, c;
};
// edge equation exactly as the device kernel fills it, src/map/map.h:216-226
inline Eq make_eq(const Pt& p1, const Pt& p2) {
  Eq e;
  e.a = p1.y - p2.y;
  e.b = p2.x - p1.x;
  e.c = -(__int128) p1.x * e.a - (__int128) p1.y * e.b;
  if (e.b < 0) {
    e.a = -e.a;
    e.b = -e.b;
    e.c = -e.c;
  }
  return e;
}
}  // namespace

extern "C" void ref_intersect_batch(const int64_t* pts, uint64_t n,
                                    uint8_t* hit, int64_t* x, int64_t* y,
                                    uint8_t* hit_pred_only) {
  for (uint64_t i = 0; i < n; i++) {
    const int64_t* p = pts + 8 * i;
    Pt a1{p[0], p[1]}, a2{p[2], p[3]}, b1{p[4], p[5]}, b2{p[6], p[7]};
    Eq e1 = make_eq(a1, a2), e2 = make_eq(b1, b2);
    tcb::rational<__int128> xx, yy;
    bool h = rayjoin::dev::intersect_test<Eq, Eq, Pt, __int128>(
        e1, a1, a2, e2, b1, b2, xx, yy);
    if (hit_pred_only)
      hit_pred_only[i] = rayjoin::dev::intersect_test<Eq, Eq, Pt, __int128>(
          e1, a1, a2, e2, b1, b2);
    hit[i] = h;
    x[i] = y[i] = 0;
    if (h) {
      rayjoin::dev::Intersection<int64_t> xsect;
      xsect.x = xx;
      xsect.y = yy;
      x[i] = xsect.x.num();
      y[i] = xsect.y.num();
      if (xsect.x.denom() != 1 || xsect.y.denom() != 1) {
        // would contradict SURVEY section 0; make it loud
        fprintf(stderr, "ref_lsi_pin: non-unit denominator\n");
        hit[i] = 2;
      }
    }
  }
}

// Scaling<double> of the reference (src/map/scaling.h:32-136) evaluated on the
// host: constants {rx, ry, rrx, rry, deltax, deltay, ddeltax, ddeltay} via
// Unscale/Scale probes are not accessible (private), so the pin exposes the
// observable behaviour: ScaleX/ScaleY (host arithmetic, no FMA) and
// UnscaleX/UnscaleY of caller-provided values.
extern "C" void ref_scaling_apply(const double* bbox, const double* xy, uint64_t n,
                                  int64_t* scaled, const int64_t* ixy, uint64_t m,
                                  double* unscaled, int64_t* limits) {
  rayjoin::BoundingBox<double> bb;
  bb.min_x = bbox[0];
  bb.min_y = bbox[1];
  bb.max_x = bbox[2];
  bb.max_y = bbox[3];
  rayjoin::Scaling<double> s(bb);
  for (uint64_t i = 0; i < n; i++) {
    scaled[2 * i] = s.ScaleX(xy[2 * i]);
    scaled[2 * i + 1] = s.ScaleY(xy[2 * i + 1]);
  }
  for (uint64_t i = 0; i < m; i++) {
    unscaled[2 * i] = s.UnscaleX(ixy[2 * i]);
    unscaled[2 * i + 1] = s.UnscaleY(ixy[2 * i + 1]);
  }
  limits[0] = s.get_internal_min();
  limits[1] = s.get_internal_max();
  limits[2] = s.get_internal_range();
}
