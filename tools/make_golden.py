"""Generates tests/golden/*.npz from the REFERENCE's own code compiled on the
host (oracle/_ref/libref_lsi.so: src/algo/lsi.h, src/util/rational.h,
src/map/scaling.h of /root/reference).  Run in the build container where
/root/reference exists:   python tools/make_golden.py
The committed vectors let the CPU test-suite pin the oracle where the reference
tree is absent (the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def lsi_cases(rng):
    cases = []
    cases.append(rng.integers(-1, 2, size=(3000, 8)))               # 3x3 lattice: all degeneracies
    cases.append(rng.integers(-3, 4, size=(6000, 8)))               # 7x7 lattice
    cases.append(rng.integers(-1000, 1000, size=(3000, 8)))
    base = rng.integers(-2**46, 2**46 - 2**31, size=(4000, 1, 2))   # 47-bit coordinates, short edges
    cases.append((base + rng.integers(0, 2**30, size=(4000, 4, 2))).reshape(-1, 8))
    base = rng.integers(-2**46, 2**46 - 2**42, size=(2000, 1, 2))   # medium edges
    cases.append((base + rng.integers(0, 2**41, size=(2000, 4, 2))).reshape(-1, 8))
    cases.append(rng.integers(-2**46, 2**46, size=(2000, 8)))       # map-spanning edges: int128 wraps
    # shared end points and T-junctions at full scale
    p = rng.integers(-2**45, 2**45, size=(2000, 8))
    p[:500, 4:6] = p[:500, 0:2]                                     # e2.p1 == e1.p1
    p[500:1000, 6:8] = p[500:1000, 2:4]                             # e2.p2 == e1.p2
    mid = (p[1000:1500, 0:2] + p[1000:1500, 2:4]) // 2              # e2.p1 (nearly) on e1
    p[1000:1500, 4:6] = mid
    p[1500:, 4:8] = p[1500:, 0:4]                                   # identical edges
    p[1750:, 4:8] = p[1750:, [2, 3, 0, 1]]                          # reversed identical edges
    cases.append(p)
    return np.concatenate(cases).astype(np.int64)


def main():
    assert O.ref_lsi_available(), "build oracle/_ref first (make -C oracle)"
    rng = np.random.default_rng(20240518)
    pts = lsi_cases(rng)
    hit, x, y = O.ref_intersect_batch(pts)
    assert hit.max() <= 1
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "lsi_kat.npz"), pts=pts, hit=hit, x=x, y=y)
    # scaling: three boxes (Brazil sample, US, a thin anisotropic one)
    boxes = np.array([[-74.0, -34.0, -34.0, 5.0], [-179.15, -14.55, 179.78, 71.39],
                      [-179.0, 10.0, 179.0, 12.0]])
    sc = {}
    for i, b in enumerate(boxes):
        xy = np.column_stack([rng.uniform(b[0], b[2], 2000), rng.uniform(b[1], b[3], 2000)])
        ixy = rng.integers(-2**46, 2**46, size=(2000, 2))
        scaled, unscaled, limits = O.ref_scaling_apply(b, xy, ixy)
        sc["xy%d" % i], sc["ixy%d" % i] = xy, ixy
        sc["scaled_host%d" % i], sc["unscaled_host%d" % i], sc["limits%d" % i] = scaled, unscaled, limits
    np.savez_compressed(os.path.join(out, "scaling_kat.npz"), boxes=boxes, **sc)
    print("golden: %d LSI cases (%d hits), %d scaling boxes" % (len(pts), int(hit.sum()), len(boxes)))


if __name__ == "__main__":
    main()
