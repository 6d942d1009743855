"""Shared helpers of the parity tests: seeded datasets and oracle runs."""
import numpy as np

from rayjoin_b200 import synth
from rayjoin_b200.capi import PlanarGraph


def lattice_map(n_chains, seed, size=24, max_len=6, face_base=1):
    """Random polylines on a small integer lattice: shared vertices, T-junctions,
    overlapping horizontal / vertical edges -- every degenerate contact the
    simulation-of-simplicity branches of intersect_test exist for."""
    rng = np.random.default_rng(seed)
    xy, rows, left, right = [], [0], [], []
    for c in range(n_chains):
        n = int(rng.integers(2, max_len + 1))
        p = rng.integers(0, size, size=2)
        pts = [p]
        while len(pts) < n:
            step = rng.integers(-3, 4, size=2)
            q = np.clip(pts[-1] + step, 0, size - 1)
            if (q == pts[-1]).all():
                continue
            pts.append(q)
        xy.extend(pts)
        rows.append(len(xy))
        left.append(face_base + int(rng.integers(0, 5)))
        right.append(face_base + int(rng.integers(0, 5)))
    return PlanarGraph(np.asarray(xy, np.float64), np.asarray(rows, np.uint32),
                       np.asarray(left, np.int64), np.asarray(right, np.int64))


def dataset(name):
    if name == "voronoi":
        R = synth.voronoi_map(60, 3000, synth.BRAZIL_BBOX, seed=1)
        S = synth.voronoi_map(400, 5000, synth.BRAZIL_BBOX, seed=2)
    elif name == "shared":
        R = synth.voronoi_map(40, 2500, synth.BRAZIL_BBOX, seed=3)
        S = synth.share_chains(R, synth.voronoi_map(150, 3000, synth.BRAZIL_BBOX, seed=4), frac=0.3)
    elif name == "soup":
        R = synth.polygon_soup(3000, "gaussian", seed=1, polysize=0.2)
        S = synth.polygon_soup(2500, "gaussian", seed=2, polysize=0.2)
    elif name == "lattice":
        R = lattice_map(150, 11)
        S = lattice_map(170, 12, face_base=10)
    elif name == "aniso":
        # strongly anisotropic box -> very different rx / ry
        R = synth.voronoi_map(30, 1500, (-179.0, 10.0, 179.0, 12.0), seed=5)
        S = synth.voronoi_map(90, 2000, (-179.0, 10.0, 179.0, 12.0), seed=6)
    elif name == "dense":
        # many short edges (leaves touch ~6 occupancy cells): the regime of the headline workload,
        # where the occupancy filter and the cell directory are in play
        R = synth.voronoi_map(1500, 600000, synth.BRAZIL_BBOX, seed=7)
        S = synth.voronoi_map(5000, 900000, synth.BRAZIL_BBOX, seed=8)
    elif name == "tiny":
        R = PlanarGraph(np.array([[0.0, 0.0], [10.0, 10.0]]), [0, 2], [1], [2])
        S = PlanarGraph(np.array([[0.0, 10.0], [10.0, 0.0], [20.0, 3.0]]), [0, 3], [3], [4])
    else:
        raise KeyError(name)
    return R, S


DATASETS = ["voronoi", "shared", "soup", "lattice", "aniso", "tiny"]


class OracleMaps:
    """Both maps scaled and numbered by the oracle."""

    def __init__(self, O, graphs):
        self.O = O
        self.graphs = graphs
        self.bbox = synth.union_bbox(*graphs)
        self.sc = O.scaling_init(*self.bbox)
        self.pts = [O.scale_points(self.sc, g.xy) for g in graphs]
        e = [O.build_edges(g.row_index) for g in graphs]
        self.p1 = [x[0] for x in e]
        self.chain = [x[1] for x in e]

    def lsi(self, q, brute=False):
        b = 1 - q
        f = self.O.lsi_brute if brute else None
        if brute:
            return self.O.lsi_brute(self.pts[q], self.p1[q], self.pts[b], self.p1[b])
        return self.O.lsi_grid(self.pts[q], self.p1[q], self.pts[b], self.p1[b], self.sc)

    def lsi_refgrid(self, q, gsize, brute=False):
        """the reference's -mode=grid pair set, ordered like sort_xsects(xs, q)"""
        return self.O.lsi_refgrid(self.pts[0], self.p1[0], self.pts[1], self.p1[1], self.sc, gsize,
                                  sort_map=q, brute=brute)

    def pip(self, q, pts, brute=False):
        b = 1 - q
        if brute:
            return self.O.pip_brute(self.pts[b], self.p1[b], pts, q)
        return self.O.pip_grid(self.pts[b], self.p1[b], self.sc, pts, q)

    def faces(self, q, eids):
        b = 1 - q
        g = self.graphs[b]
        return self.O.face_ids(self.pts[b], self.p1[b], self.chain[b], g.left, g.right, eids)


def sort_xsects(xs, q):
    """rjb_xsect structured array -> arrays sorted by (eid[q], eid[1-q])."""
    eq, eb = xs["eid"][:, q].astype(np.int64), xs["eid"][:, 1 - q].astype(np.int64)
    o = np.lexsort((eb, eq))
    return eq[o].astype(np.uint32), eb[o].astype(np.uint32), xs["x"][o], xs["y"][o]


def i128_kat(n_random=200000, seed=99):
    """Known-answer inputs for (double)(__int128): magnitudes just above 2^53, 2^64 and 2^96,
    exact ties between two doubles (with even and odd mantissas), ties +-1, powers of two
    +-1, and random values of every bit length.  Returns (python ints, (n, 2) uint64 words)."""
    rng = np.random.default_rng(seed)
    vals = [0, 1, -1, 2**53, 2**53 + 1, 2**53 + 2, 2**53 + 3, -(2**53 + 1), 2**64, 2**64 + 1, 2**64 - 1,
            2**96, 2**96 + 1, 2**96 - 1, 2**127 - 1, -(2**127), -(2**127) + 1, 2**63, -(2**63), 2**63 - 1]
    for k in range(1, 75):           # v needs 53 + k bits
        half = 1 << (k - 1)
        for m in rng.integers(2**52, 2**53, size=40).tolist():
            for mm in (m, m | 1, m & ~1):
                base = mm << k
                for d in (0, half - 1, half, half + 1, (1 << k) - 1):
                    if d < 0 or base + d >= 2**127:
                        continue
                    vals.append(base + d)
                    vals.append(-(base + d))
    for b in (53, 54, 55, 63, 64, 65, 95, 96, 97, 126):
        for d in range(-3, 4):
            vals.append(2**b + d)
            vals.append(-(2**b + d))
    bits = rng.integers(1, 127, size=n_random)
    hi = rng.integers(0, 2**62, size=n_random).tolist()
    lo = rng.integers(0, 2**62, size=n_random).tolist()
    sg = rng.integers(0, 2, size=n_random).tolist()
    for b, h, l, s in zip(bits.tolist(), hi, lo, sg):
        v = ((h << 62) | l) & ((1 << b) - 1) | (1 << (b - 1))
        vals.append(-v if s else v)
    words = np.array([[(v & (2**64 - 1)), ((v >> 64) & (2**64 - 1))] for v in vals], dtype=np.uint64)
    return vals, words
