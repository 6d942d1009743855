"""Seeded synthetic planar maps in RayJoin's chain (CDB) model.

RayJoin's own generator (reference misc/generator.py, driven by
misc/gen_polys.sh:4-22) only emits WKT/CSV polygons and needs an external
ArcGIS step to become CDB, and the bundled sample pair is missing from the
reference checkout, so the benchmark maps are synthesised here:

  voronoi_map    a planar subdivision (Voronoi cells = faces, ridges = chains
                 with left/right face ids) whose ridges are subdivided into
                 many short wiggly edges -- "county-like" (few faces, long
                 chains) or "zip-like" (many faces, short chains);
  polygon_soup   independent small polygons, one closed chain each, uniform or
                 gaussian centres -- the shape of gen_polys.sh's data
                 (radius 0.001, 3..10 vertices, affine 50,0,-119,0,30,35).
"""
import numpy as np

from .capi import PlanarGraph

US_BBOX = (-179.15, -14.55, 179.78, 71.39)     # County / Zipcode extent
BRAZIL_BBOX = (-74.0, -34.0, -34.0, 5.0)       # extent of the missing bundled sample (C1)


def voronoi_map(n_faces, n_edges, bbox=US_BBOX, seed=1, wiggle=0.03, face_id_base=1):
    """Voronoi subdivision of `bbox` with ~n_faces faces and ~n_edges edges."""
    from scipy.spatial import Voronoi
    rng = np.random.default_rng(seed)
    x0, y0, x1, y1 = bbox
    w, h = x1 - x0, y1 - y0
    pts = np.column_stack([x0 + rng.random(n_faces) * w, y0 + rng.random(n_faces) * h])
    # mirror the sites across the four sides: every original cell becomes bounded
    # and is clipped exactly to the box
    mir = [pts,
           np.column_stack([2 * x0 - pts[:, 0], pts[:, 1]]),
           np.column_stack([2 * x1 - pts[:, 0], pts[:, 1]]),
           np.column_stack([pts[:, 0], 2 * y0 - pts[:, 1]]),
           np.column_stack([pts[:, 0], 2 * y1 - pts[:, 1]])]
    vor = Voronoi(np.vstack(mir))
    rp = np.asarray(vor.ridge_points)
    rv = np.asarray(vor.ridge_vertices)
    keep = (rv[:, 0] >= 0) & (rv[:, 1] >= 0) & ((rp[:, 0] < n_faces) | (rp[:, 1] < n_faces))
    rp, rv = rp[keep], rv[keep]
    a = vor.vertices[rv[:, 0]]
    b = vor.vertices[rv[:, 1]]
    d = b - a
    length = np.hypot(d[:, 0], d[:, 1])
    ok = length > 1e-9 * max(w, h)
    rp, a, b, d, length = rp[ok], a[ok], b[ok], d[ok], length[ok]
    n_r = len(rp)
    # which site is on the left of a -> b
    s0 = vor.points[rp[:, 0]]
    cross = d[:, 0] * (s0[:, 1] - a[:, 1]) - d[:, 1] * (s0[:, 0] - a[:, 0])
    left_site = np.where(cross > 0, rp[:, 0], rp[:, 1])
    right_site = np.where(cross > 0, rp[:, 1], rp[:, 0])
    face = lambda s: np.where(s < n_faces, s + face_id_base, 0).astype(np.int64)
    left, right = face(left_site), face(right_site)
    # edges per ridge proportional to its length
    k = np.maximum(1, np.rint(length / length.sum() * n_edges)).astype(np.int64)
    npts = k + 1
    row_index = np.zeros(n_r + 1, np.int64)
    np.cumsum(npts, out=row_index[1:])
    total = int(row_index[-1])
    ridge = np.repeat(np.arange(n_r), npts)
    j = np.arange(total) - row_index[ridge]
    t = j / k[ridge]
    # smooth perpendicular displacement, zero at both ends: a few sinusoids per ridge
    nh = 4
    amp = rng.random((n_r, nh)) * (wiggle / nh)
    freq = rng.integers(1, 9, size=(n_r, nh)) * (1 + k[:, None] // 40)
    phase = rng.random((n_r, nh)) * 2 * np.pi
    disp = np.zeros(total)
    for i in range(nh):
        disp += amp[ridge, i] * np.sin(2 * np.pi * freq[ridge, i] * t + phase[ridge, i])
    disp *= np.sin(np.pi * t) * length[ridge]
    nrm = np.column_stack([-d[:, 1], d[:, 0]]) / length[:, None]
    xy = a[ridge] + d[ridge] * t[:, None] + nrm[ridge] * disp[:, None]
    # end points are exactly the Voronoi vertices (shared between chains)
    xy[row_index[:-1]] = a
    xy[row_index[1:] - 1] = b
    # the reference loader rejects consecutive duplicate points
    dup = np.zeros(total, bool)
    dup[1:] = (xy[1:] == xy[:-1]).all(axis=1)
    dup[row_index[:-1]] = False
    if dup.any():
        xy[dup] += 1e-9 * max(w, h)
    return PlanarGraph(xy, row_index.astype(np.uint32), left, right)


def polygon_soup(n_polys, dist="uniform", seed=1, polysize=0.001, maxseg=10,
                 affine=(50.0, 0.0, -119.0, 0.0, 30.0, 35.0)):
    """Independent small polygons like `generator.py distribution=<dist>
    geometry=polygon polysize=0.001 maxseg=10 affinematrix=50,0,-119,0,30,35`
    (reference misc/gen_polys.sh:4-22, misc/generator.py:174-199): the affine map
    moves the CENTRE, the polygon is then built around it with vertices on a circle
    of radius `polysize` at sorted uniform angles, 4..maxseg segments, closed ring."""
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        c = rng.random((n_polys, 2))
    elif dist == "gaussian":
        c = np.clip(rng.normal(0.5, 0.1, size=(n_polys, 2)), 0.0, 1.0)
    else:
        raise ValueError(dist)
    ax, bx, cx, ay, by, cy = affine
    c = np.column_stack([ax * c[:, 0] + bx * c[:, 1] + cx, ay * c[:, 0] + by * c[:, 1] + cy])
    nseg = rng.integers(4, maxseg + 1, size=n_polys)
    npts = nseg + 1  # closed ring: first point repeated
    row_index = np.zeros(n_polys + 1, np.int64)
    np.cumsum(npts, out=row_index[1:])
    total = int(row_index[-1])
    poly = np.repeat(np.arange(n_polys), npts)
    ang = rng.random(total) * (2 * np.pi)
    ang[row_index[1:] - 1] = 7.0  # closing slot sorts last; overwritten below
    order = np.lexsort((ang, poly))
    theta = ang[order]
    xy = c[poly] + polysize * np.column_stack([np.cos(theta), np.sin(theta)])
    xy[row_index[1:] - 1] = xy[row_index[:-1]]
    # the reference loader rejects consecutive duplicate points (equal random angles)
    d = np.zeros(total, bool)
    d[1:] = (xy[1:] == xy[:-1]).all(axis=1)
    d[row_index[:-1]] = False
    if d.any():
        xy[d] += 1e-9
        xy[row_index[1:] - 1] = xy[row_index[:-1]]
    ids = np.arange(1, n_polys + 1, dtype=np.int64)
    return PlanarGraph(xy, row_index.astype(np.uint32), np.zeros(n_polys, np.int64), ids)


def share_chains(base, other, frac=0.05, seed=3, stride=2):
    """Append to `other` copies of a fraction of `base`'s chains that keep every
    `stride`-th vertex: both maps then share vertices and near-collinear edges,
    the degenerate contacts real TIGER-derived layers are full of."""
    rng = np.random.default_rng(seed)
    pick = np.nonzero(rng.random(base.n_chains) < frac)[0]
    xs, rows, l, r = [other.xy], [other.row_index.astype(np.int64)], [other.left], [other.right]
    off = other.n_points
    new_rows = []
    for c in pick:
        b, e = int(base.row_index[c]), int(base.row_index[c + 1])
        idx = np.unique(np.concatenate([np.arange(b, e, stride), [e - 1]]))
        xs.append(base.xy[idx])
        off += len(idx)
        new_rows.append(off)
    row_index = np.concatenate([rows[0], np.asarray(new_rows, np.int64)])
    left = np.concatenate([l[0], base.left[pick] + 1000000])
    right = np.concatenate([r[0], base.right[pick] + 1000000])
    return PlanarGraph(np.vstack(xs), row_index.astype(np.uint32), left, right)


def write_cdb(g, path, precision=17):
    """CDB text (reference README.md:78-90; parser planar_graph.h:56-99)."""
    with open(path, "w") as f:
        for c in range(g.n_chains):
            b, e = int(g.row_index[c]), int(g.row_index[c + 1])
            f.write("%d %d %d %d %d %d\n" % (g.chain_id[c], e - b, g.first_point[c],
                                             g.last_point[c], g.left[c], g.right[c]))
            for p in g.xy[b:e]:
                f.write("%.*g %.*g\n" % (precision, p[0], precision, p[1]))


def union_bbox(*graphs):
    boxes = [g.bbox for g in graphs if g is not None and g.n_points]
    return (min(b[0] for b in boxes), min(b[1] for b in boxes),
            max(b[2] for b in boxes), max(b[3] for b in boxes))
