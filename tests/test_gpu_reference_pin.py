"""Parity pin against the REFERENCE ITSELF run on the GPU box: the reference's own
unmodified -mode=lbvh / -mode=grid backends (oracle/_ref/ref_exec, built from
/root/reference by oracle/Makefile) on the same inputs.

  * reference lbvh LSI uses e1 = query (map 1) edge, e2 = base (map 0) edge, the
    same argument order as this engine; its float BVH filter is documented to lose
    a few pairs, so the assertion is: every pair the reference
    reports is reported by us with bit-identical intersection coordinates, and
    any pair it misses passes the reference's OWN predicate (libref_lsi.so).
  * reference PIP (grid and lbvh) is compared on the scaled endpoints of the chosen
    edge, exactly like the reference's own -check (src/run_query.cu:49-98).
"""
import os
import sys

import numpy as np
import pytest

from helpers import OracleMaps, dataset, sort_xsects

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import ref_runner  # noqa: E402

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref/ref_exec not built")


@needs_ref
@pytest.mark.parametrize("name", ["voronoi", "shared", "aniso"])
def test_lsi_matches_reference_lbvh(rjb, oracle, name, tmp_path):
    R, S = dataset(name)
    ref = ref_runner.run_lsi(None, R, S, mode="lbvh", warmup=0, repeat=1, xsect_factor=4.0,
                             workdir=str(tmp_path), dump=True)
    a = ref["pairs"]
    assert (a[:, 3] == 1).all() and (a[:, 5] == 1).all()  # denominators are always 1
    ctx = rjb.Context([R, S])
    ctx.build_index(0, "lbvh")
    lsi = rjb.LSI(ctx, "lbvh")
    lsi.Init(4.0)
    lsi.Query(1)
    eq, eb, x, y = sort_xsects(lsi.get_xsects(), 1)
    ctx.close()
    ours = {(int(q), int(b)): (int(xx), int(yy)) for q, b, xx, yy in zip(eq, eb, x, y)}
    theirs = {(int(r[0]), int(r[1])): (int(r[2]), int(r[4])) for r in a}
    extra_ref = set(theirs) - set(ours)
    assert not extra_ref, "reference found pairs we did not: %s" % list(extra_ref)[:5]
    for k, v in theirs.items():
        assert ours[k] == v, (k, ours[k], v)
    missed_by_ref = set(ours) - set(theirs)
    # pairs the reference's float filter dropped must still satisfy its own predicate
    if missed_by_ref and oracle.ref_lsi_available():
        om = OracleMaps(oracle, [R, S])
        pts = []
        for q, b in missed_by_ref:
            pq, pb = om.p1[1][q], om.p1[0][b]
            pts.append(np.concatenate([om.pts[1][pq], om.pts[1][pq + 1], om.pts[0][pb], om.pts[0][pb + 1]]))
        hit, _, _ = oracle.ref_intersect_batch(np.asarray(pts))
        assert (hit == 1).all()
    assert len(missed_by_ref) <= max(2, len(ours) // 200)


@needs_ref
@pytest.mark.parametrize("grid_size", [64, 1000])
@pytest.mark.parametrize("name", ["voronoi", "shared", "aniso", "soup", "lattice"])
def test_lsi_grid_matches_reference_grid(rjb, oracle, name, grid_size, tmp_path):
    """-mode=grid is a drop-in for the reference's grid backend: the SAME pair set -- predicate
    evaluated as (map-0 edge, map-1 edge), pair kept only in the cell of its intersection point
    (src/app/lsi_grid.h:62-67), including the pairs that rule loses for long edges, where the
    reference's 128-bit `val - internal_min` wraps -- with the same coordinates, and equal to
    the oracle's restatement of that rule."""
    R, S = dataset(name)
    ref = ref_runner.run_lsi(None, R, S, mode="grid", warmup=0, repeat=1, xsect_factor=8.0,
                             grid_size=grid_size, workdir=str(tmp_path), dump=True)
    a = ref["pairs"]  # columns: eid(map 1), eid(map 0), x.num, x.den, y.num, y.den; sorted by (map 1, map 0)
    assert (a[:, 3] == 1).all() and (a[:, 5] == 1).all()
    ctx = rjb.Context([R, S])
    ctx.build_index(0, "grid", grid_size=grid_size)
    lsi = rjb.LSI(ctx, "grid")
    lsi.Init(8.0)
    n = lsi.Query(1)
    e1, e0, x, y = sort_xsects(lsi.get_xsects(), 1)
    ctx.close()
    om = OracleMaps(oracle, [R, S])
    want = om.lsi_refgrid(1, grid_size)
    assert n == len(want[0])
    for g, w in zip((e1, e0, x, y), want):
        assert np.array_equal(g, w)
    assert n == len(a), "reference grid reports %d pairs, this engine %d" % (len(a), n)
    assert np.array_equal(a[:, 0], e1.astype(np.int64)) and np.array_equal(a[:, 1], e0.astype(np.int64))
    assert np.array_equal(a[:, 2], x) and np.array_equal(a[:, 4], y)


@needs_ref
@pytest.mark.parametrize("mode", ["grid", "lbvh"])
@pytest.mark.parametrize("name", ["voronoi", "shared"])
def test_pip_matches_reference(rjb, oracle, name, mode, tmp_path):
    R, S = dataset(name)
    ref = ref_runner.run_pip(None, R, S, mode=mode, warmup=0, repeat=1, grid_size=256,
                             workdir=str(tmp_path))
    want = ref["closest_eids"]
    ctx = rjb.Context([R, S])
    ctx.build_index(0, "lbvh")
    pip = rjb.PIP(ctx, "lbvh")
    pip.Query(1)
    got = pip.get_closest_eids()
    ctx.close()
    assert len(got) == len(want)
    om = OracleMaps(oracle, [R, S])
    diff = np.nonzero(got != want)[0]

    def endpoints(e):
        if e == 0xFFFFFFFF:
            return None
        p = om.p1[0][e]
        return tuple(om.pts[0][p]) + tuple(om.pts[0][p + 1])
    bad = [i for i in diff if endpoints(got[i]) != endpoints(want[i])]
    assert not bad, "PIP differs from reference %s at %d points, e.g. %s" % (mode, len(bad), bad[:5])


@needs_ref
@pytest.mark.parametrize("mode,ref_mode", [("lbvh", "rjb"), ("grid", "rjbgrid")])
@pytest.mark.parametrize("name", ["voronoi", "shared"])
def test_reference_driver_calls_product_through_binding(rjb, oracle, name, mode, ref_mode, tmp_path):
    """The reference-side binding (oracle/rjb_binding.h: LSIRJB<CTX> / PIPRJB<CTX>, the classes of
    INTEGRATION.md) compiled into ref_exec: the reference's OWN driver loop -- its Context,
    PlanarGraph loader, Stream, LSI<CTX>::Query / get_xsects / CopyTo, PIP<CTX>::Query /
    get_closest_eids -- calls librjb200 through include/rjb200.h.  Its result must equal the
    ctypes path and the reference backend of the same mode."""
    R, S = dataset(name)
    out = ref_runner.run_lsi(None, R, S, mode=ref_mode, warmup=1, repeat=2, xsect_factor=4.0,
                             grid_size=256, workdir=str(tmp_path), dump=True)
    a = out["pairs"]
    ctx = rjb.Context([R, S])
    ctx.build_index(0, mode, grid_size=256)
    lsi = rjb.LSI(ctx, mode)
    lsi.Init(4.0)
    n = lsi.Query(1)
    e1, e0, x, y = sort_xsects(lsi.get_xsects(), 1)
    pip = rjb.PIP(ctx, mode)
    pip.Query(1)
    eids = pip.get_closest_eids()
    ctx.close()
    assert out["intersections"] == n == len(a)
    assert np.array_equal(a[:, 0], e1.astype(np.int64)) and np.array_equal(a[:, 1], e0.astype(np.int64))
    assert np.array_equal(a[:, 2], x) and np.array_equal(a[:, 4], y)
    assert (a[:, 3] == 1).all() and (a[:, 5] == 1).all()
    # against the reference backend of the same mode, through the same driver
    ref = ref_runner.run_lsi(None, R, S, mode=mode, warmup=0, repeat=1, xsect_factor=4.0, grid_size=256,
                             workdir=str(tmp_path), dump=True)["pairs"]
    theirs = {(int(r[0]), int(r[1])): (int(r[2]), int(r[4])) for r in ref}
    ours = {(int(r[0]), int(r[1])): (int(r[2]), int(r[4])) for r in a}
    if mode == "grid":
        assert ours == theirs
    else:  # the reference's float BVH filter may lose pairs, never find other ones
        assert set(theirs) <= set(ours) and all(ours[k] == v for k, v in theirs.items())
    p = ref_runner.run_pip(None, R, S, mode=ref_mode, warmup=0, repeat=1, grid_size=256, workdir=str(tmp_path))
    assert np.array_equal(p["closest_eids"], eids)
