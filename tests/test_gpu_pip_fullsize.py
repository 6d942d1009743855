"""PIP parity at the size of BASELINE.json configs[2]: uniform random points against the
BlockGroup-scale synthetic map (~220 k faces, 28.0 M edges).

  * 10 M points, every one of them, against the oracle (orc_pip_grid: host grid filter + the
    restated update rule of src/algo/pip.h:27-96) -- closest edge ids and face ids, bit-exact;
  * 1 M points against the REFERENCE's own -mode=grid PIP run on this box
    (oracle/_ref/ref_exec, src/app/pip_grid.h:21-77), compared like the reference's -check
    does: by the scaled end points of the chosen edge (src/run_query.cu:49-98).
The full 100 M-point run with the same check over all points is tools/pip_bench.py --check -1
(profiles/pip_r2*.json)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rayjoin_b200 import synth  # noqa: E402
from rayjoin_b200.capi import PlanarGraph  # noqa: E402
from tools import ref_runner  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(rjb, oracle):
    R = synth.voronoi_map(220_000, 28_000_000, synth.US_BBOX, seed=1)
    sc = oracle.scaling_init(*synth.US_BBOX)
    pts = oracle.scale_points(sc, R.xy)
    p1, chain = oracle.build_edges(R.row_index)
    ctx = rjb.Context(device=0)
    ctx.set_option("keep_host_graph", 0)
    ctx.set_bounding_box(*synth.US_BBOX)
    ctx.set_map(0, R)
    ctx.build_index(0, "lbvh")
    yield {"R": R, "sc": sc, "pts": pts, "p1": p1, "chain": chain, "ctx": ctx}
    ctx.close()


def random_points(sc, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = int(-2**46 * 0.98), int(2**46 * 0.98)  # inside the scaled bounding box (margin 1.0 of Scaling)
    return np.column_stack([rng.integers(lo, hi, n), rng.integers(lo, hi, n)]).astype(np.int64)


@pytest.mark.parametrize("sort_queries", [1, 0])
def test_pip_10m_points_all_checked_against_oracle(rjb, oracle, big, sort_queries):
    n = 10_000_000 if sort_queries else 2_000_000
    q = random_points(big["sc"], n, 21 + sort_queries)
    ctx = big["ctx"]
    ctx.set_option("sort_queries", sort_queries)
    try:
        pip = rjb.PIP(ctx, "lbvh")
        pip.Query(1, q)
        got, faces = pip.get_closest_eids(), pip.get_face_ids()
    finally:
        ctx.set_option("sort_queries", 0)
    want = oracle.pip_grid(big["pts"], big["p1"], big["sc"], q, 1)
    assert np.array_equal(got, want)
    R = big["R"]
    assert np.array_equal(faces, oracle.face_ids(big["pts"], big["p1"], big["chain"], R.left, R.right, want))
    assert 0.5 < (got != 0xFFFFFFFF).mean() <= 1.0


@pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref/ref_exec not built")
def test_pip_1m_points_match_reference_grid(rjb, oracle, big, tmp_path):
    R = big["R"]
    n = 1_000_000
    rng = np.random.default_rng(5)
    x0, y0, x1, y1 = synth.US_BBOX
    xy = np.column_stack([rng.uniform(x0, x1, n), rng.uniform(y0, y1, n)])
    # the reference takes its query points from the vertices of map 1: 2-point chains
    S = PlanarGraph(xy, np.arange(0, n + 1, 2, dtype=np.uint32), np.ones(n // 2, np.int64),
                    np.full(n // 2, 2, np.int64), bbox=synth.US_BBOX)
    R.bbox = synth.US_BBOX  # same union box on both sides -> same Scaling
    ref = ref_runner.run_pip(None, R, S, mode="grid", warmup=0, repeat=1, grid_size=4096,
                             workdir=str(tmp_path))
    want = ref["closest_eids"]
    assert len(want) == n
    ctx = big["ctx"]
    ctx.set_map(1, S)
    ctx.set_option("sort_queries", 1)
    try:
        pip = rjb.PIP(ctx, "lbvh")
        pip.Query(1)
        got = pip.get_closest_eids()
    finally:
        ctx.set_option("sort_queries", 0)
    diff = np.nonzero(got != want)[0]
    pts, p1 = big["pts"], big["p1"]

    def endpoints(e):
        e = e.astype(np.int64)
        none = e == 0xFFFFFFFF
        p = p1[np.where(none, 0, e)].astype(np.int64)
        out = np.concatenate([pts[p], pts[p + 1]], axis=1)
        out[none] = -1
        return out
    bad = diff[(endpoints(got[diff]) != endpoints(want[diff])).any(axis=1)]
    assert len(bad) == 0, "PIP differs from reference grid at %d of %d points, e.g. %s" % (len(bad), n, bad[:5])
