"""Parity of the CUDA path (through the C ABI) against the CPU oracle.
Bit-exact: pair sets, intersection coordinates, closest edges, face ids."""
import numpy as np
import pytest

from helpers import DATASETS, OracleMaps, dataset, sort_xsects
from rayjoin_b200 import synth


def synth_soup(n, seed):
    return synth.polygon_soup(n, "gaussian", seed=seed, polysize=0.2)

pytestmark = pytest.mark.gpu

MODES = ["brute", "lbvh", "grid"]


@pytest.fixture(scope="module")
def loaded(rjb, oracle):
    cache = {}

    def get(name):
        if name not in cache:
            R, S = dataset(name)
            ctx = rjb.Context([R, S])
            cache[name] = (ctx, OracleMaps(oracle, [R, S]))
        return cache[name]
    yield get
    for ctx, _ in cache.values():
        ctx.close()


@pytest.mark.parametrize("name", DATASETS)
def test_scaling_matches_oracle(loaded, name):
    ctx, om = loaded(name)
    s = ctx.get_scaling()
    for f in ("rx", "ry", "rrx", "rry", "deltax", "deltay", "ddeltax", "ddeltay"):
        assert getattr(s, f) == getattr(om.sc, f), f
    for im in range(2):
        assert np.array_equal(ctx.map_points(im), om.pts[im])


@pytest.mark.parametrize("q", [1, 0])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", DATASETS)
def test_lsi_pair_set_and_points(rjb, loaded, name, mode, q):
    ctx, om = loaded(name)
    ctx.build_index(1 - q, mode, grid_size=64)
    lsi = rjb.LSI(ctx, mode)
    lsi.Init(4.0)
    n = lsi.Query(q)
    got = sort_xsects(lsi.get_xsects(), q)
    # grid mode has the reference's GRID semantics (src/app/lsi_grid.h:62-67): predicate
    # evaluated as (map-0 edge, map-1 edge) whatever the query side, pair kept iff the cell of
    # its intersection point holds both edges; lbvh / brute: e1 = query edge, no cell rule
    want = om.lsi_refgrid(q, 64) if mode == "grid" else om.lsi(q)
    assert n == len(want[0])
    for g, w, what in zip(got, want, ("eid_query", "eid_base", "x", "y")):
        assert np.array_equal(g, w), what
    assert lsi.n_candidates >= n


@pytest.mark.parametrize("leaf", [1, 2, 8])
def test_lsi_lbvh_leaf_sizes_and_sorted_queries(rjb, loaded, leaf):
    ctx, om = loaded("shared")
    ctx.set_option("lbvh_leaf_size", leaf)
    ctx.set_option("sort_queries", leaf % 2)
    try:
        ctx.build_index(0, "lbvh")
        lsi = rjb.LSI(ctx, "lbvh")
        lsi.Init(4.0)
        lsi.Query(1)
        got = sort_xsects(lsi.get_xsects(), 1)
        for g, w in zip(got, om.lsi(1)):
            assert np.array_equal(g, w)
    finally:
        ctx.set_option("lbvh_leaf_size", 4)
        ctx.set_option("sort_queries", 0)


@pytest.mark.parametrize("enlarge,iters", [(5.0, 5), (3.5, 2), (1.5, 1), (1000.0, 3)])
@pytest.mark.parametrize("name", ["voronoi", "shared", "lattice", "soup"])
def test_adaptive_leaf_grouping(rjb, loaded, name, enlarge, iters):
    """Options lbvh_ag / lbvh_ag_iter / lbvh_enlarge_x1000 (RayJoin's -ag -ag_iter -enlarge,
    src/rt/primitive.h:120-260): leaves are merged runs of consecutive chain edges.  The
    grouping changes the index, never the result; a huge limit merges everything a block holds,
    a limit near 1 merges next to nothing."""
    ctx, om = loaded(name)
    try:
        ctx.set_option("lbvh_leaf_size", 1)
        ctx.build_index(0, "lbvh")
        n_edges_as_leaves = ctx.index_info(0, "lbvh")["units"]
        ctx.set_option("lbvh_ag", 1)
        ctx.set_option("lbvh_ag_iter", iters)
        ctx.set_option("lbvh_enlarge_x1000", int(enlarge * 1000))
        for q in (1, 0):
            ctx.build_index(1 - q, "lbvh")
            lsi = rjb.LSI(ctx, "lbvh")
            lsi.Init(4.0)
            n = lsi.Query(q)
            want = om.lsi(q)
            assert n == len(want[0])
            for g, w in zip(sort_xsects(lsi.get_xsects(), q), want):
                assert np.array_equal(g, w)
            pip = rjb.PIP(ctx, "lbvh")
            pip.Query(q)
            assert np.array_equal(pip.get_closest_eids(), om.pip(q, om.pts[q]))
        ctx.build_index(0, "lbvh")
        leaves = ctx.index_info(0, "lbvh")["units"]
        assert leaves <= n_edges_as_leaves
        if enlarge >= 1000:  # everything inside a block of 2^iters edges merges
            assert leaves <= n_edges_as_leaves // 2 + ctx.map_info(0)["chains"] * 4
        if name == "voronoi" and enlarge == 5.0:
            assert leaves < 0.6 * n_edges_as_leaves
    finally:
        ctx.set_option("lbvh_ag", 0)
        ctx.set_option("lbvh_leaf_size", 4)


@pytest.mark.parametrize("q", [1, 0])
@pytest.mark.parametrize("name", DATASETS + ["dense"])
def test_lsi_cell_directory_path(rjb, loaded, name, q):
    """Option lsi_cells: the survivors of the occupancy filter find their candidate leaves
    through the directory of occupied cells (k_lsi_cells), long edges through the tree
    walk.  Same result, same number of exact-predicate evaluations as the tree walk."""
    ctx, om = loaded(name)
    want = om.lsi(q)
    counts = {}
    try:
        # fused: 1 = exact + point pass as one kernel, 2 = its warp-private variant, 0 = two kernels
        for cells, tiles, fused in ((0, 1, 1), (1, 1, 1), (0, 0, 1), (1, 0, 1), (0, 1, 0), (1, 1, 0), (1, 1, 2), (0, 0, 2)):
            ctx.set_option("lsi_cells", 2 * cells)  # (2: keep the directory whatever the query finds)
            ctx.set_option("lsi_tile_filter", tiles)
            ctx.set_option("lsi_fused", min(fused, 1))
            ctx.set_option("lsi_resolve_warp", 1 if fused == 2 else 0)
            ctx.set_option("stage_timing", -1 if fused == 2 else 1)  # (also the event-free query)
            ctx.set_option("lsi_filter", 1)
            ctx.set_option("sort_queries", 0)
            ctx.build_index(1 - q, "lbvh")
            lsi = rjb.LSI(ctx, "lbvh")
            lsi.Init(4.0)
            for _ in range(2):  # the second query runs with launch sizes learnt from the first
                n = lsi.Query(q)
                got = sort_xsects(lsi.get_xsects(), q)
                assert n == len(want[0])
                for g, w in zip(got, want):
                    assert np.array_equal(g, w)
            counts[(cells, tiles, fused)] = lsi.n_candidates
            surv = counts.setdefault(("survivors", cells), ctx.last_stats()[7])
            # (the tile test is at least as tight as the edge test: a Small descriptor looks at the
            # whole 2 x 2 block, the tile box only at the cells its edges touch)
            assert ctx.last_stats()[7] <= surv if tiles else ctx.last_stats()[7] >= surv
            st = ctx.last_stats()
            # the directory is only built where leaves are small against the cells
            assert st[5] in (0, cells)
            if name == "dense":
                assert st[7] > 0 and st[5] == cells  # filter on, and the path that was asked for
    finally:
        ctx.set_option("lsi_cells", 0)
        ctx.set_option("lsi_tile_filter", 1)
        ctx.set_option("lsi_fused", 1)
        ctx.set_option("lsi_resolve_warp", 0)
        ctx.set_option("stage_timing", 1)
        ctx.set_option("lsi_filter", -1)
    assert len({counts[k] for k in counts if k[0] != "survivors"}) == 1


@pytest.mark.parametrize("mode,flt", [("lbvh", 1), ("lbvh", 2), ("lbvh", 3), ("lbvh", 0), ("grid", 0)])
@pytest.mark.parametrize("name", ["voronoi", "shared", "soup"])
def test_lsi_query_window(rjb, loaded, name, mode, flt):
    """Options lsi_window_begin / lsi_window_end: the query edges starting in a window of the
    query map's points (what one rank of the multi-GPU overlay runs).  Windows that tile the map
    -- cut at arbitrary points, not multiples of 32 -- add up to the whole result, ids global."""
    ctx, om = loaded(name)
    q = 1
    n_pts = ctx.map_info(q)["points"]
    cuts = [0, n_pts // 3 + 5, (2 * n_pts) // 3 - 7, n_pts]
    want = om.lsi_refgrid(q, 64) if mode == "grid" else om.lsi(q)
    parts = []
    try:
        # flt 1: two-level filter (tiles, then edges); 2: one-level filter; 3: two-level + cell directory
        ctx.set_option("lsi_filter", min(flt, 1))
        ctx.set_option("lsi_tile_filter", 0 if flt == 2 else 1)
        ctx.set_option("lsi_cells", 2 if flt == 3 else 0)
        ctx.set_option("sort_queries", 0)
        ctx.build_index(1 - q, mode, grid_size=64)
        lsi = rjb.LSI(ctx, mode)
        lsi.Init(4.0)
        for a, b in zip(cuts[:-1], cuts[1:]):
            ctx.set_option("lsi_window_begin", a)
            ctx.set_option("lsi_window_end", b)
            lsi.Query(q)
            xs = lsi.get_xsects()
            p1 = om.p1[q][xs["eid"][:, q]]  # start point of every reported query edge
            assert ((p1 >= a) & (p1 < b)).all()
            parts.append(xs)
    finally:
        ctx.set_option("lsi_window_begin", 0)
        ctx.set_option("lsi_window_end", 0)
        ctx.set_option("lsi_filter", -1)
        ctx.set_option("lsi_tile_filter", 1)
        ctx.set_option("lsi_cells", 0)
    got = sort_xsects(np.concatenate(parts), q)
    assert len(got[0]) == len(want[0])
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_cached_query_order_follows_the_map(rjb, oracle):
    """The Morton order of a short-chain query map is computed once and kept with the map: a
    second query reuses it, replacing the map drops it."""
    R = synth_soup(3000, 1)
    S1, S2 = synth_soup(2500, 2), synth_soup(2600, 3)
    ctx = rjb.Context(device=0)
    try:
        bbox = synth.union_bbox(R, S1, S2)
        ctx.set_bounding_box(*bbox)
        ctx.set_option("sort_queries", 1)
        ctx.set_map(0, R)
        ctx.build_index(0, "lbvh")
        lsi = rjb.LSI(ctx, "lbvh")
        lsi.Init(4.0)
        for S in (S1, S2, S1):
            ctx.set_map(1, S)
            sc = oracle.scaling_init(*bbox)
            pts = [oracle.scale_points(sc, g.xy) for g in (R, S)]
            p1 = [oracle.build_edges(g.row_index)[0] for g in (R, S)]
            want = oracle.lsi_grid(pts[1], p1[1], pts[0], p1[0], sc)
            for _ in range(2):
                assert lsi.Query(1) == len(want[0])
                for g, w in zip(sort_xsects(lsi.get_xsects(), 1), want):
                    assert np.array_equal(g, w)
    finally:
        ctx.close()


def test_chunked_upload_matches_single_chunk(rjb, oracle):
    """rjb_set_map uploads in chunks and runs the load kernel per chunk; the edge descriptors,
    chain maps and last-point bits must not depend on where the chunk boundaries fall."""
    R, S = dataset("voronoi")
    om = OracleMaps(oracle, [R, S])
    want = om.lsi(1)
    for chunk in (1024, 3072, 1 << 20):
        ctx = rjb.Context(device=0)
        try:
            ctx.set_option("load_chunk_points", chunk)
            ctx.set_bounding_box(*om.bbox)
            ctx.set_map(0, R)
            ctx.set_map(1, S)
            for im in range(2):
                assert np.array_equal(ctx.map_points(im), om.pts[im])
            ctx.set_option("lsi_filter", 1)  # the filter reads the per-edge descriptors
            ctx.build_index(0, "lbvh")
            lsi = rjb.LSI(ctx, "lbvh")
            lsi.Init(4.0)
            assert lsi.Query(1) == len(want[0])
            for g, w in zip(sort_xsects(lsi.get_xsects(), 1), want):
                assert np.array_equal(g, w)
            ctx.build_index(0, "grid", grid_size=64)  # the grid path reads last_bits / point_chain
            lsi = rjb.LSI(ctx, "grid")
            lsi.Init(4.0)
            assert lsi.Query(1) == len(om.lsi_refgrid(1, 64)[0])
            pip = rjb.PIP(ctx, "lbvh")
            pip.Query(1)
            assert np.array_equal(pip.get_closest_eids(), om.pip(1, om.pts[1]))
        finally:
            ctx.close()


def test_lsi_queue_overflow_is_detected(rjb, loaded):
    ctx, om = loaded("voronoi")
    ctx.build_index(0, "lbvh")
    want = len(om.lsi(1)[0])
    lsi = rjb.LSI(ctx, "lbvh")
    lsi.Init(1e-4)  # far too small
    with pytest.raises(rjb.RjbError) as ei:
        lsi.Query(1)
    assert ei.value.code == 3
    assert ei.value.needed == want


@pytest.mark.parametrize("q", [1, 0])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", DATASETS)
def test_pip_vertices_of_other_map(rjb, loaded, name, mode, q):
    """PIP of every vertex of map q in map 1-q (what overlay and query_exec -poly2 do)."""
    ctx, om = loaded(name)
    ctx.build_index(1 - q, mode, grid_size=64)
    pip = rjb.PIP(ctx, mode)
    pip.Query(q)
    eids = pip.get_closest_eids()
    want = om.pip(q, om.pts[q])
    assert np.array_equal(eids, want)
    assert np.array_equal(pip.get_face_ids(), om.faces(q, want))


@pytest.mark.parametrize("mode", ["lbvh", "grid"])
def test_pip_random_points_sorted(rjb, loaded, mode):
    ctx, om = loaded("voronoi")
    rng = np.random.default_rng(7)
    s = om.sc
    pts = rng.integers(s.imin // 2, s.imax // 2, size=(20000, 2))
    ctx.set_option("sort_queries", 1)
    try:
        ctx.build_index(0, mode, grid_size=128)
        pip = rjb.PIP(ctx, mode)
        pip.Query(1, pts)
        assert np.array_equal(pip.get_closest_eids(), om.pip(1, pts))
    finally:
        ctx.set_option("sort_queries", 0)


def test_oracle_brute_equals_grid_filter(oracle):
    """the oracle's own filter loses nothing (CPU only, but cheap enough to keep here)"""
    for name in ("lattice", "shared"):
        R, S = dataset(name)
        om = OracleMaps(oracle, [R, S])
        for q in (0, 1):
            for a, b in zip(om.lsi(q, brute=True), om.lsi(q)):
                assert np.array_equal(a, b)
            assert np.array_equal(om.pip(q, om.pts[q], brute=True), om.pip(q, om.pts[q]))


def test_empty_query_map(rjb, oracle):
    R, _ = dataset("tiny")
    from rayjoin_b200.capi import PlanarGraph
    E = PlanarGraph(np.zeros((0, 2)), np.zeros(0, np.uint32), [], [], bbox=R.bbox)
    ctx = rjb.Context([R, E], bbox=R.bbox)
    ctx.build_index(0, "lbvh")
    lsi = rjb.LSI(ctx, "lbvh")
    lsi.Init(1.0)
    assert lsi.Query(1) == 0
    ctx.close()


def test_query_without_index_fails(rjb):
    R, S = dataset("tiny")
    ctx = rjb.Context([R, S])
    lsi = rjb.LSI(ctx, "lbvh")
    with pytest.raises(rjb.RjbError) as ei:
        lsi.Query(1)
    assert ei.value.code == 5
    ctx.close()
