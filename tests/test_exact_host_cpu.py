"""CPU suite: the PRODUCT's exact arithmetic (rayjoin_b200/csrc/rjb_exact.cuh -- the
same source the kernels compile) built for the host by tests/native/host_exact.cu and
pinned against the golden vectors of the reference's lsi.h / rational.h and against the
oracle on random crossings.  In particular the gcd-free intersection-point path
(lsi_point_axis<true>, used by k_lsi_points) must agree bit for bit with the always-gcd
path, and defer only the coordinates whose fraction lies within 1/32 of an integer."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
SRC = os.path.join(HERE, "native", "host_exact.cu")
OUT = os.path.join(HERE, "native", "libhost_exact.so")


@pytest.fixture(scope="module")
def hx():
    nvcc = os.environ.get("RJB_NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = shutil.which("nvcc")
    deps = [SRC, os.path.join(ROOT, "rayjoin_b200", "csrc", "rjb_exact.cuh"),
            os.path.join(ROOT, "rayjoin_b200", "csrc", "rjb_common.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        assert nvcc, "nvcc is needed to build the host view of rjb_exact.cuh"
        subprocess.run([nvcc, "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets",
                        "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "rayjoin_b200", "csrc"), "-Xcompiler", "-fPIC",
                        "-shared", "-o", OUT, SRC], check=True)
    return C.CDLL(OUT)


def run(hx, pts, mode):
    pts = np.ascontiguousarray(pts, np.int64)
    n = len(pts)
    hit = np.zeros(n, np.uint8)
    ox = np.zeros(n, np.int64)
    oy = np.zeros(n, np.int64)
    nd = C.c_uint64(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    hx.hx_intersect_batch(p(pts), C.c_uint64(n), C.c_int(mode), p(hit), p(ox), p(oy), C.byref(nd))
    return hit, ox, oy, nd.value


def crossing(rng, n, span_bits, off_bits=46):
    """pairs of edges of span ~2^span_bits around a common random centre: most cross"""
    c = rng.integers(-2**off_bits + 2**(span_bits + 1), 2**off_bits - 2**(span_bits + 1), size=(n, 1, 2))
    d = rng.integers(-2**span_bits, 2**span_bits, size=(n, 2, 2))
    jb = max(1, span_bits - 3)
    j = rng.integers(-2**jb, 2**jb, size=(n, 2, 2))
    e = np.concatenate([c + j[:, :1] - d[:, :1], c + j[:, :1] + d[:, :1],
                        c + j[:, 1:] - d[:, 1:], c + j[:, 1:] + d[:, 1:]], axis=1)
    return e.reshape(n, 8)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_product_arithmetic_matches_reference_golden(hx, mode):
    z = np.load(os.path.join(GOLD, "lsi_kat.npz"))
    hit, x, y, _ = run(hx, z["pts"], mode)
    assert np.array_equal(hit, z["hit"])
    m = z["hit"] == 1
    assert np.array_equal(x[m], z["x"][m]) and np.array_equal(y[m], z["y"][m])


@pytest.mark.parametrize("span_bits", [3, 8, 16, 24, 30, 37, 38, 40])
def test_gcd_free_point_path_is_bit_exact(hx, oracle, span_bits):
    rng = np.random.default_rng(1000 + span_bits)
    pts = crossing(rng, 60000, span_bits)
    hit, x, y = oracle.intersect_batch(pts)
    m = hit == 1
    assert m.sum() > 30000
    for mode in (0, 1, 2):
        h2, x2, y2, nd = run(hx, pts, mode)
        assert np.array_equal(hit, h2)
        assert np.array_equal(x[m], x2[m]) and np.array_equal(y[m], y2[m])
        if mode == 1 and 8 <= span_bits < 38:
            # 2 * 1/32 of the fractions, nothing else, take the gcd path
            assert 0.05 < nd / (2.0 * m.sum()) < 0.075
        if mode == 1 and span_bits >= 38 + 1:
            assert nd > 1.9 * m.sum()  # long edges: always the general path


def test_gcd_free_point_path_special_cases(hx, oracle):
    rng = np.random.default_rng(5)
    sets = [crossing(rng, 50000, 4, off_bits=7), crossing(rng, 50000, 10, off_bits=13)]  # around 0: negative x
    p = crossing(rng, 50000, 10)
    p[:, 1] = p[:, 3]  # horizontal e1: y is an exact integer (rs == 0)
    sets.append(p)
    p = crossing(rng, 50000, 10)
    p[:, 4] = p[:, 6]  # vertical e2
    sets.append(p)
    p = crossing(rng, 50000, 12)
    p[:, 4:6] = p[:, 0:2]  # shared end point: the clamp decides
    sets.append(p)
    for pts in sets:
        hit, x, y = oracle.intersect_batch(pts)
        m = hit == 1
        for mode in (0, 1, 2):
            h2, x2, y2, _ = run(hx, pts, mode)
            assert np.array_equal(hit, h2)
            assert np.array_equal(x[m], x2[m]) and np.array_equal(y[m], y2[m])


def test_edge_descriptor_covers_the_edge_box(hx):
    """The LSI filter decides from a 4-byte descriptor per edge: min-corner cell + class.
    Same / Small must describe a box inside the 2 x 2 cells at the min corner (what the
    dilated bitmap covers); the cell code must agree with the cell of the quantised box
    coordinate the index is built from (occ_cell(quant(v)))."""
    rng = np.random.default_rng(3)
    n = 200000
    p1 = rng.integers(-2**46, 2**46, size=(n, 2))
    step = rng.integers(-2**36, 2**36, size=(n, 2)) >> rng.integers(0, 30, size=(n, 1))
    p2 = np.clip(p1 + step, -2**46, 2**46 - 1)
    hx.hx_occ_code.restype = C.c_uint32
    hx.hx_edge_desc.restype = C.c_uint32
    seen = set()
    for (x1, y1), (x2, y2) in zip(p1[:20000].tolist(), p2[:20000].tolist()):
        c1 = hx.hx_occ_code(C.c_longlong(x1), C.c_longlong(y1))
        c2 = hx.hx_occ_code(C.c_longlong(x2), C.c_longlong(y2))
        assert (c1 & 4095) == hx.hx_occ_cell_of_quant(C.c_longlong(x1))
        assert (c1 >> 12) == hx.hx_occ_cell_of_quant(C.c_longlong(y1))
        d = hx.hx_edge_desc(C.c_longlong(x1), C.c_longlong(y1), C.c_longlong(x2), C.c_longlong(y2))
        cls, corner = d >> 24, d & 0xFFFFFF
        cx0, cx1 = sorted((c1 & 4095, c2 & 4095))
        cy0, cy1 = sorted((c1 >> 12, c2 >> 12))
        assert corner == (cy0 << 12 | cx0)
        ex, ey = cx1 - cx0, cy1 - cy0
        want = 0 if (ex, ey) == (0, 0) else 1 if ex <= 1 and ey <= 1 else 2
        assert cls == want
        seen.add(cls)
    assert seen == {0, 1, 2}


def test_tile_descriptor_classes(hx):
    """Tile descriptor of the two-level LSI filter: class k promises that the box lies within the
    2^(k-1) x 2^(k-1) cells at its min corner (the area the k-th dilated bitmap covers)."""
    hx.hx_tile_desc.restype = C.c_uint32
    rng = np.random.default_rng(5)
    for _ in range(5000):
        x0, y0 = (int(v) for v in rng.integers(0, 4000, 2))
        ex, ey = (int(v) for v in rng.integers(0, 12, 2))
        d = hx.hx_tile_desc(x0, y0, x0 + ex, y0 + ey)
        cls, corner = d >> 24, d & 0xFFFFFF
        assert corner == (y0 << 12 | x0)
        e = max(ex, ey)
        assert cls == (1 if e == 0 else 2 if e <= 1 else 3 if e <= 3 else 4 if e <= 7 else 5)


def test_two_level_filter_never_drops_a_kept_edge(hx):
    """The invariant behind k_lsi_filter_tiles: if the edge-level test keeps an edge (its descriptor's
    bit is set in occ / occ2), the tile-level test keeps the tile that holds it -- the tile descriptor
    covers the cell boxes of its 8 edges, and bitmap K is the K x K dilation.  Random chains of short
    and long edges against random sparse occupancy bitmaps, everything restated with numpy."""
    hx.hx_occ_code.restype = C.c_uint32
    hx.hx_edge_desc.restype = C.c_uint32
    hx.hx_tile_desc.restype = C.c_uint32
    rng = np.random.default_rng(21)
    G = 4096
    occ = np.zeros((G, G), bool)  # [y, x]
    pts0 = rng.integers(0, G, size=(3000, 2))
    occ[pts0[:, 1], pts0[:, 0]] = True

    def dilate(a, k):  # bit (x, y) = OR over [x, x + k) x [y, y + k)
        out = np.zeros_like(a)
        for dy in range(k):
            for dx in range(k):
                out[:G - dy, :G - dx] |= a[dy:, dx:]
        return out
    maps = {1: occ, 2: dilate(occ, 2), 3: dilate(occ, 4), 4: dilate(occ, 8)}
    cell = 1 << 35  # coordinate units per occupancy cell
    kept_edges = live_tiles = 0
    for trial in range(60):
        # a chain of 9 points = 8 edges = one tile; steps from a fraction of a cell to several cells
        scale = int(rng.choice([cell // 8, cell // 2, cell, 3 * cell]))
        p = np.cumsum(np.vstack([rng.integers(-2**46 + 40 * cell, 2**46 - 40 * cell, size=(1, 2)),
                                 rng.integers(-scale, scale + 1, size=(8, 2))]), axis=0)
        codes = [hx.hx_occ_code(C.c_longlong(int(x)), C.c_longlong(int(y))) for x, y in p]
        cx = np.array([c & 4095 for c in codes]); cy = np.array([c >> 12 for c in codes])
        # plant an occupied cell near the chain in half of the trials, so that edges are kept
        if trial % 2 == 0:
            j = int(rng.integers(0, 9))
            occ2 = occ.copy(); occ2[cy[j], cx[j]] = True
            m = {1: occ2, 2: dilate(occ2, 2), 3: dilate(occ2, 4), 4: dilate(occ2, 8)}
        else:
            m = maps
        keep_any = False
        for e in range(8):
            d = hx.hx_edge_desc(*(C.c_longlong(int(v)) for v in (p[e, 0], p[e, 1], p[e + 1, 0], p[e + 1, 1])))
            cls, code = d >> 24, d & 0xFFFFFF
            ex0, ey0 = code & 4095, code >> 12
            if cls == 0:
                keep = m[1][ey0, ex0]
            elif cls == 1:
                keep = m[2][ey0, ex0]
            else:  # longer edge: rectangle test (occ_rect), exact over its cell box
                x0, x1 = sorted((cx[e], cx[e + 1])); y0, y1 = sorted((cy[e], cy[e + 1]))
                keep = m[1][y0:y1 + 1, x0:x1 + 1].any()
            keep_any |= bool(keep)
            kept_edges += bool(keep)
        t = hx.hx_tile_desc(int(cx.min()), int(cy.min()), int(cx.max()), int(cy.max()))
        tcls, tcode = t >> 24, t & 0xFFFFFF
        live = True if tcls == 5 else bool(m[tcls][tcode >> 12, tcode & 4095])
        live_tiles += live
        assert tcls in (1, 2, 3, 4, 5)
        if keep_any:
            assert live
    assert kept_edges > 20 and live_tiles < 60  # both outcomes occurred


def test_cell_directory_items_report_each_overlap_exactly_once(hx):
    """k_lsi_cells decides a (query, leaf) pair from the leaf's 16-byte item record (box clipped to
    the cell) and the query box clipped to the same cell.  Over all common cells of the two cell
    boxes it must fire exactly once when the quantised boxes overlap -- in the cell holding the
    min corner of the intersection of the cell boxes -- and never otherwise; the record must
    give back the leaf's first point and edge count."""
    rng = np.random.default_rng(11)
    n = 300000
    cell = 1 << 19
    # leaf boxes of 0..3 cells, query boxes of 0..2 cells, near each other; some exactly touching
    l0 = rng.integers(-2**30, 2**30 - 4 * cell, size=(n, 2))
    lsz = (rng.integers(0, 3 * cell, size=(n, 2)) >> rng.integers(0, 19, size=(n, 1)))
    q0 = l0 + rng.integers(-2 * cell, 3 * cell, size=(n, 2)) // rng.choice([1, 1, 7, 4096], size=(n, 1))
    qsz = (rng.integers(0, 2 * cell, size=(n, 2)) >> rng.integers(0, 19, size=(n, 1)))
    touch = rng.random(n) < 0.1
    q0[touch, 0] = (l0 + lsz)[touch, 0]  # query starts where the leaf ends (closed boxes: overlap)
    q0 = np.clip(q0, -2**30, 2**30 - 3 * cell)
    leaf = np.ascontiguousarray(np.concatenate([l0, l0 + lsz], axis=1).astype(np.int32))
    query = np.ascontiguousarray(np.concatenate([q0, q0 + qsz], axis=1).astype(np.int32))
    hits, cellid, first, count = (np.zeros(n, np.uint32) for _ in range(4))
    u32p = C.POINTER(C.c_uint32)
    hx.hx_cell_item_batch(leaf.ctypes.data_as(C.POINTER(C.c_int)), query.ctypes.data_as(C.POINTER(C.c_int)),
                          C.c_uint64(n), hits.ctypes.data_as(u32p), cellid.ctypes.data_as(u32p),
                          first.ctypes.data_as(u32p), count.ctypes.data_as(u32p))
    L, Q = leaf.astype(np.int64), query.astype(np.int64)
    overlap = (L[:, 0] <= Q[:, 2]) & (Q[:, 0] <= L[:, 2]) & (L[:, 1] <= Q[:, 3]) & (Q[:, 1] <= L[:, 3])
    assert 0.1 < overlap.mean() < 0.9
    assert np.array_equal(hits, overlap.astype(np.uint32))
    occ = lambda v: ((v + 2**30) >> 19) & 4095
    want_cell = np.maximum(occ(L[:, 1]), occ(Q[:, 1])) * 4096 + np.maximum(occ(L[:, 0]), occ(Q[:, 0]))
    assert np.array_equal(cellid[overlap], want_cell[overlap].astype(np.uint32))
    idx = np.arange(n)
    assert np.array_equal(first[overlap], (1000 + idx[overlap]).astype(np.uint32))
    assert np.array_equal(count[overlap], (1 + idx[overlap] % 8).astype(np.uint32))


@pytest.mark.parametrize("q", [0, 1])
def test_pip_update_rule_matches_oracle(hx, oracle, q):
    """The kernels' pip_update (closest edge above, src/algo/pip.h:27-96) scanned in eid order
    against the oracle's brute force: lattice maps (points with x equal to a vertex x,
    horizontal / vertical / coincident edges) and a Voronoi map with random points."""
    import sys
    sys.path.insert(0, HERE)
    from helpers import dataset, OracleMaps
    for name in ("lattice", "voronoi", "shared"):
        R, S = dataset(name)
        om = OracleMaps(oracle, [R, S])
        b = 1 - q
        xyb, p1 = om.pts[b], om.p1[b]
        edges = np.concatenate([xyb[p1], xyb[p1 + 1]], axis=1).astype(np.int64)
        rng = np.random.default_rng(17)
        pts = om.pts[q][:1500]
        lo, hi = xyb.min(0), xyb.max(0)
        rnd = np.column_stack([rng.integers(lo[0], hi[0] + 1, 800), rng.integers(lo[1], hi[1] + 1, 800)])
        pts = np.ascontiguousarray(np.concatenate([pts, rnd]), np.int64)
        if len(edges) > 4000:  # keep the brute-force scan short
            edges, p1 = edges[:4000], p1[:4000]
        want = oracle.pip_brute(xyb, p1, pts, q)
        got = np.zeros(len(pts), np.uint32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        edges = np.ascontiguousarray(edges)
        hx.hx_pip_batch(p(edges), C.c_uint64(len(edges)), p(pts), C.c_uint64(len(pts)), C.c_int(q), p(got))
        assert np.array_equal(got, want), name
        assert (got != 0xFFFFFFFF).any()


@pytest.mark.parametrize("gsize", [64, 2048, 15000, 32768])
def test_xsect_cell_shortcut_matches_reference_sequence(hx, oracle, gsize):
    """Grid LSI keeps a pair only in the cell of its intersection point (src/app/lsi_grid.h:62-67,
    src/grid/cell.h:15-22).  The kernel decides the cell without the gcd unless the point lies
    within 1e-6 of a cell boundary; the oracle replays the reference's rational -> double
    sequence literally (including the 128-bit wrap-around for long edges)."""
    sc = oracle.scaling_init(-179.15, -14.55, 179.78, 71.39)
    cs = float(gsize) / float(sc.irange) * 0.999
    rng = np.random.default_rng(gsize)
    sets = [crossing(rng, 60000, b) for b in (4, 12, 20, 28, 33, 36, 37, 38, 41, 44)]
    # points ON cell boundaries: lattice of multiples of the cell width (integer crossings)
    w = int(1.0 / cs) + 1
    base = rng.integers(-2**45 // w, 2**45 // w, size=(40000, 1, 2)) * w
    d = rng.integers(1, 2**10, size=(40000, 2, 2))
    e = np.concatenate([base - d[:, :1] * [[1, 0]], base + d[:, :1] * [[1, 0]],
                        base - d[:, 1:] * [[0, 1]], base + d[:, 1:] * [[0, 1]]], axis=1).reshape(-1, 8)
    sets.append(e)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for pts in sets:
        pts = np.ascontiguousarray(pts, np.int64)
        hit, cx, cy, owned = oracle.refgrid_cells(pts, sc, gsize)
        gx, gy = np.zeros(len(pts), np.int32), np.zeros(len(pts), np.int32)
        hx.hx_xsect_ref_cells(p(pts), C.c_uint64(len(pts)), C.c_longlong(sc.imin), C.c_double(cs), p(gx), p(gy))
        m = hit == 1
        assert m.sum() > 1000
        assert np.array_equal(gx[m], cx[m]) and np.array_equal(gy[m], cy[m])
