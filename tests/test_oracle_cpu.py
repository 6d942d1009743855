"""CPU suite: the oracle against the committed golden vectors (generated from the
reference's own lsi.h / rational.h / scaling.h by tools/make_golden.py), against the
reference library itself when it is present, and its own internal consistency."""
import os

import numpy as np
import pytest

from helpers import DATASETS, OracleMaps, dataset

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_lsi_predicate_and_point_match_golden(oracle):
    z = np.load(os.path.join(GOLD, "lsi_kat.npz"))
    hit, x, y = oracle.intersect_batch(z["pts"])
    assert np.array_equal(hit, z["hit"])
    m = z["hit"] == 1
    assert np.array_equal(x[m], z["x"][m]) and np.array_equal(y[m], z["y"][m])
    assert 4000 < int(m.sum()) < len(m)  # the vectors exercise both outcomes


def test_argument_order_matters_in_golden(oracle):
    """intersect_test(e1, e2) != intersect_test(e2, e1) at touching contacts
    (reference src/algo/lsi.h:42-87); the oracle keeps the asymmetry."""
    z = np.load(os.path.join(GOLD, "lsi_kat.npz"))
    pts = z["pts"][:9000]  # lattice cases
    swapped = pts[:, [4, 5, 6, 7, 0, 1, 2, 3]]
    h1, _, _ = oracle.intersect_batch(pts)
    h2, _, _ = oracle.intersect_batch(swapped)
    assert (h1 != h2).sum() > 0


def test_scaling_matches_golden(oracle):
    z = np.load(os.path.join(GOLD, "scaling_kat.npz"))
    for i, b in enumerate(z["boxes"]):
        s = oracle.scaling_init(*b)
        assert [s.imin, s.imax, s.irange] == list(z["limits%d" % i])
        got = oracle.scale_points(s, z["xy%d" % i], device_semantics=False)
        assert np.array_equal(got, z["scaled_host%d" % i])
        un = oracle.unscale_points_host(s, z["ixy%d" % i])
        assert np.array_equal(un, z["unscaled_host%d" % i])
        # device semantics (fma) may differ from host rounding by at most one unit
        dev = oracle.scale_points(s, z["xy%d" % i], device_semantics=True)
        assert np.abs(dev - got).max() <= 1


def test_oracle_matches_reference_library_when_present(oracle):
    if not oracle.ref_lsi_available():
        pytest.skip("oracle/_ref/libref_lsi.so not built (no /root/reference on this box)")
    rng = np.random.default_rng(5)
    for pts in (rng.integers(-4, 5, size=(200000, 8)), rng.integers(-2**46, 2**46, size=(100000, 8))):
        h, x, y = oracle.intersect_batch(pts)
        rh, rx, ry = oracle.ref_intersect_batch(pts)
        assert np.array_equal(h, rh)
        assert np.array_equal(x[h == 1], rx[h == 1]) and np.array_equal(y[h == 1], ry[h == 1])


@pytest.mark.parametrize("name", ["lattice", "shared", "tiny", "soup"])
def test_grid_filter_equals_brute_force(oracle, name):
    R, S = dataset(name)
    om = OracleMaps(oracle, [R, S])
    for q in (0, 1):
        for a, b in zip(om.lsi(q, brute=True), om.lsi(q)):
            assert np.array_equal(a, b)
        pts = om.pts[q][:3000]
        assert np.array_equal(om.pip(q, pts, brute=True), om.pip(q, pts))


def test_edge_numbering(oracle):
    p1, ch = oracle.build_edges(np.array([0, 3, 5, 9], np.uint32))
    assert list(p1) == [0, 1, 3, 5, 6, 7]
    assert list(ch) == [0, 0, 1, 2, 2, 2]
    assert all(int(p1[e]) == e + int(ch[e]) for e in range(6))


def test_pip_rule_hand_cases(oracle):
    # one horizontal edge y=10 from x=0..10 (chain 0) and one sloped edge above it
    xy = np.array([[0, 10], [10, 10], [0, 20], [10, 30]], np.int64)
    p1 = np.array([0, 2], np.uint32)
    pts = np.array([[5, 0], [5, 10], [5, 15], [5, 40], [0, 0], [10, 0], [11, 0]], np.int64)
    e1 = oracle.pip_brute(xy, p1, pts, 1)
    e0 = oracle.pip_brute(xy, p1, pts, 0)
    NO = oracle.NO_HIT
    # below both -> lower edge; between -> upper edge; above -> none
    assert e1[0] == 0 and e1[2] == 1 and e1[3] == NO and e1[6] == NO
    # x on the left end point counts only for query_map_id == 1 (x == xmin is skipped
    # when q == 0), x on the right end point only for q == 0 (src/algo/pip.h:44-47)
    assert e1[4] == 0 and e0[4] == NO
    assert e0[5] == 0 and e1[5] == NO
    # a point ON the horizontal edge: diff == 0 -> -a/+a == 0 -> -b/+b decides:
    # q == 0 perturbs the point below the edge (hit), q == 1 above it (next edge up)
    assert e0[1] == 0 and e1[1] == 1


def test_oracle_int128_to_double_is_correctly_rounded(oracle):
    """SURVEY section 9 Q2, host half: libgcc's __floattidf (what the oracle and the reference's
    host code use) against Python's correctly rounded int -> float on the KAT the device test
    (tests/test_gpu_device_arith.py) runs through the sm_100a conversion sequence."""
    from helpers import i128_kat
    vals, words = i128_kat(n_random=100000)
    want = np.array([float(v) for v in vals])
    assert np.array_equal(oracle.i128_to_double(words).view(np.uint64), want.view(np.uint64))
