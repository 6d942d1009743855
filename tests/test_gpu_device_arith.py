"""The exact arithmetic of the query kernels on the DEVICE (sm_100a compile of
rayjoin_b200/csrc/rjb_exact.cuh) through the debug entries of the C ABI:

  * the golden vectors of the reference's own src/algo/lsi.h + src/util/rational.h
    (tests/golden/lsi_kat.npz, made by tools/make_golden.py from /root/reference), incl.
    the int128 wrap-around regime and |coord| ~ 2^46, through lsi_intersect and both paths of
    lsi_point_axis (always-gcd and the deferring path of k_lsi_points);
  * random crossings of every span from 2^3 to 2^45 (the `else` branch of lsi_point_axis
    and rat_make only run for spans >= 2^38) against the oracle;
  * (double)(__int128), the double division and the truncating store against Python's
    correctly rounded int -> float (device PTX sequence vs libgcc
    __floattidf): |v| just above 2^53 / 2^64 / 2^96, exact ties, ties +- 1;
  * the PIP update rule (src/algo/pip.h:27-96) on lattice maps against the oracle.
"""
import os

import numpy as np
import pytest

from helpers import OracleMaps, dataset, i128_kat
from test_exact_host_cpu import crossing

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx(rjb):
    c = rjb.Context(device=0)
    yield c
    c.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_reference_golden_vectors_on_device(ctx, mode):
    z = np.load(os.path.join(GOLD, "lsi_kat.npz"))
    flags, x, y = ctx.debug_intersect_batch(z["pts"], mode)
    assert np.array_equal(flags & 1, z["hit"])
    m = z["hit"] == 1
    assert m.sum() > 3000
    assert np.array_equal(x[m], z["x"][m]) and np.array_equal(y[m], z["y"][m])
    if mode == 1:
        assert ((flags >> 1) & 3).any()  # some coordinates did go through the deferred gcd pass


@pytest.mark.parametrize("span_bits", [3, 8, 16, 24, 30, 37, 38, 40, 43, 44])
def test_device_points_match_oracle_for_every_span(ctx, oracle, span_bits):
    rng = np.random.default_rng(2000 + span_bits)
    pts = crossing(rng, 200000, span_bits, off_bits=46)
    hit, x, y = oracle.intersect_batch(pts)
    m = hit == 1
    assert m.sum() > 80000
    for mode in (0, 1, 2):
        flags, gx, gy = ctx.debug_intersect_batch(pts, mode)
        assert np.array_equal(flags & 1, hit)
        assert np.array_equal(gx[m], x[m]) and np.array_equal(gy[m], y[m])
        if mode == 1 and span_bits > 38:
            assert (((flags[m] >> 1) & 3) == 3).mean() > 0.9  # long edges: the general path


def test_device_degenerate_cases_match_oracle(ctx, oracle):
    rng = np.random.default_rng(11)
    sets = [rng.integers(-1, 2, size=(200000, 8)), rng.integers(-3, 4, size=(300000, 8)),
            rng.integers(-2**46, 2**46, size=(300000, 8))]
    p = rng.integers(-2**45, 2**45, size=(200000, 8))
    p[:50000, 4:6] = p[:50000, 0:2]
    p[50000:100000, 6:8] = p[50000:100000, 2:4]
    p[100000:150000, 4:6] = (p[100000:150000, 0:2] + p[100000:150000, 2:4]) // 2
    p[150000:, 4:8] = p[150000:, [2, 3, 0, 1]]
    sets.append(p)
    for pts in sets:
        hit, x, y = oracle.intersect_batch(pts)
        m = hit == 1
        for mode in (0, 1, 2):
            flags, gx, gy = ctx.debug_intersect_batch(pts, mode)
            assert np.array_equal(flags & 1, hit)
            assert np.array_equal(gx[m], x[m]) and np.array_equal(gy[m], y[m])


def test_device_int128_to_double_kat(ctx, oracle):
    vals, words = i128_kat(n_random=1000000)
    assert len(vals) >= 1000000
    want = np.array([float(v) for v in vals])  # Python: correctly rounded, ties to even
    got = ctx.debug_i128_batch(words)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # the oracle's conversion (libgcc __floattidf) agrees as well: the three are one function
    assert np.array_equal(oracle.i128_to_double(words).view(np.uint64), want.view(np.uint64))
    # quotient and truncating store, as PIP's y* = (double) num / (double) b and the
    # rational -> int64 conversion (src/util/rational.h:190-192) use them
    rng = np.random.default_rng(5)
    den = rng.integers(1, 2**47, size=len(vals)) * rng.choice([-1, 1], size=len(vals))
    dwords = np.column_stack([den.astype(np.int64).view(np.uint64),
                              np.where(den < 0, np.uint64(2**64 - 1), np.uint64(0))]).astype(np.uint64)
    cvt, div, tr = ctx.debug_i128_batch(words, dwords)
    assert np.array_equal(cvt.view(np.uint64), want.view(np.uint64))
    wdiv = want / den.astype(np.float64)
    assert np.array_equal(div.view(np.uint64), wdiv.view(np.uint64))
    ok = np.abs(wdiv) < 9.2e18
    assert np.array_equal(tr[ok], np.trunc(wdiv[ok]).astype(np.int64))


@pytest.mark.parametrize("q", [0, 1])
@pytest.mark.parametrize("name", ["lattice", "shared", "voronoi"])
def test_device_pip_update_rule(ctx, oracle, name, q):
    R, S = dataset(name)
    om = OracleMaps(oracle, [R, S])
    b = 1 - q
    xyb, p1 = om.pts[b], om.p1[b]
    if len(p1) > 4000:
        p1 = p1[:4000]
    edges = np.concatenate([xyb[p1], xyb[p1 + 1]], axis=1).astype(np.int64)
    rng = np.random.default_rng(17 + q)
    lo, hi = xyb.min(0), xyb.max(0)
    rnd = np.column_stack([rng.integers(lo[0], hi[0] + 1, 4000), rng.integers(lo[1], hi[1] + 1, 4000)])
    # vertices of BOTH maps: points with x equal to a vertex x, points on edges, coincident edges
    pts = np.ascontiguousarray(np.concatenate([om.pts[q][:3000], xyb[:3000], rnd]), np.int64)
    want = oracle.pip_brute(xyb, p1, pts, q)
    got = ctx.debug_pip_batch(edges, pts, q)
    assert np.array_equal(got, want)
    assert (got != 0xFFFFFFFF).any()
