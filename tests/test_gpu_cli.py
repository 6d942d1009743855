"""query_exec / polyover_exec: RayJoin's flags, CDB loader, -output and -check."""
import os
import subprocess

import numpy as np
import pytest

from helpers import OracleMaps, dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "rayjoin_b200", "bin")


@pytest.fixture(scope="module")
def cdb_pair(tmp_path_factory):
    from rayjoin_b200 import synth
    d = tmp_path_factory.mktemp("cdb")
    R, S = dataset("shared")
    p0, p1 = str(d / "r.cdb"), str(d / "s.cdb")
    synth.write_cdb(R, p0)
    synth.write_cdb(S, p1)
    return R, S, p0, p1, d


def _run(args):
    out = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stderr


@pytest.mark.parametrize("mode", ["lbvh", "grid", "rt"])
def test_query_exec_lsi(oracle, cdb_pair, mode):
    R, S, p0, p1, d = cdb_pair
    outp = str(d / ("lsi_%s.txt" % mode))
    err = _run([os.path.join(BIN, "query_exec"), "-poly1", p0, "-poly2=" + p1, "-mode=" + mode,
                "-query=lsi", "-xsect_factor", "2.0", "-grid_size=128", "-warmup=1", "-repeat=2",
                "-check", "-serialize=" + str(d / "ser"), "-output", outp])
    assert "Timing results:" in err and " - Query: " in err and " - Build Index: " in err
    assert "Intersections: " in err and "Queue Load Factor" in err
    if mode != "grid":  # count against the grid backend, reported like src/run_overlay.cu:56-68
        assert "LSI passed check" in err or "xsects (Answer)" in err
    om = OracleMaps(oracle, [R, S])
    # -mode=grid has the reference's grid semantics (src/app/lsi_grid.h:62-67)
    eq, eb, x, y = om.lsi_refgrid(1, 128) if mode == "grid" else om.lsi(1)
    got = np.loadtxt(outp, dtype=np.int64, ndmin=2)
    o = np.lexsort((eq.astype(np.int64), eb.astype(np.int64)))  # file is sorted by (eid0, eid1)
    want = np.column_stack([eb[o], eq[o], x[o], y[o]]).astype(np.int64)
    assert np.array_equal(got, want)


def test_query_exec_pip_map_vertices_and_generated(oracle, cdb_pair):
    R, S, p0, p1, d = cdb_pair
    outp = str(d / "pip.txt")
    err = _run([os.path.join(BIN, "query_exec"), "-poly1", p0, "-poly2", p1, "-mode=lbvh",
                "-query=pip", "-warmup=1", "-repeat=1", "-grid_size=128", "-output", outp])
    assert "passed check" in err
    om = OracleMaps(oracle, [R, S])
    assert np.array_equal(np.loadtxt(outp, dtype=np.int64).astype(np.uint32), om.pip(1, om.pts[1]))
    # generated workload: same seed -> same points -> same answers in both modes
    a, b = str(d / "gen_lbvh.txt"), str(d / "gen_grid.txt")
    for mode, o in (("lbvh", a), ("grid", b)):
        _run([os.path.join(BIN, "query_exec"), "-poly1", p0, "-mode=" + mode, "-query=pip",
              "-gen_n=5000", "-seed=7", "-warmup=0", "-repeat=1", "-nocheck", "-grid_size=128",
              "-output", o])
    ea, eb = np.loadtxt(a, dtype=np.int64), np.loadtxt(b, dtype=np.int64)
    assert len(ea) == 5000 and np.array_equal(ea, eb) and (ea != 0xFFFFFFFF).any()


def _std_uniform_real(raw, a, b):
    """libstdc++ uniform_real_distribution<double> on mt19937: generate_canonical<double, 53>
    takes two 32-bit draws (low word first), all in double arithmetic."""
    u = (raw[0::2].astype(np.float64) + raw[1::2].astype(np.float64) * 4294967296.0) / 18446744073709551616.0
    u = np.where(u >= 1.0, np.nextafter(1.0, 0.0), u)
    return u * (b - a) + a


def test_query_exec_lsi_generated_workload(oracle, rjb, cdb_pair):
    """-gen_n/-gen_t/-seed without -poly2 (GenerateLSIQueries, src/run_query.cu:101-144):
    the test re-creates the same std::mt19937 draws and checks the CLI's result against
    the oracle on those segments."""
    R, S, p0, p1, d = cdb_pair
    n, t, seed = 4000, 0.8, 11
    outp = str(d / "lsi_gen.txt")
    err = _run([os.path.join(BIN, "query_exec"), "-poly1", p0, "-mode=lbvh", "-query=lsi",
                "-gen_n=%d" % n, "-gen_t=%g" % t, "-seed=%d" % seed, "-xsect_factor", "2.0",
                "-warmup=1", "-repeat=1", "-check", "-grid_size=128", "-output", outp])
    assert "Generate Workloads" in err and ("LSI passed check" in err or "xsects (Answer)" in err)
    Rf = rjb.read_pgraph(p0)  # coordinates and bounding box as the CLI parsed them
    raw = np.random.RandomState(seed)._bit_generator.random_raw(n * 10).astype(np.uint64).reshape(n, 10)
    bx0, by0, bx1, by1 = Rf.bbox
    x1 = _std_uniform_real(raw[:, 0:2].ravel(), bx0, bx1)
    y1 = _std_uniform_real(raw[:, 2:4].ravel(), by0, by1)
    x2 = _std_uniform_real(raw[:, 4:6].ravel(), bx0, bx1)
    y2 = _std_uniform_real(raw[:, 6:8].ravel(), by0, by1)
    tt = _std_uniform_real(raw[:, 8:10].ravel(), 0.0, t)
    ln = np.sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1))
    xy = np.column_stack([x1, y1, x1 + tt * ((x2 - x1) / ln), y1 + tt * ((y2 - y1) / ln)]).reshape(-1, 2)
    Q = rjb.PlanarGraph(xy, np.arange(0, 2 * n + 1, 2, dtype=np.uint32), np.zeros(n, np.int64),
                        np.zeros(n, np.int64))
    om = OracleMaps(oracle, [Rf, Q])
    om.sc = oracle.scaling_init(*Rf.bbox)  # the context is built on the base map alone
    om.pts = [oracle.scale_points(om.sc, g.xy) for g in (Rf, Q)]
    eq, eb, x, y = om.lsi(1, brute=True)
    got = np.loadtxt(outp, dtype=np.int64, ndmin=2)
    o = np.lexsort((eq.astype(np.int64), eb.astype(np.int64)))
    want = np.column_stack([eb[o], eq[o], x[o], y[o]]).astype(np.int64)
    assert len(want) > 100 and np.array_equal(got, want)


def test_polyover_exec_output_matches_api(rjb, cdb_pair):
    R, S, p0, p1, d = cdb_pair
    outp = str(d / "overlay.cdb")
    err = _run([os.path.join(BIN, "polyover_exec"), "-poly1", p0, "-poly2", p1, "-mode=lbvh",
                "-xsect_factor=2.0", "-check", "-grid_size=128", "-output", outp])
    assert "LSI passed check" in err and err.count("PIP passed check") == 2
    ctx = rjb.Context([rjb.read_pgraph(p0), rjb.read_pgraph(p1)])
    ov = rjb.MapOverlay(ctx, "grid", grid_size=64, xsect_factor=2.0)
    ov.Run()
    api = str(d / "overlay_api.cdb")
    ov.WriteResult(api)
    ctx.close()
    assert open(outp).read() == open(api).read() and os.path.getsize(outp) > 1000


def test_bad_flags_fail_loudly(cdb_pair):
    R, S, p0, p1, d = cdb_pair
    for args in (["-poly1", p0, "-poly2", p1, "-mode=quadtree", "-query=lsi"],
                 ["-poly1", p0, "-poly2", p1, "-mode=lbvh", "-query=knn"],
                 ["-poly1", str(d / "nope.cdb"), "-poly2", p1, "-mode=lbvh", "-query=lsi"],
                 ["-poly1", p0, "-bogus_flag=1"]):
        out = subprocess.run([os.path.join(BIN, "query_exec")] + args, capture_output=True, text=True)
        assert out.returncode != 0 and "FATAL" in out.stderr
