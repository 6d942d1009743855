"""Runs oracle/_ref/ref_exec (the reference's own grid / lbvh CUDA backends built
from /root/reference by oracle/Makefile) on PlanarGraph inputs.

TEST / BASELINE INFRASTRUCTURE: used by tests (parity pin on the GPU box) and by
bench.py --impl reference.  Inputs are handed over as RayJoin .bin graphs
(layout of reference src/map/planar_graph.h:128-167), written here with numpy so
that nothing of the product is on the reference's path.
"""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXEC = os.path.join(ROOT, "oracle", "_ref", "ref_exec")


def available():
    return os.path.exists(REF_EXEC)


def write_bin(g, path):
    magic = np.array([0xabcdabcd], np.uint64)
    n_chains, n_points = g.n_chains, g.n_points
    n_row = n_chains + 1 if n_points else 0
    with open(path, "wb") as f:
        magic.tofile(f)
        np.array([n_chains, n_row, n_points], np.uint64).tofile(f)
        ch = np.column_stack([g.chain_id, g.first_point, g.last_point, g.left, g.right]).astype(np.int64)
        ch.tofile(f)
        g.row_index.astype(np.uint32)[:n_row].tofile(f)
        g.xy.astype(np.float64).tofile(f)
        np.array([g.bbox[0], g.bbox[1], g.bbox[2], g.bbox[3]], np.float64).tofile(f)
        magic.tofile(f)


def _maps(R, S, workdir, tag):
    os.makedirs(workdir, exist_ok=True)
    p0 = os.path.join(workdir, "ref_%s_map0.bin" % tag)
    p1 = os.path.join(workdir, "ref_%s_map1.bin" % tag)
    write_bin(R, p0)
    write_bin(S, p1)
    return p0, p1


def _run(cmd, timeout):
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError("ref_exec failed (%d): %s" % (out.returncode, out.stderr[-2000:]))
    last = [l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1]
    return json.loads(last)


def run_lsi(ref_exec, R, S, mode="lbvh", warmup=5, repeat=5, xsect_factor=0.2, grid_size=2048,
            workdir="/tmp/rjb200_cache", dump=False, timeout=1200):
    """query_exec -query=lsi semantics: base map 0 = R, query map 1 = S."""
    p0, p1 = _maps(R, S, workdir, "%d_%d" % (R.n_edges, S.n_edges))
    cmd = [ref_exec or REF_EXEC, "lsi", mode, p0, p1, "-warmup", str(warmup), "-repeat", str(repeat),
           "-xsect_factor", str(xsect_factor), "-grid_size", str(grid_size)]
    dpath = os.path.join(workdir, "ref_lsi_dump_%s.txt" % mode)
    if dump:
        cmd += ["-dump", dpath]
    res = _run(cmd, timeout)
    res["query_ms"] = res["phases"]["query_ms"]
    if dump:
        a = np.loadtxt(dpath, dtype=np.int64, ndmin=2) if os.path.getsize(dpath) else np.zeros((0, 6), np.int64)
        # columns: eid_query(map1) eid_base(map0) x.num x.den y.num y.den
        res["pairs"] = a
    return res


def run_pip(ref_exec, R, S, mode="lbvh", warmup=1, repeat=1, grid_size=2048,
            workdir="/tmp/rjb200_cache", timeout=1200):
    """query_exec -query=pip -poly2: vertices of map 1 located in map 0."""
    p0, p1 = _maps(R, S, workdir, "%d_%d" % (R.n_edges, S.n_edges))
    dpath = os.path.join(workdir, "ref_pip_dump_%s.bin" % mode)
    res = _run([ref_exec or REF_EXEC, "pip", mode, p0, p1, "-warmup", str(warmup), "-repeat",
                str(repeat), "-grid_size", str(grid_size), "-dump", dpath], timeout)
    res["query_ms"] = res["phases"]["query_ms"]
    res["closest_eids"] = np.fromfile(dpath, dtype=np.uint32)
    return res


def run_overlay(ref_exec, M0, M1, mode="lbvh", xsect_factor=0.5, grid_size=2048, output=None,
                workdir="/tmp/rjb200_cache", timeout=1200):
    p0, p1 = _maps(M0, M1, workdir, "ov_%d_%d" % (M0.n_edges, M1.n_edges))
    dpath = os.path.join(workdir, "ref_ov_dump_%s" % mode)
    cmd = [ref_exec or REF_EXEC, "overlay", mode, p0, p1, "-xsect_factor", str(xsect_factor),
           "-grid_size", str(grid_size), "-dump", dpath]
    if output:
        cmd += ["-output", output]
    res = _run(cmd, timeout)
    for im, g in enumerate((M0, M1)):
        raw = np.fromfile(dpath + ".pip%d" % im, dtype=np.uint32)
        res["closest_eids_%d" % im] = raw[:g.n_points]
        res["point_in_polygon_%d" % im] = raw[g.n_points:].view(np.int32)
    return res
