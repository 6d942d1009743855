"""CPU suite: the C-ABI library loads, exports every symbol include/rjb200.h
declares, fails loudly without a GPU, and its host-only CDB functions work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rjb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rjb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(rjb):
    lib = rjb.load_library()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "librjb200.so does not export %s" % n
    from rayjoin_b200 import capi
    assert sorted(capi.EXPORTS) == names


def test_no_cpu_fallback(rjb):
    """Without a CUDA device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rjb.RjbError) as ei:
        rjb.Context(device=0)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rayjoin_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in src.replace("oracle/", "ORACLE_DOC/") or f == "synth.py" or \
                    not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"oracle", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_cdb_text_bin_roundtrip(rjb, tmp_path):
    from rayjoin_b200 import synth
    g = synth.voronoi_map(20, 400, synth.BRAZIL_BBOX, seed=9)
    txt = str(tmp_path / "m.cdb")
    synth.write_cdb(g, txt)
    a = rjb.read_pgraph(txt)
    assert np.array_equal(a.xy, g.xy) and np.array_equal(a.row_index, g.row_index)
    assert np.array_equal(a.left, g.left) and np.array_equal(a.right, g.right)
    assert a.bbox == (g.xy[:, 0].min(), g.xy[:, 1].min(), g.xy[:, 0].max(), g.xy[:, 1].max())
    # load_from writes <prefix>/<path with '/' -> '-'>.bin and reads it next time
    prefix = str(tmp_path / "ser")
    b = rjb.load_from(txt, prefix)
    ser = os.path.join(prefix, txt.replace("/", "-") + ".bin")
    assert os.path.exists(ser)
    os.rename(txt, txt + ".moved")  # second load must come from the cache
    c = rjb.load_from(txt, prefix)
    for x in (b, c):
        assert np.array_equal(x.xy, g.xy) and np.array_equal(x.row_index, g.row_index)
    # .bin layout of the reference (planar_graph.h:129-167)
    raw = open(ser, "rb").read()
    magic, n_chains, n_row, n_points = np.frombuffer(raw[:32], np.uint64)
    assert magic == 0xabcdabcd and n_chains == g.n_chains and n_row == g.n_chains + 1
    assert n_points == g.n_points and np.frombuffer(raw[-8:], np.uint64)[0] == 0xabcdabcd
    assert len(raw) == 32 + 40 * g.n_chains + 4 * (g.n_chains + 1) + 16 * g.n_points + 32 + 8


def test_cdb_parser_rules(rjb, tmp_path):
    p = tmp_path / "a.cdb"
    p.write_text("# comment\n\n% other comment\n7 3 0 2 1 2\n0 0\n1 1\n2 0\n8 2 2 3 0 1\n2 0\n3 5\n")
    g = rjb.read_pgraph(str(p))
    assert g.n_chains == 2 and g.n_points == 5 and list(g.row_index) == [0, 3, 5]
    assert list(g.chain_id) == [7, 8] and list(g.left) == [1, 0] and list(g.right) == [2, 1]
    for bad in ("1 1 0 0 0 0\n0 0\n",                 # np < 2
                "1 2 0 1 0 0\n0 0\n0 0\n",            # consecutive duplicate point
                "1 3 0 1 0 0\n0 0\n1 1\n",            # truncated chain
                "1 2 0 1 0 0\n0 zero\n1 1\n"):        # unparsable coordinate
        p.write_text(bad)
        with pytest.raises(rjb.RjbError) as ei:
            rjb.read_pgraph(str(p))
        assert ei.value.code == 4
    p.write_text("")
    e = rjb.read_pgraph(str(p))
    assert e.n_chains == 0 and e.n_points == 0 and len(e.row_index) == 0
    with pytest.raises(rjb.RjbError):
        rjb.read_pgraph(str(tmp_path / "missing.cdb"))


def test_synthetic_maps_are_well_formed():
    from rayjoin_b200 import synth
    g = synth.voronoi_map(50, 2000, synth.US_BBOX, seed=3)
    assert g.row_index[0] == 0 and g.row_index[-1] == g.n_points
    assert (np.diff(g.row_index.astype(np.int64)) >= 2).all()
    d = np.diff(g.xy, axis=0)
    same = (d == 0).all(axis=1)
    same[g.row_index[1:-1].astype(np.int64) - 1] = False
    assert not same.any()
    assert abs(g.n_edges - 2000) < 200
    assert (g.left != g.right).all()
    s = synth.polygon_soup(100, "gaussian", seed=4)
    first, last = s.row_index[:-1].astype(np.int64), s.row_index[1:].astype(np.int64) - 1
    assert np.array_equal(s.xy[first], s.xy[last])  # closed rings


def test_parallel_cdb_parser_equals_sequential(rjb, tmp_path, monkeypatch):
    """Files > 1 MiB take the multi-threaded from_chars path; it must give the same graph
    as the sequential strtod path (and fall back to it on anything unusual)."""
    from rayjoin_b200 import synth
    g = synth.voronoi_map(300, 60000, synth.US_BBOX, seed=13)
    p = str(tmp_path / "big.cdb")
    synth.write_cdb(g, p)
    assert os.path.getsize(p) > (1 << 20)
    a = rjb.read_pgraph(p)
    monkeypatch.setenv("RJB_CDB_SEQUENTIAL", "1")
    b = rjb.read_pgraph(p)
    monkeypatch.delenv("RJB_CDB_SEQUENTIAL")
    for x in (a, b):
        assert np.array_equal(x.xy, g.xy) and np.array_equal(x.row_index, g.row_index)
        assert np.array_equal(x.left, g.left) and np.array_equal(x.right, g.right)
    assert a.bbox == b.bbox
    # an error in the middle of a big file is still reported with file and line number
    lines = open(p).read().split("\n")
    lines[40000] = "not a number"
    open(p, "w").write("\n".join(lines))
    with pytest.raises(rjb.RjbError) as ei:
        rjb.read_pgraph(p)
    assert ei.value.code == 4 and "[40001]" in str(ei.value)
