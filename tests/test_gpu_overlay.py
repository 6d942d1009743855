"""Overlay parity: device results and the written chain file against the CPU
restatement, and against the reference's own polyover path (ref_exec) when built."""
import os
import sys

import numpy as np
import pytest

from helpers import dataset

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.overlay_oracle import OverlayOracle  # noqa: E402
from tools import ref_runner  # noqa: E402

pytestmark = pytest.mark.gpu


def _rows(xs):
    return [[int(r["eid"][0]), int(r["eid"][1]), int(r["x"]), int(r["y"]), int(r["mid_point_polygon_id"])]
            for r in xs]


@pytest.mark.parametrize("mode", ["lbvh", "grid", "brute"])
@pytest.mark.parametrize("name", ["voronoi", "shared", "lattice", "tiny"])
def test_overlay_matches_oracle(rjb, name, mode, tmp_path):
    A, B = dataset(name)
    if name in ("voronoi", "shared"):  # keep the pure-Python oracle quick
        from rayjoin_b200 import synth
        A = synth.voronoi_map(25, 900, synth.BRAZIL_BBOX, seed=21)
        B = synth.voronoi_map(60, 1200, synth.BRAZIL_BBOX, seed=22)
        if name == "shared":
            B = synth.share_chains(A, B, frac=0.3)
    oo = OverlayOracle([A, B], grid_size=64 if mode == "grid" else None).run()
    ctx = rjb.Context([A, B])
    ov = rjb.MapOverlay(ctx, mode, grid_size=64, xsect_factor=4.0)
    ov.Run()
    for im in range(2):
        assert np.array_equal(ov.get_closet_eids(im), oo.closest[im])
        assert np.array_equal(ov.get_point_in_polygon(im), oo.pip[im])
        assert _rows(ov.get_xsect_edges(im)) == oo.sorted[im]
    ours, want = str(tmp_path / "ours.cdb"), str(tmp_path / "oracle.cdb")
    ov.WriteResult(ours)
    oo.write(want)
    assert open(ours).read() == open(want).read()
    ctx.close()


@pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref/ref_exec not built")
@pytest.mark.parametrize("ref_mode", ["lbvh", "grid"])
def test_overlay_output_matches_reference_binary(rjb, ref_mode, tmp_path):
    """The reference's own test is `diff` of the -output file (test/test_overlay.sh:15-21)."""
    from rayjoin_b200 import synth
    A = synth.voronoi_map(40, 2500, synth.BRAZIL_BBOX, seed=31)
    B = synth.voronoi_map(120, 3500, synth.BRAZIL_BBOX, seed=32)
    ref_out = str(tmp_path / "ref.cdb")
    ref = ref_runner.run_overlay(None, A, B, mode=ref_mode, xsect_factor=4.0, grid_size=256,
                                 output=ref_out, workdir=str(tmp_path))
    ctx = rjb.Context([A, B])
    ov = rjb.MapOverlay(ctx, "lbvh", xsect_factor=4.0)
    ov.Run()
    ours = str(tmp_path / "ours.cdb")
    ov.WriteResult(ours)
    for im in range(2):
        assert np.array_equal(ov.get_point_in_polygon(im), ref["point_in_polygon_%d" % im])
    assert ref["intersections"] == len(ov.get_xsect_edges(0))
    assert open(ours).read() == open(ref_out).read()
    ctx.close()
