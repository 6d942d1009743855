"""Full-scale overlay parity: the County x Zipcode-scale synthetic pair through the
reference's own MapOverlayLBVH / MapOverlayGrid (oracle/_ref/ref_exec) and through this
engine; compares intersection counts, vertex locations and the -output files."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import rayjoin_b200 as RJ
from tools import ref_runner

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["lbvh", "grid"]
A, B = bench.get_map("R", 1, scale), bench.get_map("S", 2, scale)
ctx = RJ.Context([A, B], device=0)
ov = RJ.MapOverlay(ctx, "lbvh", xsect_factor=0.5)
ov.Run(); ms = ov.Run()
ours = os.path.join(bench.CACHE, "overlay_ours.cdb")
t = time.perf_counter(); ov.WriteResult(ours); tw = time.perf_counter() - t
pip = [ov.get_point_in_polygon(im) for im in range(2)]
n = len(ov.get_xsect_edges(0))
md5 = lambda p: hashlib.md5(open(p, "rb").read()).hexdigest()
print(json.dumps({"impl": "rjb200", "n_xsects": n, "device_ms": ms, "write_s": tw,
                  "bytes": os.path.getsize(ours), "md5": md5(ours)}), flush=True)
for mode in modes:
    out = os.path.join(bench.CACHE, "overlay_ref_%s.cdb" % mode)
    t = time.perf_counter()
    try:
        ref = ref_runner.run_overlay(None, A, B, mode=mode, xsect_factor=0.5, grid_size=8192,
                                     output=out, workdir=bench.CACHE, timeout=1500)
    except Exception as e:
        print(json.dumps({"impl": "reference", "mode": mode, "error": str(e)[-300:]}), flush=True)
        continue
    same_pip = [bool(np.array_equal(pip[im], ref["point_in_polygon_%d" % im])) for im in range(2)]
    print(json.dumps({"impl": "reference", "mode": mode, "n_xsects": ref["intersections"],
                      "phases": ref["phases"], "wall_s": time.perf_counter() - t,
                      "point_in_polygon_equal": same_pip, "bytes": os.path.getsize(out),
                      "md5": md5(out), "output_identical": md5(out) == md5(ours)}), flush=True)
ctx.close()
