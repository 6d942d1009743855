"""Scalability sweep (BASELINE.json configs[4]): LSI on polygon soups like the reference's
synthetic runs (expr/run_scalability.sh: R = 5M polygons, S = 1M..5M polygons, uniform and
gaussian, generator.py polysize 0.001, 3..10 segments).  One JSON line per point; every
result is checked against the host oracle (pair set + coordinates)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import rayjoin_b200 as RJ
from rayjoin_b200 import synth
from oracle import oracle as O
from helpers import sort_xsects
from tools import ref_runner

ap = argparse.ArgumentParser()
ap.add_argument("--dists", default="uniform,gaussian")
ap.add_argument("--r-polys", type=int, default=5_000_000)
ap.add_argument("--s-polys", default="1000000,3000000,5000000")
ap.add_argument("--xsect-factor", type=float, default=1.0)
ap.add_argument("--modes", default="lbvh,grid")
ap.add_argument("--ref", type=int, default=1)
ap.add_argument("--repeat", type=int, default=5)
args = ap.parse_args()

for dist in args.dists.split(","):
    t = time.time()
    R = synth.polygon_soup(args.r_polys, dist, seed=1)
    print("# R %s: %d edges (%.1fs)" % (dist, R.n_edges, time.time() - t), file=sys.stderr)
    for ns in [int(x) for x in args.s_polys.split(",")]:
        S = synth.polygon_soup(ns, dist, seed=2)
        bbox = synth.union_bbox(R, S)
        ctx = RJ.Context([R, S], device=0, bbox=bbox)
        ctx.set_option("keep_host_graph", 0)
        sc = O.scaling_init(*bbox)
        r, s = O.scale_points(sc, R.xy), O.scale_points(sc, S.xy)
        rp1, _ = O.build_edges(R.row_index); sp1, _ = O.build_edges(S.row_index)
        t = time.perf_counter(); want = O.lsi_grid(s, sp1, r, rp1, sc); t_cpu = time.perf_counter() - t
        for mode in args.modes.split(","):
            for sq in ((0, 1, -1) if mode == "lbvh" else (0,)):
                ctx.set_option("sort_queries", sq)
                build = min(ctx.build_index(0, mode, 8192) for _ in range(2))
                lsi = RJ.LSI(ctx, mode); lsi.Init(args.xsect_factor)
                ts = []
                for _ in range(args.repeat + 1):
                    t = time.perf_counter(); n = lsi.Query(1); ts.append(time.perf_counter() - t)
                got = sort_xsects(lsi.get_xsects(), 1)
                ok = n == len(want[0]) and all(np.array_equal(g, w) for g, w in zip(got, want))
                print(json.dumps({"dist": dist, "mode": mode, "sort_queries": sq, "R_edges": R.n_edges,
                                  "S_edges": S.n_edges, "pairs": int(n), "build_ms": build,
                                  "query_ms": min(ts[1:]) * 1e3, "kernel_ms": ctx.last_kernel_ms(),
                                  "survivors": ctx.last_stats()[7], "candidates": lsi.n_candidates,
                                  "parity_vs_oracle": "bit-exact" if ok else "MISMATCH",
                                  "oracle_ms_%dcores" % O.num_threads(): t_cpu * 1e3}), flush=True)
        ctx.close()
        if args.ref and ref_runner.available():
            for rm in ("lbvh",):
                try:
                    out = ref_runner.run_lsi(None, R, S, mode=rm, warmup=2, repeat=3,
                                             xsect_factor=args.xsect_factor)
                    print(json.dumps({"dist": dist, "impl": "reference", "mode": rm, "S_edges": S.n_edges,
                                      "pairs": out["intersections"], "phases": out["phases"]}), flush=True)
                except Exception as e:
                    print(json.dumps({"impl": "reference", "mode": rm, "error": str(e)[:200]}), flush=True)
