"""Single-GPU polygon overlay of the County x Zipcode-scale synthetic pair: phase
times (CUDA events), warm, plus traversal statistics of the two vertex-location passes."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import rayjoin_b200 as RJ

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--modes", default="lbvh,grid")
ap.add_argument("--grid-size", type=int, default=8192)
ap.add_argument("--repeat", type=int, default=3)
ap.add_argument("--stats", type=int, default=0)
ap.add_argument("--sort", type=int, default=0)
args = ap.parse_args()
A, B = bench.get_map("R", 1, args.scale), bench.get_map("S", 2, args.scale)
ctx = RJ.Context([A, B], device=0)
ctx.set_option("sort_queries", args.sort)
for mode in args.modes.split(","):
    ov = RJ.MapOverlay(ctx, mode, grid_size=args.grid_size, xsect_factor=0.5)
    best = None
    for _ in range(args.repeat):
        ms = ov.Run()
        if best is None or ms["total"] < best["total"]:
            best = dict(ms)
    out = {"query": "overlay", "mode": mode, "edges": [A.n_edges, B.n_edges], "phase_ms": best,
           "n_xsects": int(len(ov.get_xsect_edges(0)))}
    if args.stats and mode == "lbvh":
        ctx.set_option("stats", 1)
        for q in (0, 1):
            ctx.pip_device(q, mode)
            st = ctx.last_stats(); n = ctx.map_info(q)["points"]; w = (n + 31) // 32
            out["pip%d_stats_per_warp" % q] = {"binary_nodes": st[2] / w, "leaves": st[3] / w,
                                               "top_steps": st[4] / w, "cand_per_point": st[1] / n,
                                               "kernel_ms": ctx.last_kernel_ms()[0]}
        ctx.set_option("stats", 0)
    print(json.dumps(out), flush=True)
ctx.close()
