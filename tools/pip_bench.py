"""PIP benchmark (BASELINE.json configs[2]): N uniform random points against a
BlockGroup-scale synthetic map (~220k faces, ~28M edges).  Points are generated on
the device as scaled int64 pairs; results are checked against the host oracle on a
sample.  Prints one JSON line per mode."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import rayjoin_b200 as RJ
from rayjoin_b200 import synth
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=100_000_000)
ap.add_argument("--faces", type=int, default=220_000)
ap.add_argument("--edges", type=int, default=28_000_000)
ap.add_argument("--modes", default="lbvh,grid")
ap.add_argument("--grid-size", type=int, default=8192)
ap.add_argument("--sort", default="0,1")  # PIP: points are ordered only on request
ap.add_argument("--repeat", type=int, default=3)
ap.add_argument("--check", type=int, default=200_000, help="points checked against the oracle (-1: all of them)")
ap.add_argument("--stats", type=int, default=0)
ap.add_argument("--park", default="0", help="option pip_park values to run (A/B), e.g. 0,1")
ap.add_argument("--sort-bits", default="24", help="option pip_sort_bits values to run (grid mode)")
args = ap.parse_args()

t0 = time.time()
path = os.path.join(bench.CACHE, "pipmap_%d_%d.npz" % (args.faces, args.edges))
os.makedirs(bench.CACHE, exist_ok=True)
if os.path.exists(path):
    z = np.load(path); R = RJ.PlanarGraph(z["xy"], z["row_index"], z["left"], z["right"])
else:
    R = synth.voronoi_map(args.faces, args.edges, synth.US_BBOX, seed=1)
    np.savez(path, xy=R.xy, row_index=R.row_index, left=R.left, right=R.right)
print("map: %d edges, %d chains (%.1fs)" % (R.n_edges, R.n_chains, time.time() - t0), file=sys.stderr)
dev = torch.device("cuda", 0)
ctx = RJ.Context(device=0)
ctx.set_option("keep_host_graph", 0)
ctx.set_bounding_box(*synth.US_BBOX)
ctx.set_map(0, R)
s = ctx.get_scaling()
# uniform points in the (scaled) bounding box of the map, seed 1
g = torch.Generator(device=dev); g.manual_seed(1)
sc = O.scaling_init(*synth.US_BBOX)
lo = O.scale_points(sc, np.array([[synth.US_BBOX[0], synth.US_BBOX[1]]]))[0]
hi = O.scale_points(sc, np.array([[synth.US_BBOX[2], synth.US_BBOX[3]]]))[0]
pts = torch.empty((args.points, 2), dtype=torch.int64, device=dev)
pts[:, 0] = torch.randint(int(lo[0]), int(hi[0]), (args.points,), generator=g, device=dev, dtype=torch.int64)
pts[:, 1] = torch.randint(int(lo[1]), int(hi[1]), (args.points,), generator=g, device=dev, dtype=torch.int64)
torch.cuda.synchronize()
om_pts = None
for mode in args.modes.split(","):
    build = min(ctx.build_index(0, mode, args.grid_size) for _ in range(2))
    for sq, park, sbits in [(int(x), int(y), int(z)) for x in args.sort.split(",") for y in args.park.split(",")
                            for z in (args.sort_bits.split(",") if mode == "grid" else ["24"])]:
        if sq == 0 and sbits != int(args.sort_bits.split(",")[0]):
            continue
        ctx.set_option("sort_queries", sq)
        ctx.set_option("pip_sort_bits", sbits)
        ctx.set_option("pip_park", park if mode == "lbvh" else 0)
        times, kms = [], []
        for it in range(args.repeat + 1):
            torch.cuda.synchronize(); t = time.perf_counter()
            de, df, cand = ctx.pip_device(1, mode, pts.data_ptr(), args.points)
            dt = time.perf_counter() - t
            if it: times.append(dt * 1e3); kms.append(ctx.last_kernel_ms()[0]); stg = ctx.last_stage_ms()[0]
        out = {"query": "pip", "mode": mode, "sort_queries": sq, "pip_park": park, "pip_sort_bits": sbits, "points": args.points,
               "edges": R.n_edges, "build_ms": build, "query_ms": float(np.min(times)),
               "kernel_ms": float(np.min(kms)), "points_per_s": args.points / (np.min(times) / 1e3),
               "order_ms": stg[0], "candidates": cand}
        if args.stats:
            ctx.set_option("stats", 1)
            ctx.pip_device(1, mode, pts.data_ptr(), args.points)
            st = ctx.last_stats(); w = (args.points + 31) // 32
            out["stats_per_warp"] = {"binary_nodes": st[2] / w, "leaves": st[3] / w, "top_steps": st[4] / w,
                                     "lane_leaf": st[5] / w, "max_stack": st[7], "cand_per_point": st[1] / args.points}
            ctx.set_option("stats", 0)
        if args.check:
            n = args.points if args.check < 0 else min(args.check, args.points)
            eids = ctx.copy_to_host(de, np.empty(args.points, np.uint32))[:n]
            if om_pts is None:
                om_pts = O.scale_points(sc, R.xy); om_p1, _ = O.build_edges(R.row_index)
                O.set_num_threads(len(os.sched_getaffinity(0)))
                t_or = time.perf_counter()
                want = O.pip_grid(om_pts, om_p1, sc, pts[:n].cpu().numpy(), 1)
                t_or = time.perf_counter() - t_or
            out["oracle_s_%dcores" % O.num_threads()] = t_or
            out["points_checked"] = n
            out["parity_vs_oracle"] = "bit-exact" if np.array_equal(eids, want) else "MISMATCH (%d)" % int((eids != want).sum())
            out["hit_fraction"] = float((eids != 0xFFFFFFFF).mean())
        print(json.dumps(out), flush=True)
ctx.close()
