"""Traversal statistics of the LBVH LSI kernel on the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import rayjoin_b200 as RJ
from rayjoin_b200 import synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
R, S = bench.get_map("R", 1, scale), bench.get_map("S", 2, scale)
for leaf in (4,):
    for sortq in (0, 1):
        ctx = RJ.Context(device=0)
        ctx.set_option("keep_host_graph", 0)
        ctx.set_option("lbvh_leaf_size", leaf)
        ctx.set_option("sort_queries", sortq)
        ctx.set_option("stats", 1)
        ctx.set_bounding_box(*synth.US_BBOX)
        ctx.set_map(0, R); ctx.set_map(1, S)
        ctx.build_index(0, "lbvh")
        lsi = RJ.LSI(ctx, "lbvh"); lsi.Init(0.1)
        lsi.Query(1); lsi.Query(1)
        st = ctx.last_stats()
        warps = (S.n_edges + 31) // 32
        print("leaf %d sort %d: warps %d results %d cand %d | binary node visits/warp %.2f leaf visits/warp %.2f top steps/warp %.2f lane-leaf tests/warp %.2f warps with leaf %.1f%% max stack %d kernel ms %.3f + %.3f"
              % (leaf, sortq, warps, st[0], st[1], st[2] / warps, st[3] / warps, st[4] / warps, st[5] / warps, 100.0 * st[6] / warps, st[7], *ctx.last_kernel_ms()))
        ctx.close()
