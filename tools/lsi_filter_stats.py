import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import rayjoin_b200 as RJ
from rayjoin_b200 import synth
R, S = bench.get_map("R", 1, 1.0), bench.get_map("S", 2, 1.0)
ctx = RJ.Context(device=0)
ctx.set_option("keep_host_graph", 0)
ctx.set_option("lsi_filter", 1)
ctx.set_bounding_box(*synth.US_BBOX)
ctx.set_map(0, R); ctx.set_map(1, S)
ctx.build_index(0, "lbvh")
lsi = RJ.LSI(ctx, "lbvh"); lsi.Init(0.1)
for i in range(3):
    lsi.Query(1)
    st = ctx.last_stats()
    print("results %d cand %d survivors %d of %d (%.2f%%) kernel ms %.3f + %.3f" % (st[0], st[1], st[7], S.n_edges, 100.0 * st[7] / S.n_edges, *ctx.last_kernel_ms()))
ctx.close()
