"""Phase timeline of k_lsi_resolve (a -DRJB_TRACE build selected with RJB_LIB)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, rayjoin_b200 as RJ
from rayjoin_b200 import synth, capi
R, S = bench.get_map("R", 1), bench.get_map("S", 2)
stream = torch.cuda.Stream()
ctx = RJ.Context(device=0, stream=stream.cuda_stream)
ctx.set_option("keep_host_graph", 0)
for kv in (sys.argv[1] if len(sys.argv) > 1 else "lsi_cells=1").split(","):
    k, v = kv.split("="); ctx.set_option(k, int(v))
ctx.set_bounding_box(*synth.US_BBOX); ctx.set_map(0, R); ctx.set_map(1, S)
ctx.build_index(0, "lbvh")
lsi = RJ.LSI(ctx, "lbvh"); lsi.Init(0.1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    if os.environ.get("NOFLUSH") != "1":
        with torch.cuda.stream(stream): flush.zero_()
    lsi.Query(1)
lib = ctx.lib
n_cta = 4096
buf = np.zeros(n_cta * 16, np.uint64)
lib.rjb_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert lib.rjb_debug_trace(ctx._h, buf.ctypes.data, buf.size) == 0
t = buf.reshape(n_cta, 16).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = {0: "start", 1: "count", 10: "loop done", 11: "final drain", 12: "final flush", 13: "end"}
print("CTAs", len(t), "kernel span us", (t[:, 13].max() - t0) / 1e3)
for k in range(14):
    col = t[:, k]; m = col > 0
    if m.any():
        v = (col[m] - t0) / 1e3
        print("mark %2d %-12s n=%4d  min %7.2f  median %7.2f  max %7.2f us" % (k, names.get(k, "round/drain"), m.sum(), v.min(), np.median(v), v.max()))
d = (t[:, 13] - t[:, 0]) / 1e3
print("CTA lifetime us: min %.2f median %.2f max %.2f" % (d.min(), np.median(d), d.max()))
for a, b, nm in ((10, 11, "final drain"), (11, 12, "final flush (gcd)"), (12, 6, "tail: barrier"), (6, 7, "tail: fence"), (7, 8, "tail: ticket atomic"), (8, 13, "tail: rest")):
    v = (t[:, b] - t[:, a]) / 1e3
    print("%-18s median %.2f max %.2f us" % (nm, np.median(v), v.max()))

# per-warp view: when does lane 0 of every warp pass the marks around the end of the kernel?
bw = np.zeros(n_cta * 4 * 16, np.uint64)
assert lib.rjb_debug_trace(ctx._h, bw.ctypes.data, bw.size) == 0
w = bw.reshape(n_cta, 4, 16).astype(np.int64)
w = w[w[:, 0, 0] > 0]
for k, nm in ((10, "loop done"), (11, "final drain done"), (12, "after barrier A"), (6, "after barrier B"), (13, "end")):
    col = w[:, :, k]
    spread = (col.max(1) - col.min(1)) / 1e3
    print("mark %2d %-18s spread between the warps of a CTA: median %.2f max %.2f us" % (k, nm, np.median(spread), spread.max()))
cta = 5
print("CTA", cta, "per-warp times (us since kernel start):")
for k in (10, 11, 14, 15, 12, 6, 13):
    print("  mark", k, [(round((int(x) - t0) / 1e3, 2)) for x in w[cta, :, k]])
