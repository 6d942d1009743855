"""Where the end-to-end LSI step goes: upload (rjb_set_map from pinned buffers), query, read-back."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, rayjoin_b200 as RJ
from rayjoin_b200 import synth
R, S = bench.get_map("R", 1), bench.get_map("S", 2)
stream = torch.cuda.Stream()
ctx = RJ.Context(device=0, stream=stream.cuda_stream)
ctx.set_option("keep_host_graph", 0)
ctx.set_bounding_box(*synth.US_BBOX)
ctx.set_map(0, R)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
xy, row, left, right = pin(S.xy), pin(S.row_index), pin(S.left), pin(S.right)
up = lambda: ctx.set_map_raw(1, xy.data_ptr(), S.n_points, row.data_ptr(), left.data_ptr(), right.data_ptr(), S.n_chains)
up(); ctx.build_index(0, "lbvh")
lsi = RJ.LSI(ctx, "lbvh"); lsi.Init(0.1)
n = lsi.Query(1)
out = np.empty(n, RJ.XSECT_DTYPE)
for name, f in (("upload", up), ("query", lambda: lsi.Query(1)), ("readback", lambda: ctx.copy_to_host(lsi._res[0], out))):
    for _ in range(3): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): f()
    torch.cuda.synchronize(); print("%-9s %.3f ms" % (name, (time.perf_counter() - t) / 20 * 1e3))
for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21):
    ctx.set_option("load_chunk_points", chunk)
    for _ in range(3): up()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): up()
    torch.cuda.synchronize(); print("upload, chunk %8d points: %.3f ms" % (chunk, (time.perf_counter() - t) / 20 * 1e3))
