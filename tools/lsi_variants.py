"""LSI query on the bench workload under several option sets, one process (maps generated once):
per-kernel CUDA-event times, counts, and a result check between the variants.
  python tools/lsi_variants.py [--steps 20] "lsi_tile_filter=0,lsi_cells=0" "lsi_tile_filter=1,lsi_cells=1" ..."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, rayjoin_b200 as RJ
from rayjoin_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--share-chains", type=float, default=0.0)
ap.add_argument("--s-seed", type=int, default=2, help="seed of the query map (bench rank r uses 2 + r)")
ap.add_argument("variants", nargs="+")
args = ap.parse_args()
R, S = bench.get_map("R", 1), bench.get_map("S", args.s_seed)
if args.share_chains > 0:
    S = synth.share_chains(R, S, frac=args.share_chains, seed=3)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for var in args.variants:
    ctx = RJ.Context(device=0, stream=stream.cuda_stream)
    ctx.set_option("keep_host_graph", 0)
    ctx.set_option("stage_timing", 1)
    for kv in var.split(","):
        if kv:
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
    ctx.set_bounding_box(*synth.US_BBOX)
    ctx.set_map(0, R)
    ctx.set_map(1, S)
    build = min(ctx.build_index(0, "lbvh") for _ in range(3))
    lsi = RJ.LSI(ctx, "lbvh")
    lsi.Init(0.1)
    for _ in range(5):
        n = lsi.Query(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage = []
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
            ev[i][0].record(stream)
            lsi.Launch(1)
            ev[i][1].record(stream)
        n = lsi.Wait()
        stage.append(ctx.last_stage_ms()[0])
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    xs = lsi.get_xsects()
    key = np.sort(xs["eid"][:, 0].astype(np.int64) << 32 | xs["eid"][:, 1])
    if ref is None:
        ref = key
    st = ctx.last_stats()
    print(json.dumps({"variant": var, "query_ms": float(np.mean(ms)), "median_ms": float(np.median(ms)),
                      "stage_ms": [round(float(v), 4) for v in np.mean(np.asarray(stage), axis=0)],
                      "pairs": int(n), "candidates": int(lsi.n_candidates), "ql_pairs": int(st[2]),
                      "survivors": int(st[7]), "cells_path": int(st[5]), "long": int(st[6]),
                      "build_ms": build, "index_MB": ctx.index_info(0, "lbvh")["bytes"] / 1e6,
                      "same_pairs_as_first": bool(np.array_equal(key, ref))}), flush=True)
    ctx.close()
