"""LSI with map 0 (R, 4 M edges) as the query side against the index of map 1 (S, 9 M edges): the
direction the overlay runs.  Option sets as in tools/lsi_variants.py."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, rayjoin_b200 as RJ
from rayjoin_b200 import synth
R, S = bench.get_map("R", 1), bench.get_map("S", 2)
stream = torch.cuda.Stream()
for var in sys.argv[1:]:
    ctx = RJ.Context(device=0, stream=stream.cuda_stream)
    ctx.set_option("keep_host_graph", 0)
    for kv in var.split(","):
        if kv:
            k, v = kv.split("="); ctx.set_option(k, int(v))
    ctx.set_bounding_box(*synth.US_BBOX); ctx.set_map(0, R); ctx.set_map(1, S)
    ctx.build_index(1, "lbvh")
    lsi = RJ.LSI(ctx, "lbvh"); lsi.Init(0.5)
    for _ in range(3): n = lsi.Query(0)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        with torch.cuda.stream(stream):
            a.record(stream); lsi.Launch(0); b.record(stream)
        lsi.Wait()
    torch.cuda.synchronize()
    st = ctx.last_stats()
    print(json.dumps({"variant": var, "ms": float(np.median([a.elapsed_time(b) for a, b in ev])), "pairs": int(n),
                      "stage_ms": [round(v, 4) for v in ctx.last_stage_ms()[0]], "survivors": int(st[7]), "cells": int(st[5])}), flush=True)
    ctx.close()
