"""Scalability sweep (BASELINE.json configs[4]) on 1..8 GPUs: LSI on polygon soups shaped like
the reference's synthetic runs (expr/run_scalability.sh:12-23: R = 5 M polygons fixed, S swept;
misc/gen_polys.sh: radius 0.001, 4..10 segments, uniform and gaussian centres), S from 1 M to
100 M segments, -xsect_factor 0.1 / 0.2 / 0.5, plus a shared-chain variant (S shares 5 % of R's
chains, every second vertex: the degenerate contacts real layers are full of).

One process per GPU (torchrun); R and its index are replicated, ONE S is cut by whole chains
over the ranks (rayjoin_b200.dist.shard_graph), NCCL carries the count all-gather and the
grouped device-side gather of the result records.  Rank 0 prints one JSON line per point:
join ms (CUDA events, max over ranks), build ms, pairs, parity of the GATHERED result against
the host oracle (pair set + coordinates, bit-exact), and -- at N = 1, where the single-GPU
reference can be compared -- the reference's -mode=lbvh and -mode=grid on the same input.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_multi.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import rayjoin_b200 as RJ  # noqa: E402
from rayjoin_b200 import dist as rd, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dists", default="uniform,gaussian")
ap.add_argument("--r-polys", type=int, default=5_000_000)
ap.add_argument("--s-segments", default="1000000,10000000,35000000,100000000")
ap.add_argument("--xsect-factors", default="0.1,0.2,0.5")
ap.add_argument("--modes", default="lbvh,grid")
ap.add_argument("--grid-size", type=int, default=8192)
ap.add_argument("--repeat", type=int, default=5)
ap.add_argument("--check-max", type=int, default=40_000_000, help="largest S (segments) checked against the oracle")
ap.add_argument("--ref", type=int, default=1)
ap.add_argument("--ref-max", type=int, default=40_000_000)
ap.add_argument("--shared", type=int, default=1, help="also run the shared-chain variant (at the 10 M point)")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
bench.pin_to_cores(local, world)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)


N_PIECES = 8  # S is always the same union of 8 seeded pieces, whatever the number of ranks


def concat(gs):
    rows, off = [np.zeros(1, np.int64)], 0
    for g in gs:
        rows.append(g.row_index.astype(np.int64)[1:] + off)
        off += g.n_points
    return RJ.PlanarGraph(np.vstack([g.xy for g in gs]), np.concatenate(rows).astype(np.uint32),
                          np.concatenate([g.left for g in gs]), np.concatenate([g.right for g in gs]))


def make_S(d, npoly, need_full):
    """-> (this rank's shard, its eid offset in the whole S, the whole S or None, total edges).
    Rank r owns pieces [r * 8 / N, (r + 1) * 8 / N): every rank generates only its own."""
    per = max(1, npoly // N_PIECES)
    mine = range(rank * N_PIECES // world, (rank + 1) * N_PIECES // world)
    pieces = {k: synth.polygon_soup(per, d, seed=100 + k) for k in (range(N_PIECES) if need_full and rank == 0 else mine)}
    sizes = torch.zeros(N_PIECES, dtype=torch.int64, device=dev)
    for k in mine:
        sizes[k] = pieces[k].n_edges
    if world > 1:
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    sizes = sizes.cpu().numpy()
    shard = concat([pieces[k] for k in mine])
    full = concat([pieces[k] for k in range(N_PIECES)]) if need_full and rank == 0 else None
    return shard, int(sizes[:mine[0]].sum()), full, int(sizes.sum())


def oracle_lsi(R, S, bbox):
    from oracle import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))
    sc = O.scaling_init(*bbox)
    r, s = O.scale_points(sc, R.xy), O.scale_points(sc, S.xy)
    rp1, _ = O.build_edges(R.row_index)
    sp1, _ = O.build_edges(S.row_index)
    t = time.perf_counter()
    want = O.lsi_grid(s, sp1, r, rp1, sc)
    return want, time.perf_counter() - t, O.num_threads()


def run_point(tag, R, shard, eid_off, S, n_S_edges, xfs, check, ref, bbox):
    """S: the whole query map on rank 0 when the result is to be checked, else None"""
    from helpers import sort_xsects
    stream = torch.cuda.Stream(device=dev)  # the engine works on this stream: the events below time it
    ctx = RJ.Context(device=local, stream=stream.cuda_stream)
    ctx.set_option("keep_host_graph", 0)
    ctx.set_bounding_box(*bbox)
    ctx.set_map(0, R)
    ctx.set_map(1, shard)
    want = None
    for mode in args.modes.split(","):
        build = min(ctx.build_index(0, mode, args.grid_size) for _ in range(2))
        for xf in xfs:
            lsi = RJ.LSI(ctx, mode)
            lsi.Init(xf)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                  for _ in range(args.repeat)]
            err = None
            n = 0
            try:
                lsi.Query(1)  # warm-up (orders the query map once, sizes the internal queues)
                lsi.Query(1)
                if world > 1:
                    dist.barrier()
                for a, b in ev:
                    a.record(stream)
                    lsi.Launch(1)
                    b.record(stream)
                    n = lsi.Wait()
                torch.cuda.synchronize()
                ms = min(a.elapsed_time(b) for a, b in ev)
            except RJ.RjbError as e:  # queue overflow at a small xsect_factor: reported, not fatal
                err, ms, n = "%s (needed %d)" % (str(e)[:80], getattr(e, "needed", -1)), float("nan"), 0
            t = torch.tensor([ms if err is None else 1e30], dtype=torch.float64, device=dev)
            c = torch.tensor([n, lsi.n_candidates if err is None else 0, 1 if err else 0], dtype=torch.float64,
                             device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(c, op=dist.ReduceOp.SUM)
            line = {"point": tag, "n_gpus": world, "mode": mode, "xsect_factor": xf, "R_edges": R.n_edges,
                    "S_edges": n_S_edges, "join_ms": t.item() if c[2].item() == 0 else None,
                    "build_ms": build, "pairs": int(c[0].item()), "candidates": int(c[1].item()),
                    "edges_per_s": n_S_edges / (t.item() / 1e3) if c[2].item() == 0 else None}
            if c[2].item():
                line["error"] = err or "queue overflow on another rank"
            elif check and xf == xfs[-1]:
                # device-side gather of the result records, checked on rank 0
                xs_local = rd.device_view(lsi._res[0], 32 * n, dev).clone() if n else \
                    torch.empty(0, dtype=torch.uint8, device=dev)
                if n:
                    xs_local.view(torch.int32).view(-1, 8)[:, 5] += eid_off  # eid[1] -> id in the unsharded S
                counts = rd.allgather_counts(dist, [n], dev)[:, 0] if world > 1 else np.array([n])
                allx = rd.gather_bytes(dist, xs_local, counts * 32) if world > 1 else xs_local
                if rank == 0:
                    got = sort_xsects(allx.cpu().numpy().view(RJ.XSECT_DTYPE), 1)
                    if want is None:
                        want = {}
                    if mode not in want:
                        if mode == "grid":
                            from oracle import oracle as O
                            sc = O.scaling_init(*bbox)
                            pr, ps = O.scale_points(sc, R.xy), O.scale_points(sc, S.xy)
                            want[mode] = (O.lsi_refgrid(pr, O.build_edges(R.row_index)[0], ps,
                                                        O.build_edges(S.row_index)[0], sc, args.grid_size), 0, 0)
                        else:
                            want[mode] = oracle_lsi(R, S, bbox)
                    w, t_cpu, cores = want[mode]
                    ok = len(got[0]) == len(w[0]) and all(np.array_equal(g, x) for g, x in zip(got, w))
                    line["parity_vs_oracle"] = "bit-exact" if ok else "MISMATCH (%d vs %d)" % (len(got[0]), len(w[0]))
                    if t_cpu:
                        line["oracle_ms"] = t_cpu * 1e3
                        line["oracle_cores"] = cores
            if rank == 0:
                print(json.dumps(line), flush=True)
    ctx.close()
    if ref and rank == 0 and world == 1:
        from tools import ref_runner
        if ref_runner.available():
            for rm in ("lbvh", "grid"):
                try:
                    out = ref_runner.run_lsi(None, R, S, mode=rm, warmup=2, repeat=3, xsect_factor=xfs[-1],
                                             grid_size=args.grid_size, workdir=bench.CACHE, timeout=600)
                    print(json.dumps({"point": tag, "impl": "reference", "mode": rm, "S_edges": n_S_edges,
                                      "pairs": out["intersections"], "query_ms": out["query_ms"],
                                      "build_ms": out["phases"]["build_index_ms"]}), flush=True)
                except Exception as e:
                    print(json.dumps({"point": tag, "impl": "reference", "mode": rm, "error": str(e)[:200]}), flush=True)
    if world > 1:
        dist.barrier()


xfs = [float(x) for x in args.xsect_factors.split(",")]
for d in args.dists.split(","):
    R = synth.polygon_soup(args.r_polys, d, seed=1)  # every rank needs all of R: generated in parallel
    # the affine image of the unit square (misc/gen_polys.sh: 50,0,-119,0,30,35) + the polygon radius:
    # the same Scaling on every rank, whatever part of S it holds
    bbox = (-119.01, 34.99, -68.99, 65.01)
    for nseg in [int(x) for x in args.s_segments.split(",")]:
        npoly = max(N_PIECES, nseg // 7)
        check = nseg <= args.check_max
        ref = bool(args.ref) and nseg <= args.ref_max and world == 1
        shard, eid_off, S, n_S = make_S(d, npoly, check or ref)
        run_point("%s/%dM" % (d, round(n_S / 1e6)), R, shard, eid_off, S, n_S, xfs, check, ref, bbox)
        if args.shared and 5_000_000 < n_S < 20_000_000 and d == "uniform" and world == 1:
            S2 = synth.share_chains(R, S, frac=0.05, seed=3)
            run_point("%s/%dM+shared" % (d, round(S2.n_edges / 1e6)), R, S2, 0, S2, S2.n_edges, xfs[-1:], True,
                      bool(args.ref), bbox)
if world > 1:
    dist.destroy_process_group()
