#!/usr/bin/env python
"""Headline benchmark: LSI join of County x Zipcode-scale synthetic maps
(BASELINE.json configs[1]: ~4M base edges R vs ~9M query edges S) on B200, with the
metric's other quantities -- PIP (configs[2]), polygon overlay, index build -- as extra
keys of the same JSON line.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun)
  python bench.py --impl reference ...                   reference arm (see run_reference)

One "step" = one LSI Query(): every S edge against the LBVH of R, including the exact
intersection points (the reference's "Query" phase, src/run_query.cu:297-303).  `value`
times it on the device (CUDA events around the enqueued query, rjb_lsi_launch /
rjb_lsi_wait) with S resident in HBM; `e2e` times the same query through the synchronous
C ABI from pinned HOST buffers: H2D of the S batch + scaling + query + D2H of the
rjb_xsect results, every step.

Multi-GPU: R and its LBVH are replicated, S is sharded.  Default = weak scaling (rank r
owns its own 9M-edge S, seed 2 + r); the line also carries a strong-scaling leg (ONE
9M-edge S cut by whole chains over the ranks, rayjoin_b200.dist.shard_graph).  NCCL
carries only the count all-gather (one per timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "synthetic County x Zipcode-scale LSI: |R|~4.0M edges (3100 faces), |S|~9.0M edges (33000 faces) per GPU, US bbox"
PIP_WORKLOAD = "PIP: 100M uniform random points vs BlockGroup-scale synthetic map (220k faces, ~28.0M edges), US bbox"
OVERLAY_WORKLOAD = "polygon overlay (polyover_exec protocol) of the County x Zipcode-scale pair, xsect_factor 0.5"
R_FACES, R_EDGES, S_FACES, S_EDGES = 3100, 4_000_000, 33_000, 9_000_000
PIP_FACES, PIP_EDGES, PIP_POINTS = 220_000, 28_000_000, 100_000_000
XSECT_FACTOR = 0.1  # expr/env.sh:16 of the reference
OVERLAY_XSECT_FACTOR = 0.5
CACHE = os.environ.get("RJB_CACHE", "/tmp/rjb200_cache")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_config(args, world):
    """The workload description -- identical keys and values on both arms."""
    legs = args.legs if args.legs else ("lsi,pip,overlay" if world == 1 else "lsi")
    cfg = {"workload": WORKLOAD, "xsect_factor": XSECT_FACTOR, "scaling": "weak", "legs": legs}
    if "pip" in legs:
        cfg["pip_workload"] = PIP_WORKLOAD
    if "overlay" in legs:
        cfg["overlay_workload"] = OVERLAY_WORKLOAD
    return cfg


def _cached_map(path, make):
    from rayjoin_b200.capi import PlanarGraph
    os.makedirs(CACHE, exist_ok=True)
    if os.path.exists(path):
        try:
            z = np.load(path)
            return PlanarGraph(z["xy"], z["row_index"], z["left"], z["right"])
        except Exception:
            pass
    g = make()
    try:
        tmp = path + ".%d.tmp.npz" % os.getpid()
        np.savez(tmp, xy=g.xy, row_index=g.row_index, left=g.left, right=g.right)
        os.replace(tmp, path)
    except Exception:
        pass
    return g


def get_map(kind, seed, scale=1.0):
    """Seeded synthetic map, cached as .npz for the second arm on the same box."""
    from rayjoin_b200 import synth
    faces, edges = (R_FACES, R_EDGES) if kind == "R" else (S_FACES, S_EDGES)
    faces, edges = max(8, int(faces * scale)), max(64, int(edges * scale))
    path = os.path.join(CACHE, "%s_%d_%d_%d.npz" % (kind, faces, edges, seed))
    return _cached_map(path, lambda: synth.voronoi_map(faces, edges, synth.US_BBOX, seed=seed))


def get_pip_map(scale=1.0):
    from rayjoin_b200 import synth
    faces, edges = max(8, int(PIP_FACES * scale)), max(64, int(PIP_EDGES * scale))
    path = os.path.join(CACHE, "pipmap_%d_%d.npz" % (faces, edges))
    g = _cached_map(path, lambda: synth.voronoi_map(faces, edges, synth.US_BBOX, seed=1))
    g.bbox = synth.US_BBOX
    return g


def pip_points_host(n, seed=1):
    """Uniform random query points inside the scaled US box (int64 internal coordinates)."""
    rng = np.random.default_rng(seed)
    lim = int(2**46 * 0.97)
    return np.column_stack([rng.integers(-lim, lim, n), rng.integers(-lim, lim, n)]).astype(np.int64)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML in-process,
    ~1 ms per sample; nvidia-smi as a fallback)."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
           0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self.stop_flag = index, [], [], set(), False
        self.power = []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
            try:
                self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            try:
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.BAD.items():
                if r & bit:
                    self.reasons.add(name)
        else:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index),
                                  "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            self.sm.append(float(out[0]))
            self.mx.append(float(out[1]))

    def run(self):
        while not self.stop_flag:
            try:
                self.sample()
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        self.stop_flag = True
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "power_w_max": max(self.power) if self.power else None,
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the
    committed `ncu --set full` capture of this workload (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return sum(t[k] for k in kernel.split("+"))
    except Exception:
        return None


def gpu_cpu_sets(n_gpus, allowed):
    """Per local GPU: the allowed host cores NVML names as close to it (its NUMA node), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = [int(v) for v in vis.split(",")] if vis and all(v.strip().isdigit() for v in vis.split(",")) else None
        words = (max(os.cpu_count() or 1, max(allowed) + 1) + 63) // 64
        out = []
        for i in range(n_gpus):
            h = pynvml.nvmlDeviceGetHandleByIndex(phys[i] if phys else i)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            near = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
            out.append(sorted(near & set(allowed)))
        return out
    except Exception:
        return None


def pin_to_cores(local_rank, world):
    """Each rank gets its own slice of the host cores: the completion wait polls the
    stream, and eight unpinned pollers next to each other showed up as single steps of
    1-2 ms (round 1, N = 8).  The slice is taken from the cores NEAR the rank's GPU (NVML CPU
    affinity = its NUMA node), shared evenly among the ranks whose GPUs sit on the same node, so
    that the pinned upload buffers -- allocated after this call -- are node-local to the GPU that
    reads them; without NVML: equal slices of the allowed cores in rank order."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        if local_world > 1:
            sets = gpu_cpu_sets(local_world, cpus)
            mine = sets[local_rank] if sets else None
            if mine and len(mine) < len(cpus):  # (one node for everything: nothing to choose)
                peers = [i for i in range(local_world) if sets[i] == mine]
                per = len(mine) // len(peers)
                if per >= 1:
                    k = peers.index(local_rank)
                    os.sched_setaffinity(0, mine[k * per:(k + 1) * per])
                    return len(os.sched_getaffinity(0))
            per = len(cpus) // max(1, local_world)
            if per >= 1:
                os.sched_setaffinity(0, cpus[local_rank * per:(local_rank + 1) * per])
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def cpu_oracle_lsi(R, S, bbox, repeats=1):
    """The multithreaded host exact-predicate oracle on the same workload."""
    from oracle import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))  # torch.set_num_threads(1) above lowered it
    sc = O.scaling_init(*bbox)
    r, s = O.scale_points(sc, R.xy), O.scale_points(sc, S.xy)
    rp1, _ = O.build_edges(R.row_index)
    sp1, _ = O.build_edges(S.row_index)
    best, res = None, None
    for _ in range(repeats):
        t = time.perf_counter()
        res = O.lsi_grid(s, sp1, r, rp1, sc, return_candidates=True)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best, len(res[0]), res[4], O.num_threads(), res


# ---------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------
def run_reference(args, rank, world, json_out):
    """Reference arm.  RayJoin has no CPU implementation of this path; its only code that
    can run on a B200 box are the -mode=lbvh / -mode=grid CUDA backends, built unmodified
    from /root/reference with stub OptiX/glog headers into oracle/_ref/ref_exec
    (oracle/Makefile).  When that binary is present it is timed (GPU, the reference's own
    phase timers: Query = wall clock over `repeat` iterations); otherwise the host oracle
    port is timed on all host cores.  At N = 1 the line also carries the reference's PIP
    (bounded samples: its lbvh PIP visits every leaf above the point) and overlay phases."""
    if rank != 0:
        return
    from rayjoin_b200 import synth
    from rayjoin_b200.capi import PlanarGraph
    cfg = make_config(args, world)
    R, S = get_map("R", 1, args.scale), get_map("S", 2, args.scale)
    bbox = synth.union_bbox(R, S)
    line = {"impl": "reference", "metric": "LSI join throughput", "unit": "query_edges/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/int128",
            "data": "synthetic", "config": cfg}
    ref_exec = os.path.join(ROOT, "oracle", "_ref", "ref_exec")
    used = None
    have_ref = os.path.exists(ref_exec) and args.ref_mode != "cpu"
    if have_ref:
        try:
            from tools import ref_runner
            out = ref_runner.run_lsi(ref_exec, R, S, mode=args.ref_mode, warmup=args.warmup,
                                     repeat=args.steps, xsect_factor=max(XSECT_FACTOR, 0.1),
                                     workdir=CACHE)
            ms = out["query_ms"]
            line.update({"value": S.n_edges / (ms / 1e3), "ms_per_step": ms, "join_ms": ms,
                         "result_pairs": out.get("intersections"),
                         "index_build_ms": out["phases"].get("build_index_ms"),
                         "reference_phases_ms": out.get("phases"),
                         "cpu_baseline": {"value": S.n_edges / (ms / 1e3), "unit": "query_edges/s",
                                          "cores": 0, "kind": "reference",
                                          "sample": "full workload on the GPU through the reference's own "
                                                    "-mode=%s CUDA backend (it has no CPU path); "
                                                    "its 'Query' phase timer, warmup=%d repeat=%d"
                                                    % (args.ref_mode, args.warmup, args.steps)}})
            used = "ref_exec"
        except Exception as e:  # fall through to the port
            log("reference binary failed (%s); timing the oracle port instead" % e)
    if used is None:
        dt, n, cand, cores, _ = cpu_oracle_lsi(R, S, bbox, repeats=max(1, min(args.steps, 3)))
        line.update({"value": S.n_edges / dt, "ms_per_step": dt * 1e3, "join_ms": dt * 1e3, "result_pairs": n,
                     "candidate_pairs": cand,
                     "cpu_baseline": {"value": S.n_edges / dt, "unit": "query_edges/s", "cores": cores,
                                      "kind": "port",
                                      "sample": "full workload (%d x %d edges), best of %d runs, "
                                                "grid filter + exact predicate, OpenMP"
                                                % (R.n_edges, S.n_edges, max(1, min(args.steps, 3)))}})
    line["e2e"] = {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0}
    if used == "ref_exec":
        from tools import ref_runner
        if "lsi" in cfg["legs"] and world == 1:
            try:  # the reference's other backend on the same input
                other = "grid" if args.ref_mode == "lbvh" else "lbvh"
                out = ref_runner.run_lsi(ref_exec, R, S, mode=other, warmup=2, repeat=3,
                                         xsect_factor=max(XSECT_FACTOR, 0.1), grid_size=2048, workdir=CACHE)
                line["lsi_%s" % other] = {"query_ms": out["query_ms"], "build_index_ms": out["phases"]["build_index_ms"],
                                          "result_pairs": out.get("intersections"), "grid_size": 2048}
            except Exception as e:
                line["lsi_other_error"] = str(e)[:200]
        if "pip" in cfg["legs"]:
            try:
                M = get_pip_map(args.scale)
                pip = {"workload": PIP_WORKLOAD, "edges": M.n_edges}
                x0, y0, x1, y1 = synth.US_BBOX
                for mode, n in (("grid", int(10_000_000 * args.scale)), ("lbvh", int(1_000_000 * args.scale))):
                    n = max(2, n // 2 * 2)
                    rng = np.random.default_rng(1)
                    xy = np.column_stack([rng.uniform(x0, x1, n), rng.uniform(y0, y1, n)])
                    Q = PlanarGraph(xy, np.arange(0, n + 1, 2, dtype=np.uint32), np.ones(n // 2, np.int64),
                                    np.full(n // 2, 2, np.int64), bbox=synth.US_BBOX)
                    out = ref_runner.run_pip(ref_exec, M, Q, mode=mode, warmup=1, repeat=2, grid_size=4096,
                                             workdir=CACHE)
                    pip[mode] = {"points": n, "query_ms": out["query_ms"],
                                 "points_per_s": n / (out["query_ms"] / 1e3),
                                 "build_index_ms": out["phases"]["build_index_ms"],
                                 "sample": "bounded: %d of the 100M points (vertices of a 2-point-chain map)" % n}
                line["pip"] = pip
            except Exception as e:
                line["pip"] = {"error": str(e)[:300]}
        if "overlay" in cfg["legs"]:
            ov = {"workload": OVERLAY_WORKLOAD}
            for mode in ("lbvh", "grid"):
                try:
                    out = ref_runner.run_overlay(ref_exec, R, S, mode=mode, xsect_factor=OVERLAY_XSECT_FACTOR,
                                                 grid_size=2048, workdir=CACHE)
                    ph = out["phases"]
                    ov[mode] = {"phase_ms": ph, "intersections": out["intersections"],
                                "device_total_ms": sum(ph[k] for k in ("build_index_ms", "lsi_ms", "pip0_ms",
                                                                       "pip1_ms", "polygons_ms"))}
                except Exception as e:
                    ov[mode] = {"error": str(e)[:300]}
            line["overlay"] = ov
    print(json.dumps(line), file=json_out, flush=True)


# ---------------------------------------------------------------------------------
# extra legs of the repo arm (N = 1)
# ---------------------------------------------------------------------------------
def pip_leg(args, RJ, torch, dev, stream, peak):
    """BASELINE.json configs[2]: 100M uniform random points against the 28M-edge map."""
    from oracle import oracle as O
    M = get_pip_map(args.scale)
    n = max(1024, int(PIP_POINTS * args.scale))
    ctx = RJ.Context(device=dev.index, stream=stream.cuda_stream)
    out = {"workload": PIP_WORKLOAD, "points": n, "edges": M.n_edges, "chains": M.n_chains}
    try:
        ctx.set_option("keep_host_graph", 0)
        ctx.set_bounding_box(*M.bbox)
        ctx.set_map(0, M)
        h_pts = torch.from_numpy(pip_points_host(n)).pin_memory()
        with torch.cuda.stream(stream):
            d_pts = h_pts.to(dev, non_blocking=True)
        stream.synchronize()
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for mode in args.pip_modes.split(","):
            res = {}
            build = [ctx.build_index(0, mode, args.pip_grid_size) for _ in range(3)]
            res["index_build_ms"] = float(np.min(build))
            idx = ctx.index_info(0, mode)
            res["index_bytes"] = int(idx["bytes"])
            ctx.set_option("sort_queries", 1)
            steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            kms = []
            de = df = None
            for i in range(warm + steps):
                with torch.cuda.stream(stream):
                    flush_buf.zero_()
                    if i >= warm:
                        ev[i - warm][0].record(stream)
                    de, df, cand = ctx.pip_device(1, mode, d_pts.data_ptr(), n)
                    if i >= warm:
                        ev[i - warm][1].record(stream)
                if i >= warm:
                    kms.append(ctx.last_stage_ms()[0][:2])
            stream.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
            kms = np.mean(np.asarray(kms), axis=0)
            res.update({"query_ms": ms, "points_per_s": n / (ms / 1e3), "order_ms": float(kms[0]),
                        "kernel_ms": float(kms[1]), "candidates": int(cand), "gpu_launches": ctx.last_launches()})
            # compulsory traffic of the query with this layout: 16 B per point in, eid + face id out,
            # the index and the base vertices once (SURVEY 8(d) pip_bvh_traverse / pip_grid)
            alg = 16 * n + 8 * n + idx["bytes"] + 16 * M.n_points
            res["roofline"] = {"bound": "hbm", "algorithmic_bytes": int(alg), "achieved": alg / (ms / 1e3) / 1e9,
                               "peak": peak, "unit": "GB/s", "frac": alg / (ms / 1e3) / 1e9 / peak,
                               "traffic": ncu_traffic("k_pip_%s" % mode)}
            # every point against the oracle (host grid filter + the restated update rule)
            if not args.no_check:
                ncheck = n if args.pip_check < 0 else min(n, args.pip_check)
                eids = ctx.copy_to_host(de, np.empty(n, np.uint32))[:ncheck]
                sc = O.scaling_init(*M.bbox)
                O.set_num_threads(len(os.sched_getaffinity(0)))
                t = time.perf_counter()
                want = O.pip_grid(O.scale_points(sc, M.xy), O.build_edges(M.row_index)[0], sc,
                                  h_pts.numpy()[:ncheck], 1)
                res["oracle_s"] = time.perf_counter() - t
                bad = int((eids != want).sum())
                res["parity_vs_oracle"] = ("bit-exact (%d points checked)" % ncheck) if bad == 0 else \
                    "MISMATCH (%d of %d)" % (bad, ncheck)
                res["hit_fraction"] = float((eids != 0xFFFFFFFF).mean())
            out[mode] = res
        # end to end: scaled int64 points from pinned host memory, eids + face ids back
        mode = args.pip_modes.split(",")[0]
        h_e = torch.empty(n, dtype=torch.int32).pin_memory()
        h_f = torch.empty(n, dtype=torch.int32).pin_memory()
        ctx.build_index(0, mode, args.pip_grid_size)
        ts = []
        for _ in range(3):
            t = time.perf_counter()
            ctx.pip_host_scaled(1, mode, h_pts.data_ptr(), n, h_e.data_ptr(), h_f.data_ptr())
            ts.append(time.perf_counter() - t)
        out["e2e"] = {"mode": mode, "ms": float(np.min(ts[1:])) * 1e3,
                      "points_per_s": n / float(np.min(ts[1:])), "h2d_bytes": 16 * n, "d2h_bytes": 8 * n}
        out["best_mode"] = min((m for m in args.pip_modes.split(",")), key=lambda m: out[m]["query_ms"])
        out["query_ms"] = out[out["best_mode"]]["query_ms"]
        out["points_per_s"] = out[out["best_mode"]]["points_per_s"]
    finally:
        ctx.close()
    return out


def overlay_leg(args, RJ, R, S, dev, stream):
    ctx = RJ.Context([R, S], device=dev.index, stream=stream.cuda_stream)
    out = {"workload": OVERLAY_WORKLOAD}
    try:
        for mode in ("lbvh", "grid"):
            ov = RJ.MapOverlay(ctx, mode, grid_size=args.grid_size, xsect_factor=OVERLAY_XSECT_FACTOR)
            best = None
            for _ in range(3):
                ms = ov.Run()
                if best is None or ms["total"] < best["total"]:
                    best = dict(ms)
            out[mode] = {"phase_ms": best, "device_total_ms": best["total"],
                         "intersections": int(len(ov.get_xsect_edges(0)))}
    finally:
        ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="rjb200", choices=["rjb200", "reference"])
    ap.add_argument("--ref-mode", default="lbvh", choices=["lbvh", "grid", "cpu"])
    ap.add_argument("--mode", default="lbvh", choices=["lbvh", "grid"])
    ap.add_argument("--grid-size", type=int, default=8192)
    ap.add_argument("--leaf-size", type=int, default=4)
    ap.add_argument("--sort-queries", type=int, default=-1)
    ap.add_argument("--filter", type=int, default=-1, help="occupancy pre-filter: -1 auto, 0 off, 1 on")
    ap.add_argument("--cells", type=int, default=1, help="cell-directory candidate path for the filter's survivors (option lsi_cells)")
    ap.add_argument("--tile-filter", type=int, default=1, help="two-level occupancy filter (option lsi_tile_filter)")
    ap.add_argument("--fused", type=int, default=1, help="exact + point pass as one kernel (option lsi_fused)")
    ap.add_argument("--stage-timing", type=int, default=-1, help="engine-internal CUDA events during the headline loop: -1 none (default), 0 per phase, 1 per kernel; the per-kernel times always come from a separate pass")
    ap.add_argument("--ag", type=int, default=0, help="adaptive leaf grouping (option lbvh_ag)")
    ap.add_argument("--ag-iter", type=int, default=5)
    ap.add_argument("--enlarge", type=float, default=3.5)
    ap.add_argument("--share-chains", type=float, default=0.0,
                    help="append copies of this fraction of R's chains (every 2nd vertex) to S: shared "
                         "vertices and near-collinear edges (experiments only)")
    ap.add_argument("--legs", default="", help="lsi,pip,overlay (default: all at N = 1, lsi at N > 1)")
    ap.add_argument("--pip-modes", default="lbvh,grid")
    ap.add_argument("--pip-grid-size", type=int, default=16384)
    ap.add_argument("--pip-check", type=int, default=10_000_000, help="points checked against the oracle (-1: all)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg at N > 1")
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: everything libraries print there (NCCL's
    # version banner, ...) is sent to stderr, the line goes to the saved descriptor
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank, world, json_out)
        return

    n_cores = pin_to_cores(local_rank, world)
    import torch
    import torch.distributed as dist
    import rayjoin_b200 as RJ
    from rayjoin_b200 import dist as rdist
    from rayjoin_b200 import synth

    RJ.load_library()  # fail loudly when the CUDA extension is missing
    torch.set_num_threads(1)  # N ranks share the host cores; the host side is launch-bound
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = make_config(args, world)
    legs = cfg["legs"].split(",")

    t0 = time.time()
    R = get_map("R", 1, args.scale)
    S = get_map("S", 2 + rank, args.scale)
    if args.share_chains > 0:
        S = synth.share_chains(R, S, frac=args.share_chains, seed=3)
    bbox = synth.US_BBOX  # same scaling on every rank
    log("[rank %d] maps ready in %.1fs: R %d edges / %d chains, S %d edges / %d chains; host cores of this rank: %s %s"
        % (rank, time.time() - t0, R.n_edges, R.n_chains, S.n_edges, S.n_chains, n_cores,
           sorted(os.sched_getaffinity(0)) if world > 1 else ""))

    stream = torch.cuda.Stream(device=dev)
    ctx = RJ.Context(device=local_rank, stream=stream.cuda_stream)
    ctx.set_option("keep_host_graph", 0)
    ctx.set_option("lbvh_leaf_size", args.leaf_size)
    ctx.set_option("sort_queries", args.sort_queries)
    ctx.set_option("lsi_filter", args.filter)
    ctx.set_option("lsi_cells", args.cells)
    ctx.set_option("lsi_tile_filter", args.tile_filter)
    ctx.set_option("lsi_fused", args.fused)
    ctx.set_option("stage_timing", args.stage_timing)
    ctx.set_option("lbvh_ag", args.ag)
    ctx.set_option("lbvh_ag_iter", args.ag_iter)
    ctx.set_option("lbvh_enlarge_x1000", int(args.enlarge * 1000))
    ctx.set_bounding_box(*bbox)
    ctx.set_map(0, R)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    def host_buffers(g):
        # pinned host copies of a query map for the end-to-end leg
        return g, (pin(g.xy), pin(g.row_index), pin(g.left), pin(g.right))

    def upload(hb):
        g, (xy, row, left, right) = hb
        ctx.set_map_raw(1, xy.data_ptr(), g.n_points, row.data_ptr(), left.data_ptr(), right.data_ptr(),
                        g.n_chains)

    hb_S = host_buffers(S)
    upload(hb_S)
    build_ms = [ctx.build_index(0, args.mode, args.grid_size) for _ in range(3)]
    idx = ctx.index_info(0, args.mode)
    lsi = RJ.LSI(ctx, args.mode)
    lsi.Init(XSECT_FACTOR)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # the only data-path collective: the per-rank {result, candidate} counts of the steps of a
    # timed region, ONE NCCL all-gather of 16 B per step and rank after the last step (nothing
    # in a query depends on another rank's counts; a small all-gather costs 0.1-0.2 ms of latency,
    # more than a whole step, so it is not issued per step).  Fixed shape: NCCL sets its
    # connections up lazily on the first use of a shape, which the warm-up absorbs.
    n_rows = max(args.steps, args.warmup, 1)
    counts_host = torch.zeros((n_rows, 2), dtype=torch.int64).pin_memory()
    counts_dev = torch.zeros_like(counts_host, device=dev)
    counts_all = torch.zeros((world, n_rows, 2), dtype=torch.int64, device=dev)

    def exchange(rows):
        if world == 1:
            return
        counts_host.zero_()
        counts_host[:len(rows)] = torch.tensor(rows[-n_rows:], dtype=torch.int64)
        with torch.cuda.stream(stream):
            counts_dev.copy_(counts_host, non_blocking=True)
            dist.all_gather_into_tensor(counts_all, counts_dev)

    def kernel_times(n_steps):
        """Per-kernel CUDA-event times from a SEPARATE pass with the engine's own stage events on
        (option stage_timing = 1: an event after every kernel adds 2-3 us of stream time each, 13 us
        per query, so the headline loop below runs without them)."""
        ctx.set_option("stage_timing", 1)
        stage = []
        for i in range(n_steps + 2):
            with torch.cuda.stream(stream):
                flush_buf.zero_()
                lsi.Launch(1)
            lsi.Wait()
            if i >= 2:
                stage.append(ctx.last_stage_ms()[0])
        ctx.set_option("stage_timing", args.stage_timing)
        return np.mean(np.asarray(stage), axis=0)

    def timed_steps(n_steps, n_warm):
        """-> (per-step device ms, ms of the count exchange, per-kernel ms, pairs, cand)"""
        rows = []
        for i in range(n_warm):
            lsi.Launch(1)
            rows.append((lsi.Wait(), lsi.n_candidates))
            if i < 2:
                exchange(rows)  # twice: connection set-up, then the steady state
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(n_steps)]
        xch = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        stage, rows = [], []
        n = 0
        gc.disable()  # a collection inside a 0.16 ms step shows up as a 5x outlier
        for i in range(n_steps):
            with torch.cuda.stream(stream):
                flush_buf.zero_()  # evict L2 between timed iterations (outside the events)
                ev[i][0].record(stream)
                lsi.Launch(1)     # the whole query is enqueued ...
                ev[i][1].record(stream)
            n = lsi.Wait()        # ... and completed; the events bracket its device time
            rows.append((n, lsi.n_candidates))
        with torch.cuda.stream(stream):
            xch[0].record(stream)
            exchange(rows)
            xch[1].record(stream)
        torch.cuda.synchronize()
        gc.enable()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        drain = xch[0].elapsed_time(xch[1]) if world > 1 else 0.0
        return step_ms, drain, kernel_times(min(10, n_steps)), n, lsi.n_candidates

    sampler = ClockSampler(local_rank)
    if rank == 0:  # one sampler per job: NVML calls from N processes serialise in the driver
        sampler.start()
    step_ms, drain_ms, stage_ms, n_pairs, n_cand = timed_steps(args.steps, args.warmup)
    total_ms = sum(step_ms) + drain_ms
    launches = ctx.last_launches()
    lst = ctx.last_stats()
    log("[rank %d] step ms min/median/max %.4f/%.4f/%.4f, count exchange %.4f ms"
        % (rank, min(step_ms), float(np.median(step_ms)), max(step_ms), drain_ms))

    # ---- end to end through the synchronous C ABI from pinned host buffers ---------
    out_host = torch.empty(max(1, n_pairs) * 32, dtype=torch.uint8).pin_memory()
    out_np = out_host.numpy().view(RJ.XSECT_DTYPE)

    def e2e_run(hb, n_steps):
        def e2e_step():
            upload(hb)                                   # H2D + scale + edge numbering
            n = lsi.Query(1)                             # traversal + intersection points
            ctx.copy_to_host(lsi._res[0], out_np[:n])    # D2H of the rjb_xsect records
            return n
        for _ in range(min(3, args.warmup)):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        for _ in range(n_steps):
            n = e2e_step()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) * 1e3, n
    e2e_total_ms, _ = e2e_run(hb_S, args.steps)
    weak_result = out_np[:n_pairs].copy()  # checked against the oracle below
    clocks = sampler.summary()

    times = torch.tensor([total_ms, e2e_total_ms, max(step_ms), float(np.median(step_ms))], dtype=torch.float64, device=dev)
    sizes = torch.tensor([S.n_edges, n_pairs, n_cand], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, worst_step, worst_median = times.tolist()
    all_edges, all_pairs, all_cand = sizes.tolist()
    h2d = S.xy.nbytes + S.row_index.nbytes + 4 * 2 * S.n_chains
    d2h = n_pairs * 32 + 16

    # ---- strong scaling: ONE S map cut by whole chains over the ranks ------------------
    strong = None
    if world > 1 and not args.no_strong:
        S_full = get_map("S", 2, args.scale)
        shard, _, _ = rdist.shard_graph(S_full, rank, world)
        hb_sh = host_buffers(shard)
        upload(hb_sh)
        s_ms, s_drain, _, s_pairs, _ = timed_steps(args.steps, args.warmup)
        s_e2e, _ = e2e_run(hb_sh, args.steps)
        t = torch.tensor([sum(s_ms) + s_drain, s_e2e], dtype=torch.float64, device=dev)
        z = torch.tensor([s_pairs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(z, op=dist.ReduceOp.SUM)
        strong = {"workload": "ONE |S|=%d-edge map sharded by whole chains over %d ranks, R replicated"
                              % (S_full.n_edges, world),
                  "join_ms": t[0].item() / args.steps, "value": S_full.n_edges / (t[0].item() / args.steps / 1e3),
                  "unit": "query_edges/s", "e2e_ms_per_step": t[1].item() / args.steps,
                  "e2e_value": S_full.n_edges / (t[1].item() / args.steps / 1e3),
                  "result_pairs": int(z[0].item()), "h2d_bytes_per_step_per_rank": int(shard.xy.nbytes)}
        upload(hb_S)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = all_edges / (ms_per_step / 1e3)
        peak, peak_src = measured_peak()
        f_ms, t_ms, x_ms, p_ms = [float(v) for v in stage_ms]
        surv, ql_pairs = (lst[7] or S.n_edges), lst[2]
        n_leaves = idx["units"]
        if args.mode == "lbvh":
            # Algorithmic (compulsory) bytes per kernel with this layout.
            #  filter   : 4 B descriptor per S point, both 2 MiB bitmaps, 4 B per survivor out
            #  traversal: 4 B slot + one 16 B vertex per survivor (the second is the neighbour's),
            #             the tree once (node boxes + children + top tree), 8 B per (query, leaf) pair out
            #  exact    : the pair list, per distinct leaf its record and <= 5 vertices, per distinct
            #             query edge its two vertices, 8 B per hit out
            #  points   : per hit 8 B in, four vertices, two chain ids, 32 B record out
            n_maps = 4  # occupancy bitmaps: occ and its 2 x 2 / 4 x 4 / 8 x 8 dilations (2 MiB each)
            cells_on = bool(lst[5]) and args.cells > 0
            # (the cell directory = what k_lsi_cells reads instead of the tree: {word, rank} pairs,
            # list bounds per occupied cell, 16-byte item records)
            dir_bytes = idx.get("cell_directory_bytes", 0)
            tree_bytes = idx["bytes"] - n_maps * (4096 * 4096 // 8) - 8 * n_leaves - dir_bytes
            pair_in = 8 * ql_pairs
            # exact pass: per distinct leaf <= 5 vertices (+ its 8-byte record unless the pairs are
            # direct), per distinct query edge two vertices; point pass: four vertices, two chain
            # ids, the 32-byte record
            exact_b = pair_in + (80 if cells_on else 88) * min(ql_pairs, n_leaves) + 32 * min(ql_pairs, surv)
            points_b = (64 + 8 + 32) * n_pairs
            kern = [
                ("k_lsi_filter_tiles" if args.tile_filter else "k_lsi_filter", f_ms,
                 4 * S.n_points + 2 * (4096 * 4096 // 8) + 4 * surv),
                ("k_lsi_cells" if cells_on else "k_lsi_bvh", t_ms,
                 20 * surv + (dir_bytes if cells_on else tree_bytes) + 8 * ql_pairs),
            ]
            if args.fused and p_ms < 0.005:  # k_lsi_resolve = exact + point pass in one kernel
                kern.append(("k_lsi_resolve", x_ms + p_ms, exact_b + points_b))
            else:
                kern += [("k_lsi_exact", x_ms, exact_b + 8 * n_pairs), ("k_lsi_points", p_ms, 8 * n_pairs + points_b)]
            if not lst[7]:
                kern = kern[1:]
                kern[0] = ("k_lsi_bvh", t_ms, 4 * S.n_edges + 16 * S.n_points + tree_bytes + 8 * ql_pairs)
            # whole query, SURVEY 8(d): every S vertex, every S edge id, the index and every R vertex once
            # ("the index" = what this query path reads of it: the two bitmaps of the filter and the
            # cell directory, or the tree -- not both, and not the dilated bitmaps it does not use)
            idx_read = 2 * (4096 * 4096 // 8) + (dir_bytes if cells_on else tree_bytes + 8 * n_leaves)
            alg_query = 16 * S.n_points + 4 * S.n_edges + idx_read + 16 * R.n_points + 8 * n_pairs
        else:
            kern = [("k_lsi_grid", f_ms, 16 * S.n_points + 4 * S.n_edges + idx["bytes"] + 16 * R.n_points + 8 * n_pairs),
                    ("k_xsect_points_dyn", t_ms, (8 + 64 + 8 + 32) * n_pairs)]
            alg_query = kern[0][2]
        rk = []
        for name, ms, b in kern:
            rk.append({"kernel": name, "ms": ms, "algorithmic_bytes": int(b),
                       "achieved": (b / (ms / 1e3) / 1e9) if ms > 0 else None,
                       "frac": (b / (ms / 1e3) / 1e9 / peak) if ms > 0 else None,
                       "traffic": ncu_traffic(name)})
        dom = max(rk, key=lambda r: r["ms"])
        line = {
            "metric": "LSI join throughput", "value": value, "unit": "query_edges/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64/int128", "data": "synthetic",
            "config": cfg,
            "engine": {"mode": args.mode, "lbvh_leaf_size": args.leaf_size, "sort_queries": args.sort_queries,
                       "lbvh_ag": args.ag, "share_chains": args.share_chains, "lsi_cells": args.cells,
                       "lsi_fused": args.fused, "lsi_tile_filter": args.tile_filter,
                       "timing": "CUDA events on the launch stream around the enqueued query (rjb_lsi_launch); "
                                 "max over ranks of the sum of the K step times + the count exchange; "
                                 "kernel_ms / roofline_kernels: a separate pass with an event after every "
                                 "kernel (those events cost 13 us per query and are off in the headline loop)",
                       "l2": "256 MiB flush write between timed iterations; inputs (S descriptors + survivors' "
                             "vertices + index, ~150 MB) also exceed L2",
                       "sharding": "R + index replicated, S sharded per rank (seed 2+rank); NCCL: one all-gather "
                                   "of the per-step counts (16 B per step and rank) per timed region, its "
                                   "CUDA-event time added to the K step times",
                       "host_cores_per_rank": n_cores},
            "join_ms": ms_per_step, "result_pairs": int(all_pairs),
            "candidate_pairs": int(all_cand),
            "candidate_pairs_per_s": all_cand / (ms_per_step / 1e3),
            "index_build_ms": float(np.min(build_ms)), "index_bytes": int(idx["bytes"]),
            "index_units": int(idx["units"]),
            "kernel_ms": {r["kernel"]: r["ms"] for r in rk},
            "step_ms": {"median_worst_rank": worst_median, "max_worst_rank": worst_step,
                        "count_exchange": drain_ms},
            "e2e": {"value": all_edges / (e2e_ms / 1e3 / args.steps), "unit": "query_edges/s",
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak,
                         "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic"],
                         "algorithmic_bytes": dom["algorithmic_bytes"], "peak_source": peak_src,
                         "note": "dominant kernel of the query: its own algorithmic bytes over its own "
                                 "CUDA-event time; latency bound by design (a chain of dependent loads per "
                                 "warp), see roofline_kernels / roofline_query"},
            "roofline_kernels": rk,
            "roofline_query": {"algorithmic_bytes": int(alg_query), "ms": ms_per_step,
                               "achieved": alg_query / (ms_per_step / 1e3) / 1e9,
                               "frac": alg_query / (ms_per_step / 1e3) / 1e9 / peak,
                               "formula": "16*Np(S) + 4*Ne(S) + index + 16*Np(R) + 8*K over ms_per_step (SURVEY 8d)"},
        }
        if strong:
            line["strong_scaling"] = strong
        # check the last result against the host oracle, and time it: the CPU baseline
        if not args.no_cpu_baseline:
            dt, n_ref, cand_ref, cores, res = cpu_oracle_lsi(R, S, bbox)
            line["cpu_baseline"] = {"value": S.n_edges / dt, "unit": "query_edges/s", "cores": cores,
                                    "kind": "port", "ms": dt * 1e3,
                                    "sample": "full rank-0 workload (%d x %d edges), 1 run, host grid "
                                              "filter + exact predicate, OpenMP" % (R.n_edges, S.n_edges)}
            if not args.no_check:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from helpers import sort_xsects
                got = sort_xsects(weak_result, 1)
                ok = n_pairs == n_ref and all(np.array_equal(g, w) for g, w in zip(got, res[:4]))
                line["parity_vs_oracle"] = "bit-exact" if ok else "MISMATCH"
                if not ok:
                    log("PARITY MISMATCH against the oracle: %d vs %d pairs" % (n_pairs, n_ref))
            del res
    ctx.close()
    if rank == 0 and world == 1:
        del flush_buf
        torch.cuda.empty_cache()
        if "pip" in legs:
            try:
                line["pip"] = pip_leg(args, RJ, torch, dev, stream, peak)
            except Exception as e:  # the headline line must survive a failing extra leg
                line["pip"] = {"error": repr(e)[:300]}
        if "overlay" in legs:
            try:
                line["overlay"] = overlay_leg(args, RJ, R, get_map("S", 2, args.scale), dev, stream)
            except Exception as e:
                line["overlay"] = {"error": repr(e)[:300]}
    if rank == 0:
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
