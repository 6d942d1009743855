#!/usr/bin/env python
"""Headline benchmark: LSI join of County x Zipcode-scale synthetic maps
(BASELINE.json configs[1]: ~4M base edges R vs ~9M query edges S) on B200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun)
  python bench.py --impl reference ...                   reference arm (see below)

One "step" = one LSI Query(): every S edge against the LBVH of R, including
the exact intersection points (the reference's "Query" phase,
src/run_query.cu:297-303).  `value` times it with S resident in HBM; `e2e`
times the same query through the C ABI from pinned HOST buffers: H2D of the S
batch + scaling + query + D2H of the rjb_xsect results, every step.

Multi-GPU: R and its LBVH are replicated, S is sharded (rank r owns its own
9M-edge shard, seed 2 + r: weak scaling); NCCL carries only the per-step count
all-gather.  Prints ONE JSON line on rank 0.
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
R_FACES, R_EDGES, S_FACES, S_EDGES = 3100, 4_000_000, 33_000, 9_000_000
XSECT_FACTOR = 0.1  # expr/env.sh:16 of the reference
CACHE = os.environ.get("RJB_CACHE", "/tmp/rjb200_cache")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def get_map(kind, seed, scale=1.0):
    """Seeded synthetic map, cached as .npz for the second arm on the same box."""
    from rayjoin_b200 import synth
    from rayjoin_b200.capi import PlanarGraph
    faces, edges = (R_FACES, R_EDGES) if kind == "R" else (S_FACES, S_EDGES)
    faces, edges = max(8, int(faces * scale)), max(64, int(edges * scale))
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, "%s_%d_%d_%d.npz" % (kind, faces, edges, seed))
    if os.path.exists(path):
        try:
            z = np.load(path)
            return PlanarGraph(z["xy"], z["row_index"], z["left"], z["right"])
        except Exception:
            pass
    g = synth.voronoi_map(faces, edges, synth.US_BBOX, seed=seed)
    try:
        tmp = path + ".%d.tmp.npz" % os.getpid()
        np.savez(tmp, xy=g.xy, row_index=g.row_index, left=g.left, right=g.right)
        os.replace(tmp, path)
    except Exception:
        pass
    return g


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


def run_reference(args, rank, world, json_out):
    """Reference arm.  RayJoin has no CPU implementation of this path; its only
    code that can run on a B200 box are the -mode=lbvh / -mode=grid CUDA
    backends, built unmodified from /root/reference with stub OptiX/glog headers
    into oracle/_ref/ref_exec (oracle/Makefile).  When that binary is present it
    is timed (GPU, the reference's own 'Query' phase timer); otherwise the host
    oracle port is timed on all host cores."""
    if rank != 0:
        return
    from rayjoin_b200 import synth
    R, S = get_map("R", 1, args.scale), get_map("S", 2, args.scale)
    bbox = synth.union_bbox(R, S)
    line = {"impl": "reference", "metric": "LSI join throughput", "unit": "query_edges/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/int128",
            "data": "synthetic", "config": {"workload": WORKLOAD, "xsect_factor": XSECT_FACTOR}}
    ref_exec = os.path.join(ROOT, "oracle", "_ref", "ref_exec")
    used = None
    if os.path.exists(ref_exec) and args.ref_mode != "cpu":
        try:
            from tools import ref_runner
            out = ref_runner.run_lsi(ref_exec, R, S, mode=args.ref_mode, warmup=args.warmup,
                                     repeat=args.steps, xsect_factor=max(XSECT_FACTOR, 0.1),
                                     workdir=CACHE)
            ms = out["query_ms"]
            line.update({"value": S.n_edges / (ms / 1e3), "ms_per_step": ms,
                         "result_pairs": out.get("intersections"),
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
        line.update({"value": S.n_edges / dt, "ms_per_step": dt * 1e3, "result_pairs": n,
                     "candidate_pairs": cand,
                     "cpu_baseline": {"value": S.n_edges / dt, "unit": "query_edges/s", "cores": cores,
                                      "kind": "port",
                                      "sample": "full workload (%d x %d edges), best of %d runs, "
                                                "grid filter + exact predicate, OpenMP"
                                                % (R.n_edges, S.n_edges, max(1, min(args.steps, 3)))}})
    line["e2e"] = {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0}
    print(json.dumps(line), file=json_out, flush=True)


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
    ap.add_argument("--cells", type=int, default=0, help="experimental cell-directory candidate path (option lsi_cells)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
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

    import torch
    import torch.distributed as dist
    import rayjoin_b200 as RJ
    from rayjoin_b200 import synth

    RJ.load_library()  # fail loudly when the CUDA extension is missing
    torch.set_num_threads(1)  # N ranks share the host cores; the host side is launch-bound
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    t0 = time.time()
    R = get_map("R", 1, args.scale)
    S = get_map("S", 2 + rank, args.scale)
    bbox = synth.US_BBOX  # same scaling on every rank
    log("[rank %d] maps ready in %.1fs: R %d edges / %d chains, S %d edges / %d chains"
        % (rank, time.time() - t0, R.n_edges, R.n_chains, S.n_edges, S.n_chains))

    stream = torch.cuda.Stream(device=dev)
    ctx = RJ.Context(device=local_rank, stream=stream.cuda_stream)
    ctx.set_option("keep_host_graph", 0)
    ctx.set_option("lbvh_leaf_size", args.leaf_size)
    ctx.set_option("sort_queries", args.sort_queries)
    ctx.set_option("lsi_filter", args.filter)
    ctx.set_option("lsi_cells", args.cells)
    ctx.set_bounding_box(*bbox)
    ctx.set_map(0, R)
    # pinned host copies of the S batch for the end-to-end leg
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_xy, h_row, h_left, h_right = pin(S.xy), pin(S.row_index), pin(S.left), pin(S.right)

    def upload_S():
        ctx.set_map_raw(1, h_xy.data_ptr(), S.n_points, h_row.data_ptr(), h_left.data_ptr(),
                        h_right.data_ptr(), S.n_chains)
    upload_S()
    build_ms = [ctx.build_index(0, args.mode, args.grid_size) for _ in range(3)]
    idx = ctx.index_info(0, args.mode)
    lsi = RJ.LSI(ctx, args.mode)
    lsi.Init(XSECT_FACTOR)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    step_counts = []  # (results, candidates) of every step, exchanged in one all-gather

    def step():
        n = lsi.Query(1)
        step_counts.append((n, lsi.n_candidates))
        return n

    # fixed-size exchange buffers: the warm-up exchange has the shape of the timed one (NCCL
    # sets its connections up lazily, 2.7 ms on first use -- tools/nccl_small.py)
    n_rows = max(args.steps, args.warmup, 1)
    counts_host = torch.zeros((n_rows, 2), dtype=torch.int64).pin_memory()
    counts_dev = torch.zeros((n_rows, 2), dtype=torch.int64, device=dev)
    counts_all = torch.zeros((world, n_rows, 2), dtype=torch.int64, device=dev)

    def wait_counts():
        """The only data-path collective: the per-rank {result, candidate} counts of the
        steps since the last call, ONE NCCL all-gather (16 B per step and rank).  Nothing
        in a query depends on another rank's counts, so they are exchanged once per batch
        of steps instead of stalling every 0.16 ms step on a collective launch."""
        if world > 1 and step_counts:
            counts_host.zero_()
            counts_host[:len(step_counts)] = torch.tensor(step_counts[-n_rows:], dtype=torch.int64)
            with torch.cuda.stream(stream):
                counts_dev.copy_(counts_host, non_blocking=True)
                dist.all_gather_into_tensor(counts_all, counts_dev)
        step_counts.clear()

    for _ in range(args.warmup):
        step()
        if world > 1 and len(step_counts) == 1:
            wait_counts()  # twice during warm-up: connection set-up, then the steady state
            step_counts.append((0, 0))
    wait_counts()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:  # one sampler per job: NVML calls from N processes serialise in the driver
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    k_ms, p_ms = [], []
    n_pairs = 0
    torch.cuda.synchronize()
    gc.disable()  # a collection inside a 0.16 ms step shows up as a 5x outlier
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush_buf.zero_()  # evict L2 between timed iterations (outside the events)
            ev[i][0].record(stream)
            n_pairs = step()
            ev[i][1].record(stream)
        a, b = ctx.last_kernel_ms()
        k_ms.append(a)
        p_ms.append(b)
    torch.cuda.synchronize()
    gc.enable()
    # The count exchange of the K steps (the only data-path collective) is timed on its own,
    # with the ranks aligned first: inside the last step's interval it would mostly measure
    # how far the ranks had drifted apart over the untimed L2 flushes between the steps.
    xch = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        xch[0].record(stream)
        wait_counts()
        xch[1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    xch_ms = xch[0].elapsed_time(xch[1])
    total_ms = sum(step_ms) + xch_ms
    log("[rank %d] step ms min/median/max %.4f/%.4f/%.4f, count exchange %.4f ms; host cpus: %s (affinity %d)"
        % (rank, min(step_ms), float(np.median(step_ms)), max(step_ms), xch_ms, os.cpu_count(),
           len(os.sched_getaffinity(0))))
    n_cand = lsi.n_candidates

    # ---- end to end through the C ABI from pinned host buffers ----------------
    out_host = torch.empty(max(1, n_pairs) * 32, dtype=torch.uint8).pin_memory()
    out_np = out_host.numpy().view(RJ.XSECT_DTYPE)

    def e2e_step():
        upload_S()                                   # H2D + scale + edge numbering
        n = lsi.Query(1)                             # traversal + intersection points
        ctx.copy_to_host(lsi._res[0], out_np[:n])    # D2H of the rjb_xsect records
        return n
    for _ in range(min(3, args.warmup)):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t
    clocks = sampler.summary()

    times = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    sizes = torch.tensor([S.n_edges, n_pairs, n_cand], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms = times.tolist()
    all_edges, all_pairs, all_cand = sizes.tolist()

    h2d = S.xy.nbytes + S.row_index.nbytes + 4 * 2 * S.n_chains
    d2h = n_pairs * 32 + 16

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = all_edges / (ms_per_step / 1e3)
        peak, peak_src = measured_peak()
        # algorithmic (compulsory) bytes of one traversal launch with this layout
        # (DESIGN.md "Kernels"): S vertices + S edge->chain ids, the index once,
        # R vertices once, 8 B per result pair
        if args.mode == "lbvh":
            alg = 16 * S.n_points + 4 * S.n_edges + idx["bytes"] + 16 * R.n_points + 8 * n_pairs
            lst = ctx.last_stats()
            log("last_stats (results, candidates, ..., [5] cells path, [6] long survivors, [7] survivors):", lst)
            kname = ("k_lsi_filter+k_lsi_cells+k_lsi_bvh" if lst[5] else "k_lsi_filter+k_lsi_bvh") if lst[7] \
                else "k_lsi_bvh"
        else:
            alg = 16 * S.n_points + 4 * S.n_edges + idx["bytes"] + 16 * R.n_points + 4 * R.n_edges + 8 * n_pairs
            kname = "k_lsi_grid"
        sorted_queries = args.sort_queries > 0 or (args.sort_queries < 0 and S.n_edges / max(1, S.n_chains) < 32)
        k_avg = float(np.mean(k_ms))
        achieved = alg / (k_avg / 1e3) / 1e9
        line = {
            "metric": "LSI join throughput", "value": value, "unit": "query_edges/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64/int128", "data": "synthetic",
            "config": {"workload": WORKLOAD, "mode": args.mode, "xsect_factor": XSECT_FACTOR,
                       "lbvh_leaf_size": args.leaf_size, "sort_queries": args.sort_queries,
                       "l2": "256 MiB flush write between timed iterations; inputs (S vertices 146 MB) also exceed L2",
                       "sharding": "R + index replicated, S sharded per rank (seed 2+rank); NCCL: one "
                                   "all-gather of the per-step counts per timed region (timed after "
                                   "aligning the ranks, added to the K step times)"},
            "join_ms": ms_per_step, "result_pairs": int(all_pairs),
            "candidate_pairs": int(all_cand),
            "candidate_pairs_per_s": all_cand / (ms_per_step / 1e3),
            "index_build_ms": float(np.min(build_ms)), "index_bytes": int(idx["bytes"]),
            "index_units": int(idx["units"]),
            "kernel_ms": {kname: k_avg, "exact_pass(k_lsi_exact+k_lsi_points)" if args.mode == "lbvh"
                          else "k_xsect_points_dyn": float(np.mean(p_ms))},
            "e2e": {"value": all_edges / (e2e_ms / 1e3 / args.steps), "unit": "query_edges/s",
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": ((2 + len(kname.split("+")) if args.mode == "lbvh" else 2)
                             + (6 if sorted_queries else 0)) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(kname),
                         "algorithmic_bytes": int(alg), "peak_source": peak_src},
        }
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
                got = sort_xsects(out_np[:n_pairs].copy(), 1)
                ok = n_pairs == n_ref and all(np.array_equal(g, w) for g, w in zip(got, res[:4]))
                line["parity_vs_oracle"] = "bit-exact" if ok else "MISMATCH"
                if not ok:
                    log("PARITY MISMATCH against the oracle: %d vs %d pairs" % (n_pairs, n_ref))
        print(json.dumps(line), file=json_out, flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
