"""Polygon overlay of the County x Zipcode-scale synthetic pair over N GPUs
(BASELINE.json configs[3]).  Launch with torchrun (one rank per GPU, NCCL):
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/overlay_multi.py
Rank 0 prints one JSON line with the phase times and, with --check, compares the
result with a single-GPU run."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
import rayjoin_b200 as RJ
from rayjoin_b200 import dist as rd, synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--mode", default="lbvh")
ap.add_argument("--xsect-factor", type=float, default=0.5)
ap.add_argument("--output", default="/tmp/rjb200_overlay_out.cdb")
ap.add_argument("--check", type=int, default=1)
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
A, B = bench.get_map("R", 1, args.scale), bench.get_map("S", 2, args.scale)
# first pass: NCCL builds its point-to-point connections lazily (0.3 s on first use of a peer);
# the timed pass is the steady state of a resident service
ov, cold = rd.distributed_overlay(dist, [A, B], mode=args.mode, xsect_factor=args.xsect_factor,
                                  device=local, torch_device=dev, output=None)
if ov is not None:
    ov.ctx.close()
dist.barrier(); torch.cuda.synchronize()
t = time.perf_counter()
ov, phases = rd.distributed_overlay(dist, [A, B], mode=args.mode, xsect_factor=args.xsect_factor,
                                    device=local, torch_device=dev,
                                    output=args.output if rank == 0 else None)
phases["cold_gather_s"] = cold["gather_s"]
torch.cuda.synchronize(); dist.barrier()
total = time.perf_counter() - t
if rank == 0:
    line = {"query": "overlay", "n_gpus": world, "mode": args.mode, "edges": [A.n_edges, B.n_edges],
            "total_s": total, "phases": phases, "device_phase_ms_rank0": ov.phase_ms,
            "output_bytes": os.path.getsize(args.output)}
    if args.check:
        ctx = RJ.Context([A, B], device=local)
        single = RJ.MapOverlay(ctx, args.mode, xsect_factor=args.xsect_factor)
        single.Run()  # warm-up (first-use allocations)
        t = time.perf_counter(); single.Run(); line["single_gpu_run_s"] = time.perf_counter() - t
        line["single_gpu_phase_ms"] = single.phase_ms
        ref = args.output + ".single"
        t = time.perf_counter(); single.WriteResult(ref); line["write_s"] = time.perf_counter() - t
        line["identical_to_single_gpu"] = open(ref, "rb").read() == open(args.output, "rb").read()
        ctx.close()
    print(json.dumps(line), flush=True)
dist.destroy_process_group()
