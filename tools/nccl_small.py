"""Latency of the count exchange (a tiny NCCL all-gather) under torchrun: first call vs steady
state, with and without the host-to-device staging of the counts."""
import os, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev)
for mode in ("fresh", "prealloc"):
    mine = torch.zeros((20, 2), dtype=torch.int64, device=dev)
    out = torch.empty((world, 20, 2), dtype=torch.int64, device=dev)
    for it in range(6):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            if mode == "fresh":
                mine = torch.tensor([[i, i] for i in range(20)], dtype=torch.int64).to(dev, non_blocking=True)
                out = torch.empty((world, 20, 2), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(out, mine)
            e1.record(stream)
        torch.cuda.synchronize()
        if rank == 0:
            print(mode, it, "device ms %.3f" % e0.elapsed_time(e1), "host ms %.3f" % ((time.perf_counter() - t) * 1e3), flush=True)
dist.destroy_process_group()
