"""Multi-rank overlay protocol (sharded LSI / vertex location, gather, finish on rank 0)
must reproduce the single-process overlay byte for byte.  Two ranks share cuda:0 here
(they only synchronise through host-side gloo collectives, never through kernels); the
NCCL path is the same code with device tensors (tools/overlay_multi.py)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from rayjoin_b200 import dist as rd, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = synth.voronoi_map(40, 2500, synth.BRAZIL_BBOX, seed=31)
    B = synth.share_chains(A, synth.voronoi_map(120, 3500, synth.BRAZIL_BBOX, seed=32), frac=0.2)
    ov, phases = rd.distributed_overlay(dist, [A, B], mode="lbvh", xsect_factor=4.0, device=0,
                                        torch_device=torch.device("cpu"), output=out_path)
    assert (ov is not None) == (rank == 0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_distributed_overlay_equals_single_process(rjb, world, tmp_path):
    from rayjoin_b200 import synth
    out = str(tmp_path / "dist.cdb")
    mp.spawn(_worker, args=(world, 29600 + world + (os.getpid() % 300), out), nprocs=world, join=True)
    A = synth.voronoi_map(40, 2500, synth.BRAZIL_BBOX, seed=31)
    B = synth.share_chains(A, synth.voronoi_map(120, 3500, synth.BRAZIL_BBOX, seed=32), frac=0.2)
    ctx = rjb.Context([A, B])
    ov = rjb.MapOverlay(ctx, "lbvh", xsect_factor=4.0)
    ov.Run()
    single = str(tmp_path / "single.cdb")
    ov.WriteResult(single)
    ctx.close()
    assert open(out).read() == open(single).read() and os.path.getsize(out) > 1000
