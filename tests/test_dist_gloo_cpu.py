"""CPU suite: the N>1 host logic (sharding, count all-gather, result gather) over a
world_size-2 gloo group.  The per-rank "query" is the CPU oracle so that the whole
distributed protocol is exercised without a GPU: sharded result == unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import OracleMaps, dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as O
    from rayjoin_b200 import dist as rd
    from rayjoin_b200 import synth
    from rayjoin_b200.capi import XSECT_DTYPE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    R, S = dataset(name)
    bbox = synth.union_bbox(R, S)
    shard, eid_off, pt_off = rd.shard_graph(S, rank, world)
    om = OracleMaps(O, [R, shard])
    om.sc = O.scaling_init(*bbox)               # scaling of the UNSHARDED pair on every rank
    om.pts = [O.scale_points(om.sc, g.xy) for g in (R, shard)]
    eq, eb, x, y = om.lsi(1)
    xs = np.zeros(len(eq), XSECT_DTYPE)
    xs["eid"][:, 1], xs["eid"][:, 0], xs["x"], xs["y"] = eq, eb, x, y
    counts = rd.allgather_counts(dist, [len(xs), shard.n_points], torch.device("cpu"))
    assert counts.shape == (world, 2) and counts[rank, 0] == len(xs)
    allx = rd.gather_xsects(dist, xs, counts[:, 0], eid_off, 1, torch.device("cpu"))
    pip = om.pip(1, om.pts[1])
    allp = rd.gather_array(dist, pip, counts[:, 1], torch.device("cpu"))
    if rank == 0:
        np.savez(out_path, xs=allx, pip=allp)
    else:
        assert allx is None and allp is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["voronoi", "soup"])
def test_sharded_query_equals_unsharded(oracle, name, tmp_path):
    world, port = 2, 29500 + (os.getpid() % 500)
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, port, name, out), nprocs=world, join=True)
    z = np.load(out)
    R, S = dataset(name)
    om = OracleMaps(oracle, [R, S])
    eq, eb, x, y = om.lsi(1)
    xs = z["xs"]
    o = np.lexsort((xs["eid"][:, 0], xs["eid"][:, 1]))
    assert np.array_equal(xs["eid"][o, 1], eq) and np.array_equal(xs["eid"][o, 0], eb)
    assert np.array_equal(xs["x"][o], x) and np.array_equal(xs["y"][o], y)
    assert np.array_equal(z["pip"], om.pip(1, om.pts[1]))


def test_shard_bounds_cover_and_balance():
    from rayjoin_b200 import dist as rd
    R, S = dataset("voronoi")
    for world in (1, 2, 3, 8):
        b = rd.shard_bounds(S, world)
        assert b[0][0] == 0 and b[-1][1] == S.n_chains
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        total_e, total_p = 0, 0
        for r in range(world):
            sh, eo, po = rd.shard_graph(S, r, world)
            assert eo == total_e and po == total_p
            total_e += sh.n_edges
            total_p += sh.n_points
            if sh.n_chains:
                assert sh.row_index[0] == 0 and sh.row_index[-1] == sh.n_points
        assert total_e == S.n_edges and total_p == S.n_points


def _gather_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rayjoin_b200 import dist as rd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [(7 * r) % 5 * 1000 + (0 if r == 1 else 13) for r in range(world)]  # rank 1 sends nothing
    sizes[1] = 0
    local = torch.full((sizes[rank],), rank + 1, dtype=torch.uint8)
    counts = rd.allgather_counts(dist, [sizes[rank]], torch.device("cpu"))[:, 0]
    assert counts.tolist() == sizes
    out = rd.gather_bytes(dist, local, counts)
    if rank == 0:
        np.save(out_path, out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_variable_length_gather(world, tmp_path):
    """gather_bytes: the grouped point-to-point gather of the overlay results (ragged, with an
    empty contribution) -- on NCCL the same calls move device memory."""
    out = str(tmp_path / "g.npy")
    mp.spawn(_gather_worker, args=(world, 29700 + world + (os.getpid() % 200), out), nprocs=world, join=True)
    got = np.load(out)
    sizes = [(7 * r) % 5 * 1000 + 13 for r in range(world)]
    sizes[1] = 0
    want = np.concatenate([np.full(sizes[r], r + 1, np.uint8) for r in range(world)])
    assert np.array_equal(got, want)
