"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink on
the GPU box, gloo in the CPU tests).

The reference is single-GPU (no collective anywhere in /root/reference).  Here the
base map R and its index are replicated on every rank, the query map S (edges for
LSI, vertices / points for PIP) is sharded by whole chains, and the only
collectives are
  * an all-gather of the per-rank {result, candidate} counts, and
  * a gather of the (variable length) rjb_xsect records / PIP answers to rank 0,
    with local edge ids mapped back to ids of the unsharded map.
"""
import numpy as np

from .capi import XSECT_DTYPE, PlanarGraph


def shard_bounds(g, world):
    """Chain ranges [c0, c1) per rank, balanced by point count (chains stay whole so
    RayJoin's eid = point - chain numbering survives with a constant offset)."""
    row = g.row_index.astype(np.int64)
    n_chains = g.n_chains
    if n_chains == 0:
        return [(0, 0)] * world
    targets = (np.arange(1, world) * row[-1]) // world
    cuts = np.searchsorted(row[:-1], targets, side="left")
    cuts = np.clip(cuts, 0, n_chains)
    edges = np.concatenate([[0], cuts, [n_chains]])
    edges = np.maximum.accumulate(edges)
    return [(int(edges[r]), int(edges[r + 1])) for r in range(world)]


def shard_graph(g, rank, world):
    """-> (shard PlanarGraph, eid_offset, point_offset): local eid + eid_offset is the
    eid in the full map; local point id + point_offset the global point id."""
    c0, c1 = shard_bounds(g, world)[rank]
    row = g.row_index.astype(np.int64)
    p0 = int(row[c0]) if g.n_chains else 0
    p1 = int(row[c1]) if g.n_chains else 0
    shard = PlanarGraph(g.xy[p0:p1], (row[c0:c1 + 1] - p0).astype(np.uint32) if c1 > c0 else
                        np.zeros(0, np.uint32), g.left[c0:c1], g.right[c0:c1], g.chain_id[c0:c1],
                        g.first_point[c0:c1], g.last_point[c0:c1], bbox=g.bbox)
    return shard, p0 - c0, p0


def allgather_counts(dist, values, device):
    """values: per-rank python ints -> int64 array [world, len(values)]."""
    import torch
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def gather_xsects(dist, xs_local, counts, eid_offset, query_map_id, device, dst=0):
    """Gathers the rjb_xsect records of every rank on `dst` (None elsewhere).  The
    query-side edge ids are rebased to the unsharded map.  Padded gather: the volume
    is tiny (32 B per intersection) against NVLink bandwidth, so latency dominates."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    xs = np.array(xs_local, dtype=XSECT_DTYPE, copy=True)
    xs["eid"][:, query_map_id] += np.uint32(eid_offset)
    n_max = int(max(counts)) if len(counts) else 0
    buf = np.zeros(max(n_max, 1), XSECT_DTYPE)
    buf[:len(xs)] = xs
    t = torch.from_numpy(buf.view(np.uint8).copy()).to(device)
    recv = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, recv, dst=dst)
    if rank != dst:
        return None
    parts = [recv[r].cpu().numpy().view(XSECT_DTYPE)[:int(counts[r])] for r in range(world)]
    return np.concatenate(parts) if parts else np.zeros(0, XSECT_DTYPE)


def gather_array(dist, local, counts, device, dst=0):
    """Variable-length 1-D arrays (PIP closest eids / face ids) -> concatenation on dst."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    local = np.ascontiguousarray(local)
    n_max = int(max(counts)) if len(counts) else 0
    buf = np.zeros(max(n_max, 1), local.dtype)
    buf[:len(local)] = local
    t = torch.from_numpy(buf.view(np.uint8).copy()).to(device)  # bytes: any dtype, any backend
    recv = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, recv, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([recv[r].cpu().numpy().view(local.dtype)[:int(counts[r])]
                           for r in range(world)])


def distributed_overlay(dist, graphs, mode="lbvh", grid_size=2048, xsect_factor=0.5, device=0,
                        torch_device=None, output=None, bbox=None):
    """Polygon overlay of graphs[0] x graphs[1] over all ranks of `dist` (one process
    per GPU).  Every rank holds two contexts -- (shard of map 0, full map 1) for
    IntersectEdge(0) and the location of map 0's vertices, (full map 0, shard of map 1)
    for the location of map 1's vertices -- so the base side and its index are always
    complete and replicated while the query side is sharded by whole chains.  Counts go
    through an all-gather, the results through padded gathers to rank 0, which imports
    them (rjb_overlay_finish), runs ComputeOutputPolygons and writes the chains.
    Returns (MapOverlay on rank 0 | None, phase dict)."""
    import time
    import torch
    from . import capi, synth
    world, rank = dist.get_world_size(), dist.get_rank()
    tdev = torch_device if torch_device is not None else torch.device("cuda", device)
    bbox = bbox or synth.union_bbox(*graphs)
    t0 = time.perf_counter()
    sh0, eoff0, poff0 = shard_graph(graphs[0], rank, world)
    sh1, eoff1, poff1 = shard_graph(graphs[1], rank, world)
    ctx_a = capi.Context([sh0, graphs[1]], device=device, bbox=bbox)
    ctx_b = capi.Context([graphs[0], sh1], device=device, bbox=bbox)
    ctx_a.build_index(1, mode, grid_size)
    ctx_b.build_index(0, mode, grid_size)
    t1 = time.perf_counter()
    lsi = capi.LSI(ctx_a, mode)
    lsi.Init(xsect_factor)
    n = lsi.Query(0)
    xs = lsi.get_xsects()
    pip_a = capi.PIP(ctx_a, mode)
    pip_a.Query(0)
    e0, f0 = pip_a.get_closest_eids(), pip_a.get_face_ids()
    pip_b = capi.PIP(ctx_b, mode)
    pip_b.Query(1)
    e1, f1 = pip_b.get_closest_eids(), pip_b.get_face_ids()
    t2 = time.perf_counter()
    counts = allgather_counts(dist, [n, sh0.n_points, sh1.n_points], tdev)
    all_xs = gather_xsects(dist, xs, counts[:, 0], eoff0, 0, tdev)
    g_e0 = gather_array(dist, e0, counts[:, 1], tdev)
    g_f0 = gather_array(dist, f0, counts[:, 1], tdev)
    g_e1 = gather_array(dist, e1, counts[:, 2], tdev)
    g_f1 = gather_array(dist, f1, counts[:, 2], tdev)
    t3 = time.perf_counter()
    ctx_a.close()
    ctx_b.close()
    phases = {"load_build_s": t1 - t0, "sharded_queries_s": t2 - t1, "gather_s": t3 - t2,
              "n_xsects_local": int(n)}
    if rank != 0:
        return None, phases
    ctx = capi.Context(list(graphs), device=device, bbox=bbox)
    ov = capi.MapOverlay(ctx, mode, grid_size, xsect_factor)
    ov.Finish(all_xs, [g_e0, g_e1], [g_f0, g_f1])
    if output:
        ov.WriteResult(output)
    phases["finish_s"] = time.perf_counter() - t3
    phases["n_xsects"] = int(len(all_xs))
    return ov, phases
