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


class _DevBytes:
    """A device allocation owned by the engine, as a zero-copy uint8 torch view."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def device_view(ptr, nbytes, torch_device):
    import torch
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=torch_device)
    return torch.as_tensor(_DevBytes(ptr, nbytes), device=torch_device)


def gather_bytes(dist, local, nbytes_per_rank, dst=0):
    """Variable-length gather of uint8 tensors to `dst` as ONE batch of point-to-point
    operations (NCCL: a grouped ncclSend / ncclRecv, device memory to device memory; gloo: the
    same calls on host tensors).  -> concatenation in rank order on dst, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [int(x) for x in nbytes_per_rank]
    if rank != dst:
        if sizes[rank]:
            for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst)]):
                r.wait()
        return None
    out = torch.empty(sum(sizes), dtype=torch.uint8, device=local.device)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ops = [dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r) for r in range(world) if r != dst and sizes[r]]
    out[offs[dst]:offs[dst + 1]].copy_(local)
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return out


def distributed_overlay(dist, graphs, mode="lbvh", grid_size=2048, xsect_factor=0.5, device=0,
                        torch_device=None, output=None, bbox=None):
    """Polygon overlay of graphs[0] x graphs[1] over all ranks of `dist` (one process per GPU;
    BASELINE.json configs[3]).  Every rank holds ONE context with both maps whole and both
    indexes (the build is cheaper than any broadcast); the query side of each phase is cut by
    whole chains: rank r runs IntersectEdge(0) on its window of map 0 (option lsi_window_*; edge
    ids stay global) and locates its range of the vertices of both maps.  NCCL carries the
    count all-gather and one grouped send / recv per result array into device memory of rank 0,
    which hands the device pointers to rjb_overlay_finish_device -- its resident maps and
    indexes are reused -- runs ComputeOutputPolygons and writes the chains.  With a host
    (gloo) group the same protocol runs on host copies.
    Returns (MapOverlay on rank 0 | None, phase dict)."""
    import time
    import torch
    from . import capi, synth
    world, rank = dist.get_world_size(), dist.get_rank()
    tdev = torch_device if torch_device is not None else torch.device("cuda", device)
    on_device = tdev.type == "cuda"
    bbox = bbox or synth.union_bbox(*graphs)
    t0 = time.perf_counter()
    ctx = capi.Context(list(graphs), device=device, bbox=bbox)
    ctx.set_option("sort_queries", 0)  # map vertices are coherent along their chains
    for im in range(2):
        ctx.build_index(im, mode, grid_size)
    # windows of this rank: whole chains -> point ranges of map 0 and map 1
    win = []
    for g in graphs:
        c0, c1 = shard_bounds(g, world)[rank]
        row = g.row_index.astype(np.int64)
        win.append((int(row[c0]), int(row[c1])) if g.n_chains else (0, 0))
    ctx.sync()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()

    def view(ptr, nbytes):
        # results live in buffers of the context that the next query reuses: take a copy
        if on_device:
            return device_view(ptr, nbytes, tdev).clone()
        out = np.empty(nbytes, np.uint8)
        ctx.copy_to_host(ptr, out)
        return torch.from_numpy(out)

    lsi = capi.LSI(ctx, mode)
    lsi.Init(xsect_factor)
    ctx.set_option("lsi_window_begin", win[0][0])
    ctx.set_option("lsi_window_end", win[0][1])
    n = lsi.Query(0) if win[0][1] > win[0][0] else 0
    ctx.set_option("lsi_window_begin", 0)
    ctx.set_option("lsi_window_end", 0)
    xs = view(lsi._res[0], 32 * n) if n else torch.empty(0, dtype=torch.uint8, device=tdev)
    loc = []
    for im in range(2):
        p0, p1 = win[im]
        if p1 > p0:
            de, df, _ = ctx.pip_device(im, mode, ctx.map_points_device_ptr(im) + 16 * p0, p1 - p0)
            loc.append((view(de, 4 * (p1 - p0)), view(df, 4 * (p1 - p0))))
        else:
            e = torch.empty(0, dtype=torch.uint8, device=tdev)
            loc.append((e, e.clone()))
    if on_device:
        torch.cuda.synchronize(tdev)
    t2 = time.perf_counter()
    counts = allgather_counts(dist, [n, win[0][1] - win[0][0], win[1][1] - win[1][0]], tdev)
    g_xs = gather_bytes(dist, xs, counts[:, 0] * 32)
    g_loc = [(gather_bytes(dist, loc[im][0], counts[:, 1 + im] * 4),
              gather_bytes(dist, loc[im][1], counts[:, 1 + im] * 4)) for im in range(2)]
    if on_device:
        torch.cuda.synchronize(tdev)
    t3 = time.perf_counter()
    phases = {"load_build_s": t1 - t0, "sharded_queries_s": t2 - t1, "gather_s": t3 - t2,
              "n_xsects_local": int(n), "collective": "nccl grouped send/recv (device)" if on_device
              else "host point-to-point (gloo)"}
    if rank != 0:
        ctx.close()
        return None, phases
    n_all = int(counts[:, 0].sum())
    ov = capi.MapOverlay(ctx, mode, grid_size, xsect_factor)
    if on_device:
        ov.FinishDevice(g_xs.data_ptr(), n_all, [g_loc[0][0].data_ptr(), g_loc[1][0].data_ptr()],
                        [g_loc[0][1].data_ptr(), g_loc[1][1].data_ptr()])
    else:
        ov.Finish(g_xs.numpy().view(XSECT_DTYPE),
                  [g_loc[0][0].numpy().view(np.uint32), g_loc[1][0].numpy().view(np.uint32)],
                  [g_loc[0][1].numpy().view(np.int32), g_loc[1][1].numpy().view(np.int32)])
    phases["finish_device_ms"] = dict(ov.phase_ms)
    t4 = time.perf_counter()
    if output:
        ov.WriteResult(output)
    phases["finish_s"] = t4 - t3
    phases["write_s"] = time.perf_counter() - t4
    phases["n_xsects"] = n_all
    return ov, phases
