"""Multi-GPU plumbing: one process per GPU (torch.distributed), queries sharded, map replicated once.

The path shards by query: each query carries its own chain state and the map is read-only during a
plan, so there is NO collective in the data path.  The only collective is the map replication after a map
update: one broadcast per layer over NCCL/NVLink (gloo on CPU for the tests).  Results are disjoint row blocks;
`gather_rows` concatenates them on rank 0 when a single response array is wanted.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from ._abi import FOOTHOLD_DTYPE, STATUS_DTYPE, HostMap


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of ceil(n/world) queries per rank (last ranks may be short or empty)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def geometry_only(hmap: HostMap) -> HostMap:
    """rows/cols/resolution/position/start index without layer data (what non-root ranks know up front)."""
    return HostMap(hmap.rows, hmap.cols, hmap.resolution, hmap.pos, hmap.start)


def replicate_map(hmap: Optional[HostMap], geometry: HostMap, names: Sequence[str], device_index: Optional[int],
                  root: int = 0):
    """Broadcast the named layers from `root` to every rank.  -> (list of torch tensors, milliseconds)

    device_index None => CPU tensors over gloo (tests); otherwise CUDA tensors over NCCL.  A layer travels as the
    (cols, rows) C-contiguous view of the column-major (rows, cols) matrix, i.e. byte-identical to grid_map's
    Eigen storage."""
    import time

    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    dev = "cpu" if device_index is None else f"cuda:{device_index}"
    tensors: List["torch.Tensor"] = []
    for n in names:
        if rank == root:
            assert hmap is not None
            t = torch.from_numpy(np.ascontiguousarray(hmap.layers[n].T)).to(dev)
        else:
            t = torch.empty((geometry.cols, geometry.rows), dtype=torch.float32, device=dev)
        tensors.append(t)
    if device_index is not None:
        # communicator set-up / first-use cost is not map replication: pay it on a tiny tensor first
        warm = torch.zeros(8, dtype=torch.float32, device=dev)
        dist.broadcast(warm, src=root)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in tensors:
            dist.broadcast(t, src=root)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        t0 = time.perf_counter()
        for t in tensors:
            dist.broadcast(t, src=root)
        ms = (time.perf_counter() - t0) * 1e3
    return tensors, ms


def plan_sharded(plan_fn: Callable, queries: np.ndarray, world: int, rank: int):
    """Run `plan_fn(local_queries) -> (nominal, centroid|None, status)` on this rank's block of the GLOBAL query
    array.  No communication."""
    lo, hi = shard_bounds(len(queries), world, rank)
    return (lo, hi), plan_fn(queries[lo:hi])


def gather_rows(local, n_total: int, rows_per_query: int, world: int, rank: int, root: int = 0):
    """Concatenate per-rank (nominal, centroid|None, status) blocks on `root` in global query order (host arrays;
    object gather, works over gloo and NCCL).  Returns the global triple on root, None elsewhere."""
    import torch.distributed as dist

    nom, cen, st = local
    parts = [None] * world if rank == root else None
    dist.gather_object((nom, cen, st), parts, dst=root)
    if rank != root:
        return None
    gn = np.zeros((n_total, rows_per_query), dtype=FOOTHOLD_DTYPE)
    gc = np.zeros((n_total, rows_per_query), dtype=FOOTHOLD_DTYPE) if cen is not None else None
    gs = np.zeros(n_total, dtype=STATUS_DTYPE)
    for r, (pn, pc, ps) in enumerate(parts):
        lo, hi = shard_bounds(n_total, world, r)
        gn[lo:hi] = pn
        if gc is not None:
            gc[lo:hi] = pc
        gs[lo:hi] = ps
    return gn, gc, gs
