"""N>1 path on CPU: world_size-2 gloo processes replicate the map, plan their query shard with NO data-path
collective, and gather the rows; the result must equal the unsharded plan.  The per-rank planner is injected
(`plan_fn`): the product passes Fhp.plan, this CPU test passes the oracle (test infrastructure)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as td

    from oracle import binding as ob
    from quadrupedal_foothold_planner_b200 import _abi, dist, synth

    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = synth.rough_map(200, 160, 0.02, seed=9, start=(17, 150))
    geometry = dist.geometry_only(full)
    names = ["traversability", "elevation"]
    tensors, ms = dist.replicate_map(full if rank == 0 else None, geometry, names, None)
    # rebuild a HostMap from the received layers: byte-identical to the root's
    local = _abi.HostMap(geometry.rows, geometry.cols, geometry.resolution, geometry.pos, geometry.start)
    for n, t in zip(names, tensors):
        local.set_layer(n, t.numpy().T)
    same = all(np.array_equal(local.layers[n], full.layers[n], equal_nan=True) for n in names)
    queries = synth.uniform_queries(101, (-1.5, 0.2), (-1.2, 1.2), seed=2)
    p = _abi.yaml_params()
    (lo, hi), part = dist.plan_sharded(lambda q: ob.plan(local, p, q, 6, centroid=True), queries, world, rank)
    glob = dist.gather_rows(part, len(queries), _abi.rows_per_query(6), world, rank)
    if rank == 0:
        ref = ob.plan(full, p, queries, 6, centroid=True)
        ok = all(np.array_equal(a, b) for a, b in zip(glob, ref))
        ret.put((same, ok, lo, hi))
    else:
        ret.put((same, True, lo, hi))
    td.destroy_process_group()


def test_gloo_world2_replicate_shard_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    results = [ret.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in results), "replicated layers differ from the root's"
    assert all(r[1] for r in results), "sharded + gathered plan differs from the unsharded plan"
    assert sorted((r[2], r[3]) for r in results) == [(0, 51), (51, 101)]
