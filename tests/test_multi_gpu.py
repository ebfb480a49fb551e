"""Two-GPU tests (skipped when fewer than 2 CUDA devices are visible): the C-ABI map replication
`fhp_broadcast_map` — ncclBroadcast resolved at run time from the NCCL already in the process — as a C++ host would
use it, i.e. WITHOUT torch.distributed: the ranks are plain processes that build an ncclComm_t from a shared
ncclUniqueId (ctypes on the same libnccl torch bundles)."""
import ctypes as C
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def _find_nccl():
    import torch
    site = os.path.dirname(os.path.dirname(torch.__file__))
    for pat in (os.path.join(site, "nvidia", "nccl", "lib", "libnccl.so.2"), "/usr/lib/x86_64-linux-gnu/libnccl.so.2"):
        hits = glob.glob(pat)
        if hits:
            return hits[0]
    return None


def _worker(rank, world, uid_bytes, ret):
    sys.path.insert(0, ROOT)
    import torch

    from oracle import binding as ob
    from quadrupedal_foothold_planner_b200 import _abi, dist, synth
    from quadrupedal_foothold_planner_b200.planner import Fhp
    from tests.parity import compare_plans

    torch.cuda.set_device(rank)
    nccl = C.CDLL(_find_nccl(), mode=C.RTLD_GLOBAL)
    uid = NcclUniqueId()
    C.memmove(C.byref(uid), uid_bytes, 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclUniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0

    full = synth.rough_map(400, 360, 0.02, seed=19, start=(77, 300))
    ctx = Fhp(rank)
    if rank == 0:
        ctx.set_map(full)  # only the root holds the map
    d, _, keep = (full if rank == 0 else dist.geometry_only(full)).desc(names=[])
    names = ["traversability", "elevation"]
    arr = (C.c_char_p * 2)(*[n.encode() for n in names])
    d.n_layers = 2
    d.layer_names = C.cast(arr, C.POINTER(C.c_char_p))
    rc = ctx.L.fhp_broadcast_map(ctx.h, C.byref(d), comm, 0, rank)
    assert rc == 0, ctx.L.fhp_last_error(ctx.h)
    q = synth.uniform_queries(300, (-3.0, 0.5), (-2.6, 2.6), seed=20)
    lo, hi = dist.shard_bounds(len(q), world, rank)
    got = ctx.plan(_abi.yaml_params(), q[lo:hi], 8, centroid=True)
    exp = ob.plan(full, _abi.yaml_params(), q[lo:hi], 8, centroid=True)
    ret.put((rank, compare_plans(got, exp, f"rank {rank}")))
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    ctx.close()


@pytest.mark.gpu
def test_fhp_broadcast_map_nccl_world2():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    path = _find_nccl()
    if path is None:
        pytest.skip("libnccl not found")
    nccl = C.CDLL(path)
    uid = NcclUniqueId()
    nccl.ncclGetUniqueId.argtypes = [C.POINTER(NcclUniqueId)]
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, bytes(uid.internal), ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(ret.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r for r, _ in out] == [0, 1] and all(d == 0.0 for _, d in out)
