#!/usr/bin/env python
"""bench.py — footholds/sec of the batched foothold planner (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c1|c2|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (config.workload): BASELINE.json configs[2] "C3" — trot gait, 65,536 queries x 16 gait cycles on a
4000x4000 @ 0.02 m rough-terrain map with a wrapped circular-buffer start index (1234, 2777); one STEP = one plan of the
whole batch = 65,536 x 16 x 4 = 4,194,304 footholds (foothold = one (query, cycle, leg) search attempt).
Multi-GPU ("query-sharded at 1/2/4/8 B200"): the map is replicated once by an NCCL broadcast and the SAME 65,536 queries
are sharded over the ranks (strong scaling, no collective in the data path); `weak_scaling` in the line is the other
reading (every rank plans its own 65,536 queries).

value        = footholds/s with queries, map and outputs resident in HBM (CUDA events around K launches on the
               launching stream, max over ranks).
e2e          = the same metric through the C-ABI call a host makes (fhp_plan_stream on pinned HOST buffers: H2D of the
               queries, the plan, and every produced response row + cycle masks + statuses written into host memory
               inside the timed region; e2e.compact24 / e2e.fixed_pitch are the copy-engine variants of the same call).
roofline     = algorithmic bytes per launch / average launch duration vs the measured HBM copy bandwidth (the contract's
               definition; `dram` next to it is the kernel's real DRAM traffic from the committed ncu capture).
cpu_baseline = the CPU oracle (restatement of the reference's CPU path, pinned against the reference's own
               FootholdPlanner.cpp compiled in oracle/_ref) on the host cores of this box, bounded sample, rank 0, N=1.
other_workloads = BASELINE configs 1, 2, 4 (the full gait x step-length x heading sweep) and 5, a few steps each.
The oracle is used ONLY in the cpu_baseline / algorithmic-bytes legs and in --impl reference; never in the measured path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from quadrupedal_foothold_planner_b200 import _abi, synth  # noqa: E402
from quadrupedal_foothold_planner_b200.dist import shard_bounds  # noqa: E402

METRIC = "footholds/sec (batched queries)"
UNIT = "footholds/s"
KERNEL_NAMES = {"mask": "plan_kernel<false,kStageMask,false,false> (cell-class mask planes)",
                "tma": "plan_kernel<false,kStageTma,false,false> (2-D TMA boxes of the float layer)",
                "ldg": "plan_kernel<false,kStageLdg,false,false> (coalesced loads of the float layer)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic():
    """DRAM bytes per launch of the dominant kernel on the headline workload from the committed `ncu --set full`
    capture (profiles/plan_kernel_traffic.json) — a constant of that capture, NOT measured in this run."""
    p = os.path.join(ROOT, "profiles", "plan_kernel_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["traffic_bytes_per_launch"]), d.get("source", "profiles/plan_kernel_traffic.json")
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
C4_SETS = dict(n_steps=64, n_headings=64)


def make_workload(name: str, rank: int = 0, world: int = 1, with_map: bool = True, weak: bool = False,
                  c4_poses: int = 4096):
    """-> dict(map, queries (THIS rank's), n_total, cycles, params, desc, names).  Strong sharding: the config's query set
    is cut into contiguous blocks (dist.shard_bounds); weak=True (C3 only): every rank draws its own 65,536 queries.
    with_map=False: geometry only (non-root ranks receive the layers by NCCL broadcast)."""
    names = ["traversability", "elevation"]
    if name == "c3":
        seed_q = 4 + (1000 * rank if weak else 0)
        if with_map:
            m, q, cycles = synth.config_c3(seed_q=seed_q)
        else:
            m = _abi.HostMap(4000, 4000, 0.02, (0.0, 0.0), (1234, 2777))
            q = synth.uniform_queries(65536, (-40 + 1.0, 40 - 4.5), (-40 + 1.0, 40 - 1.0), seed=seed_q)
            cycles = 16
        params = [_abi.yaml_params()]
        desc = "C3: trot, 65536 queries x 16 cycles, 4000x4000 @0.02 m rough map, start index (1234,2777)"
    elif name == "c4":
        m, params, q, cycles = synth.config_c4(n_poses=c4_poses, rows=4000, cols=4000, with_map=with_map, **C4_SETS)
        desc = (f"C4: {{trot,walk}} x 64 step lengths x 64 headings x {c4_poses} poses = {len(q)} queries x 8 cycles, "
                f"{len(params)} parameter sets, 4000x4000 @0.02 m rough map")
    elif name == "c4h0":  # the heading-0 slice (the reference's own behaviour)
        m, params, q, cycles = synth.config_c4(n_poses=4096, rows=4000, cols=4000, with_map=with_map, n_steps=64)
        desc = "C4 heading-0 slice: {trot,walk} x 64 step lengths x 4096 poses = 524288 queries x 8 cycles, 4000x4000 @0.02 m"
    elif name == "c5":
        m, q, cycles = synth.config_c5(with_map=with_map)
        params = [_abi.yaml_params()]
        names = names + ["slope"]
        desc = ("C5: trot, 262144 queries x 32 cycles, 16000x16000 @0.01 m rough map, 3 layers (3.07 GB), "
                "start index (4321,12345)")
    elif name == "c5s":  # a C5-shaped workload on a 1/16 map (kernel experiments: same resolution, geometry and case mix)
        m, q, cycles = synth.config_c5(n_queries=65536, rows=4000, cols=4000, start=(1321, 2345), with_map=with_map)
        params = [_abi.yaml_params()]
        names = names + ["slope"]
        desc = "C5-small: trot, 65536 queries x 32 cycles, 4000x4000 @0.01 m rough map, 3 layers"
    elif name == "c2":
        m, q, cycles = synth.config_c2()
        params = [_abi.yaml_params(gait=_abi.FHP_GAIT_WALK)]
        desc = "C2: walk, 1024 queries x 32 cycles, 2000x2000 @0.01 m stairs/stepping stones"
    elif name == "c1":
        m, q, cycles = synth.config_c1()
        params = [_abi.yaml_params()]
        desc = "C1: trot, 1 query x 8 cycles, flat 500x500 @0.02 m"
    else:
        raise SystemExit(f"unknown workload {name}")
    n_total = len(q)
    if world > 1 and not weak:
        lo, hi = shard_bounds(n_total, world, rank)
        q = q[lo:hi]
    return {"name": name, "map": m, "queries": np.ascontiguousarray(q), "n_total": n_total, "cycles": cycles,
            "params": params, "desc": desc, "names": names}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = []
        smax = 0.0
        reasons = set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = max(smax, float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only users of oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def oracle_footholds_per_s(hmap, params, queries, cycles, n_threads):
    """The CPU oracle on a bounded sample.  -> (footholds/s, seconds)"""
    from oracle import binding as ob  # reported CPU baseline
    ob.lib()
    t0 = time.perf_counter()
    ob.plan(hmap, params, queries, cycles, n_threads=n_threads, centroid=False)
    dt = time.perf_counter() - t0
    return len(queries) * cycles * 4 / dt, dt


def algorithmic_bytes(hmap, params, queries, cycles, sample=256):
    """W_in: unique map-cell bytes the reference algorithm reads per foothold (oracle-instrumented, response-only
    variant = what fhp_plan computes when out_centroid is NULL), averaged over a query sample."""
    from oracle import binding as ob
    ob.lib()
    sel = queries[:: max(1, len(queries) // sample)][:sample]
    w, sp = ob.touch_bytes(hmap, params, sel, cycles, with_centroid_height=False)
    n = len(sel) * cycles * 4
    return float(w.sum()) / n, float(sp.sum()) / n


def cpu_single_plan_latency(runs=200):
    """BASELINE.md B1 on this box's host cores: the CPU port on C1, one thread and with the reference's threading (four
    std::threads per gait cycle, FP.cpp:863-909), and the as-written reference class (oracle/_ref) when it is built."""
    from oracle import binding as ob
    m, q, cycles = synth.config_c1()
    p = _abi.yaml_params()
    out = {}
    for key, lt in (("port_1_thread", False), ("port_4_leg_threads", True)):
        ms, st = ob.plan_latency(m, p, q[0], cycles, repeats=runs, leg_threads=lt)
        ms.sort()
        out[key] = {"p50_ms": float(ms[len(ms) // 2]), "p99_ms": float(ms[int(len(ms) * 0.99) - 1]), "runs": runs}
    try:
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfhp_ref.so")):
            ts = []
            for _ in range(12):
                t0 = time.perf_counter()
                ob.ref_plan(m, p, q, cycles, wrap_correct=False, centroid=False)
                ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            out["reference_as_written"] = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[-1], "runs": len(ts),
                                           "what": "FootholdPlanner.cpp itself (4 std::threads per cycle, whole-GridMap "
                                                   "copies per foothold test), one service call incl. the shim's setup"}
    except Exception as e:  # the checker must never take the bench down
        out["reference_as_written"] = {"error": str(e)[:200]}
    return out


def as_written_reference(hmap, params, queries, cycles, n=1):
    """The UNMODIFIED reference FootholdPlanner.cpp (oracle/_ref/libfhp_ref.so, compiled from /root/reference against
    header shims) on `n` queries of the workload: it copies the whole GridMap by value for every foothold test
    (FP.cpp:863-903, :2039, :2085, :2117, :2520), so one query on the 4000x4000 map costs seconds.  Reported next
    to the port for context; None when the prebuilt library is not in the tree."""
    try:
        from oracle import binding as ob
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfhp_ref.so")):
            return None
        t0 = time.perf_counter()
        ob.ref_plan(hmap, params, queries[:n], cycles, wrap_correct=True, centroid=False)
        dt = time.perf_counter() - t0
        return {"value": n * cycles * 4 / dt, "unit": UNIT, "cores": 4, "kind": "reference",
                "sample": f"{n} query x {cycles} cycles through the reference class (4 std::threads per cycle), {dt:.1f} s"}
    except Exception as e:
        return {"error": str(e)[:200]}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on this box's host cores, same metric/config, bounded sample per
    step.  Throughput is measured on the CPU port (oracle/, the reference's algorithm without its by-value map
    copies, all host threads); the as-written reference build (oracle/_ref) is timed on one query and reported in
    `as_written_reference` — at ~16 s per query on this map it cannot fill K steps within minutes."""
    if rank != 0:
        return
    wl = make_workload(args.workload, 0, 1, c4_poses=args.c4_poses)
    hmap, queries, cycles, params, desc = wl["map"], wl["queries"], wl["cycles"], wl["params"], wl["desc"]
    cores = os.cpu_count() or 1
    sample_q = min(len(queries), args.ref_sample)
    q = queries[:: max(1, len(queries) // sample_q)][:sample_q]  # strided: covers every parameter set of a sweep
    for _ in range(args.warmup):
        oracle_footholds_per_s(hmap, params, q[: max(64, sample_q // 8)], cycles, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_footholds_per_s(hmap, params, q, cycles, cores)
    dt = time.perf_counter() - t0
    fh = len(q) * cycles * 4 * args.steps
    v = fh / dt
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sample": f"{len(q)} of {len(queries)} queries per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(q)} queries x {cycles} cycles per step, all host threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_as_written:
        line["as_written_reference"] = as_written_reference(hmap, params, queries[1:], cycles, 1)
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries ONE JSON line: anything a library prints there during the run (NCCL's version banner under
    torchrun, for one) goes to stderr instead; emit() writes the line to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(1 if _REAL_STDOUT is None else _REAL_STDOUT, data)


def map_update_timing(ctx, hmap, params):
    """Cost of one map update through the host-facing ingest calls (pageable host memory, like a ROS callback holds):
    fhp_set_map on the converted layers, fhp_set_map_wire on the serialised grid_map_msgs/GridMap, and the one-off
    class-plane build that the first plan after an update pays."""
    from quadrupedal_foothold_planner_b200.gridmap_msg import serialize_gridmap
    nbytes = sum(a.nbytes for a in hmap.layers.values())
    out = {"layers_bytes": nbytes}
    ctx.set_map(hmap)  # allocates; the timed updates below reuse the buffers (steady state)
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.set_map(hmap)
        t.append(time.perf_counter() - t0)
    out["set_map_ms"] = round(min(t) * 1e3, 2)
    out["set_map_GBps"] = round(nbytes / min(t) / 1e9, 1)
    wire = serialize_gridmap(hmap)
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.set_map_wire(wire)
        t.append(time.perf_counter() - t0)
    out["set_map_wire_ms"] = round(min(t) * 1e3, 2)
    out["set_map_wire_GBps"] = round(len(wire) / min(t) / 1e9, 1)
    q = _abi.make_queries([[0.0, 0.0, 0.0]])
    p = params[0] if isinstance(params, (list, tuple)) else params
    ctx.plan(p, q, 1)  # first use of this geometry: tables, kernel attributes (not a per-update cost)
    best = None
    for _ in range(3):
        ctx.set_map(hmap)
        t0 = time.perf_counter()
        ctx.plan(p, q, 1)  # builds the class planes of the new map
        first = time.perf_counter() - t0
        t0 = time.perf_counter()
        ctx.plan(p, q, 1)
        d = max(0.0, first - (time.perf_counter() - t0))
        best = d if best is None else min(best, d)
    out["class_plane_build_ms"] = round(best * 1e3, 3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c1", "c4", "c4h0", "c5"])
    ap.add_argument("--ref-sample", type=int, default=16384, help="queries per step of the CPU arm")
    ap.add_argument("--cpu-sample", type=int, default=65536, help="queries of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-as-written", action="store_true", help="skip timing oracle/_ref in --impl reference")
    ap.add_argument("--no-others", action="store_true", help="skip other_workloads (profiling runs)")
    ap.add_argument("--others", default="", help="comma list overriding the default other_workloads")
    ap.add_argument("--other-steps", type=int, default=5)
    ap.add_argument("--c4-poses", type=int, default=4096, help="start poses per parameter set of C4 (config: 4096)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 10)")
    args = ap.parse_args()
    quiet_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)  # profiling runs pass --warmup 1; the driver passes >= 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from quadrupedal_foothold_planner_b200.planner import Fhp
    from quadrupedal_foothold_planner_b200 import dist as fdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: libfhp has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM) if world > 1 else x

    peak, peak_src = measured_peaks()
    staging = os.environ.get("FHP_STAGING", "mask")
    # an explicit non-default stream: fhp_plan_device treats a NULL stream as "use the context's own stream", and the
    # timing events must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0

    # the C5 map takes ~30 s of host time to generate: start on it now (rank 0), while the GPU legs of the headline run
    others = [w for w in (args.others.split(",") if args.others else
                          (["c1", "c2", "c4", "c5"] if world == 1 else ["c4", "c5"])) if w] if not args.no_others else []
    if args.workload != "c3":
        others = []
    prefetch = {}

    def _prefetch(name):
        t0 = time.perf_counter()
        prefetch[name] = (make_workload(name, rank, world, with_map=(rank == 0), c4_poses=args.c4_poses), time.perf_counter() - t0)

    pre_threads = []
    if "c5" in others and rank == 0:
        th = threading.Thread(target=_prefetch, args=("c5",), daemon=True)
        th.start()
        pre_threads.append(th)

    class Resident:
        """One workload resident on this rank's GPU: map adopted in place, queries / rows / statuses in HBM."""

        def __init__(self, wl, t_gen):
            self.wl = wl
            hmap, names = wl["map"], wl["names"]
            self.bcast_ms = None
            if world > 1:
                self.layers, self.bcast_ms = fdist.replicate_map(hmap if rank == 0 else None, hmap, names, local_rank)
            else:
                # (cols, rows) C-contiguous == column-major (rows, cols)
                self.layers = [torch.from_numpy(np.ascontiguousarray(hmap.layers[n].T)).cuda() for n in names]
            self.fhp = Fhp(local_rank)
            self.fhp.set_map_device(hmap, names, [t.data_ptr() for t in self.layers], keepalive=self.layers)
            q = wl["queries"]
            self.nq, self.cycles = len(q), wl["cycles"]
            self.R = _abi.rows_per_query(self.cycles)
            self.parr, self.npar = Fhp._params_array(wl["params"])
            self.d_q = torch.from_numpy(q.view(np.uint8)).cuda()
            self.d_nom = torch.empty(max(1, self.nq) * self.R * _abi.FOOTHOLD_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            self.d_st = torch.empty(max(1, self.nq) * _abi.STATUS_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            self.footholds_local = self.nq * self.cycles * 4
            self.footholds_total = wl["n_total"] * self.cycles * 4 if not wl.get("weak") else None
            self.t_gen = t_gen

        def step(self):
            self.fhp.plan_device(self.parr, self.npar, self.d_q, self.nq, self.cycles, self.d_nom, None, self.d_st,
                                 stream=stream.cuda_stream)

        def timed(self, steps, warmup):
            """-> (ms per step, max over ranks; libfhp launches on this rank)"""
            for _ in range(warmup):
                self.step()
            stream.synchronize()
            self.fhp.synchronize()
            barrier()
            l0 = self.fhp.kernel_launches()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(steps):
                self.step()
            ev1.record(stream)
            stream.synchronize()
            barrier()
            ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
            self.fhp.synchronize()
            return ms, self.fhp.kernel_launches() - l0

        def status(self):
            st = self.d_st.cpu().numpy().view(_abi.STATUS_DTYPE)[: self.nq]
            if self.nq and int((st["reserved"] & _abi.FHP_STATUS_ERR_MASK).max()) != 0:
                raise SystemExit("kernel reported error bits")
            return st

        def close(self):
            self.fhp.close()
            del self.layers, self.d_q, self.d_nom, self.d_st
            torch.cuda.empty_cache()

    def roofline_of(res, ms, with_w=True):
        """The contract's roofline for this rank's launch (rank 0 evaluates it; every rank's launch is the same size)."""
        wl = res.wl
        st = res.status()
        rows_written = float(st["n_rows"].sum())
        out = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src}
        if rank == 0 and with_w and res.nq:
            w_in, spiral_frac = algorithmic_bytes(wl["map"], wl["params"], wl["queries"], res.cycles)
            w_out = rows_written * _abi.FOOTHOLD_DTYPE.itemsize / res.footholds_local + 8.0 / (res.cycles * 4)
            achieved = res.footholds_local * (w_in + w_out) / (ms * 1e-3) / 1e9
            out.update({"achieved": achieved, "frac": achieved / peak,
                        "algorithmic_bytes_per_foothold": {"read_unique_cells": w_in, "rows_written": w_out},
                        "spiral_taken_frac": spiral_frac, "avg_launch_ms": ms})
            if achieved > peak:
                out["note"] = ("above 1 by construction, not by magic: the algorithmic bytes are the float32 cells the "
                               "reference reads (SURVEY 8d), the kernel reads their 2-bit classes (one sixteenth of the "
                               "bytes) and neighbouring windows share cache lines; its real DRAM traffic is in the ncu "
                               "captures under profiles/")
        cyc_ok = float((st["n_rows"].astype(np.int64) - 4).sum()) / 4.0
        out["produced_rows_per_launch"] = rows_written
        out["successful_cycle_frac"] = cyc_ok / max(1, res.nq * res.cycles)
        return out

    # ---- headline: C3 (or --workload), strong-sharded
    t_gen = time.perf_counter()
    wl = make_workload(args.workload, rank, world, with_map=(rank == 0 or args.workload in ("c1", "c2")),
                       c4_poses=args.c4_poses)
    t_gen = time.perf_counter() - t_gen
    res = Resident(wl, t_gen)
    nq, cycles, R = res.nq, res.cycles, res.R

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ms_per_step, launches = res.timed(args.steps, args.warmup)
    total_fh = float(wl["n_total"] * cycles * 4)
    value = total_fh / (ms_per_step * 1e-3)

    # ---- e2e through the host-facing C-ABI (pinned host buffers, copies inside the timed region)
    h_q = torch.from_numpy(wl["queries"].view(np.uint8).copy()).pin_memory()
    h_rows = torch.empty(max(1, nq) * R * _abi.FOOTHOLD_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_st = torch.empty(max(1, nq) * _abi.STATUS_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_offs = torch.empty(nq + 1, dtype=torch.int64).pin_memory()
    fhp = res.fhp

    h_mask = torch.zeros(max(1, nq) * ((cycles + 31) // 32) * 8, dtype=torch.int32).pin_memory()
    h_blocks = h_rows[(-h_rows.data_ptr()) % 128:]  # 128-byte aligned view

    def step_e2e():  # streamed response: the kernel writes the produced 20-byte points straight into the pinned host buffer
        fhp.plan_stream_raw(res.parr, res.npar, h_q, nq, cycles, h_blocks, h_mask, h_st)

    def step_e2e_compact():  # the service-shaped call: every query returns the rows it produced back to back, 24 bytes each
        fhp.plan_compact_raw(res.parr, res.npar, h_q, nq, cycles, h_rows, nq * R, h_offs, h_st, fmt=24)

    def step_e2e_fixed():  # fixed-pitch rows (fhp_plan): all nq*R row slots cross PCIe
        fhp.plan_raw(res.parr, res.npar, h_q, nq, cycles, h_rows, None, h_st)

    def time_host_calls(fn, n):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return dt / n

    e2e_steps = args.e2e_steps or min(args.steps, 10)
    e2e_s = time_host_calls(step_e2e, e2e_steps)
    h_x = torch.empty(max(1, nq) * R * _abi.FOOTHOLD24_DTYPE.itemsize, dtype=torch.uint8)  # pageable: a consumer's array

    def step_e2e_expanded():  # streamed response + the host-side expander into service-shaped 24-byte rows
        step_e2e()
        fhp._check(fhp.L.fhp_stream_expand(h_blocks.data_ptr(), h_mask.data_ptr(), h_st.data_ptr(), nq, cycles, h_x.data_ptr(),
                                           nq * R, h_offs.data_ptr(), 0))

    e2e_expanded_s = time_host_calls(step_e2e_expanded, e2e_steps)
    e2e_compact_s = time_host_calls(step_e2e_compact, e2e_steps)
    e2e_fixed_s = time_host_calls(step_e2e_fixed, e2e_steps) if world == 1 else None
    e2e_pageable_s = None
    if world == 1:  # a ROS node's std::vector is PAGEABLE memory: the same compact call on plain numpy buffers
        p_q = wl["queries"].copy()
        p_rows = np.empty(max(1, nq) * R, dtype=_abi.FOOTHOLD24_DTYPE)
        p_offs = np.empty(nq + 1, dtype=np.int64)
        p_st = np.empty(max(1, nq), dtype=_abi.STATUS_DTYPE)
        e2e_pageable_s = time_host_calls(
            lambda: fhp.plan_compact_raw(res.parr, res.npar, p_q, nq, cycles, p_rows, nq * R, p_offs, p_st, fmt=24),
            max(3, e2e_steps // 2))
        del p_q, p_rows, p_offs, p_st
    clocks = sampler.stop()
    st = res.status()
    rows_local = float(st["n_rows"].sum())
    d2h_local = int(rows_local * _abi.POINT20_DTYPE.itemsize + nq * _abi.STATUS_DTYPE.itemsize + 4 * h_mask.numel())
    # host-ingest ceiling of the same bytes: every rank copies its e2e D2H volume device -> pinned host, no kernels
    d_blob = torch.empty(max(64, d2h_local), dtype=torch.uint8, device="cuda")
    h_blob = h_rows[: max(64, d2h_local)]

    def copy_only():
        h_blob.copy_(d_blob, non_blocking=True)
        torch.cuda.synchronize()

    ceil_s = time_host_calls(copy_only, e2e_steps)
    d2h_total = sum_over_ranks(float(d2h_local))
    h2d_total = sum_over_ranks(float(h_q.numel()))
    del d_blob

    # ---- weak-scaling reading of the same config at N > 1 (every rank its own 65,536 queries)
    weak = None
    if world > 1 and args.workload == "c3":
        res.close()
        wlw = make_workload("c3", rank, world, with_map=(rank == 0), weak=True)
        wlw["weak"] = True
        resw = Resident(wlw, 0.0)
        ms_w, _ = resw.timed(max(5, args.steps // 2), 3)
        weak = {"value": sum_over_ranks(float(resw.footholds_local)) / (ms_w * 1e-3), "ms_per_step": ms_w,
                "queries_per_gpu": resw.nq, "what": "every rank plans its own 65,536 queries (device-timed, max over ranks)"}
        resw.close()
        res = None

    line = None
    if rank == 0:
        roof = roofline_of(res, ms_per_step) if res is not None else {"bound": "hbm", "peak": peak, "unit": "GB/s"}
        if res is None:  # N > 1: this rank's shard (the launch the roofline is about) was closed for the weak leg
            w_in, spiral_frac = algorithmic_bytes(wl["map"], wl["params"], wl["queries"], cycles)
            w_out = rows_local * _abi.FOOTHOLD_DTYPE.itemsize / max(1, nq * cycles * 4) + 8.0 / (cycles * 4)
            achieved = nq * cycles * 4 * (w_in + w_out) / (ms_per_step * 1e-3) / 1e9
            roof.update({"achieved": achieved, "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_foothold": {"read_unique_cells": w_in, "rows_written": w_out},
                         "spiral_taken_frac": spiral_frac, "avg_launch_ms": ms_per_step,
                         "produced_rows_per_launch": rows_local})
        traffic, traffic_src = committed_traffic()
        roof["traffic"] = traffic if (args.workload == "c3" and world == 1) else None
        roof["traffic_source"] = (f"constant of the committed ncu --set full capture ({traffic_src}); NOT measured in this run"
                                  if roof["traffic"] else None)
        if roof["traffic"]:
            roof["dram"] = {"GBps": traffic / (ms_per_step * 1e-3) / 1e9, "frac_of_peak": traffic / (ms_per_step * 1e-3) / 1e9 / peak,
                            "note": "the kernel's real DRAM rate: it reads 0.5-byte cell classes instead of the float "
                                    "layer and is L2-resident; the bound in practice is instruction issue (profiles/)"}
        roof["kernel"] = KERNEL_NAMES.get(staging, staging)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "queries_total": wl["n_total"], "queries_per_gpu": nq,
                       "gait_cycles": cycles, "footholds_per_step": int(total_fh),
                       "sharding": "contiguous blocks of the same query set over the ranks (dist.shard_bounds)",
                       "l2": "no explicit flush: every step streams %d MB of response rows through the 126 MB L2 "
                             "between its reads of the map (%d MB class planes + %d MB elevation layer)"
                             % (nq * R * 40 // 2 ** 20, wl["map"].rows * wl["map"].cols // 2 // 2 ** 20,
                                wl["map"].rows * wl["map"].cols * 4 // 2 ** 20),
                       "staging": staging, "map_generation_s": round(t_gen, 2)},
            "clocks": clocks,
            "e2e": {"value": total_fh / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_total),
                    "d2h_bytes_per_step": int(d2h_total), "steps": e2e_steps, "ms_per_step": e2e_s * 1e3,
                    "call": "fhp_plan_stream on pinned host buffers (queries H2D by the copy engine; the kernel writes the "
                            "produced 20-byte points {x, y f64, z f32} in cycle-major blocks, the success/kind masks and the "
                            "statuses straight into the mapped host buffers while it plans); wall clock, max over ranks",
                    "d2h_ceiling": {"GBps": d2h_total / ceil_s / 1e9, "ms": ceil_s * 1e3,
                                    "what": "the same D2H bytes copied device -> pinned host by every rank at once, no kernels"},
                    "frac_of_d2h_ceiling": ceil_s / e2e_s},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        line["e2e"]["stream_plus_expander"] = {"value": total_fh / e2e_expanded_s, "ms_per_step": e2e_expanded_s * 1e3,
                                               "call": "fhp_plan_stream, then fhp_stream_expand (host threads) into the "
                                                       "service-shaped 24-byte rows + offsets of fhp_plan_compact24; the "
                                                       "streamed blocks themselves need no expansion to be read (fhp_stream_row)"}
        line["e2e"]["compact24"] = {"value": total_fh / e2e_compact_s, "ms_per_step": e2e_compact_s * 1e3,
                                    "call": "fhp_plan_compact24 (rows back to back + offsets; chunked plan/scan/pack kernels, "
                                            "copy-engine D2H)"}
        if e2e_pageable_s:
            line["e2e"]["pageable"] = {"value": total_fh / e2e_pageable_s, "ms_per_step": e2e_pageable_s * 1e3,
                                       "call": "fhp_plan_compact24 on PAGEABLE host buffers (what a ROS node's std::vector "
                                               "is): the driver stages every copy through its own bounce buffers"}
        if e2e_fixed_s:
            line["e2e"]["fixed_pitch"] = {"value": total_fh / e2e_fixed_s, "d2h_bytes_per_step": int(nq * R * 40 + nq * 8),
                                          "call": "fhp_plan (every 40-byte row slot returned)"}
        if weak:
            line["weak_scaling"] = weak
    if res is not None:
        res.close()
    del h_q, h_blocks, h_rows, h_st, h_offs, h_mask, h_x

    # ---- the other BASELINE configs, a few steps each
    if others:
        ow = {}
        for name in others:
            try:
                if name == "c1":
                    if rank == 0:
                        ow["c1"] = {"workload": "C1: trot, 1 query x 8 cycles, flat 500x500 @0.02 m (latency config)",
                                    "see": "latency_single_plan"}
                    continue
                for th in pre_threads:
                    th.join()
                t0 = time.perf_counter()
                if name in prefetch:
                    w, tg = prefetch.pop(name)
                else:
                    w = make_workload(name, rank, world, with_map=(rank == 0 or name == "c2"), c4_poses=args.c4_poses)
                    tg = time.perf_counter() - t0
                if name == "c2" and rank != 0:
                    continue
                sm = ClockSampler(local_rank)
                sm.start()
                r = Resident(w, tg)
                ms, nl = r.timed(args.other_steps, 3)
                ck = sm.stop()
                fh_total = float(w["n_total"] * w["cycles"] * 4) if name != "c2" else float(r.footholds_local)
                roof = roofline_of(r, ms, with_w=True)
                if rank == 0:
                    ow[name] = {"workload": w["desc"], "value": fh_total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                "steps": args.other_steps, "warmup": 3, "queries_per_gpu": r.nq, "gpu_launches": int(nl),
                                "roofline": roof, "clocks": ck, "map_generation_s": round(tg, 2),
                                "map_broadcast_ms": r.bcast_ms,
                                "map_broadcast_GBps": (None if not r.bcast_ms else round(
                                    sum(t.numel() * 4 for t in r.layers) / (r.bcast_ms * 1e-3) / 1e9, 1))}
                r.close()
                del w, r
            except Exception as e:  # an extra workload must never take the headline down
                if rank == 0:
                    ow[name] = {"error": repr(e)[:300]}
        if rank == 0:
            line["other_workloads"] = ow

    if rank == 0:
        if world == 1:
            lat_ctx = Fhp(local_rank)
            m1, q1, c1 = synth.config_c1()
            lat_ctx.set_map(m1)
            p1 = _abi.yaml_params()
            for _ in range(10):
                lat_ctx.plan(p1, q1, c1)
            ts = []
            for _ in range(200):
                t0 = time.perf_counter()
                lat_ctx.plan(p1, q1, c1)
                ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            line["latency_single_plan"] = {"p50_ms": ts[100], "p99_ms": ts[197], "runs": 200,
                                           "what": "single query, 8 gait cycles, fhp_plan host->device->host, C1 flat map"}
            if not args.no_cpu_baseline:
                line["latency_single_plan"]["cpu"] = cpu_single_plan_latency()
            if args.workload == "c3":
                line["map_update"] = map_update_timing(lat_ctx, wl["map"], wl["params"])
            lat_ctx.close()
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sq = wl["queries"][:: max(1, nq // max(1, min(nq, args.cpu_sample)))][: args.cpu_sample]
            v, dt = oracle_footholds_per_s(wl["map"], wl["params"], sq, cycles, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{len(sq)} of {nq} queries x {cycles} cycles, {dt:.1f} s, all host threads"}
        emit(line)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
