"""Fuzz the CPU restatement (oracle/) against the REFERENCE ITSELF (oracle/_ref: FootholdPlanner.cpp compiled in place by
oracle/ref_shim): random small maps (resolution 0.01 / 0.02 / 0.025 / 0.04, random circular-buffer start index), random
thresholds, step lengths and leg order, both spiral variants, random poses and 1-6 gait cycles; every service call the
reference completes must give the same rows BIT FOR BIT (x, y, z, foot and cycle ids, status), and the oracle's
FHP_STATUS_OPT_SUBMAP_FAIL may only be set on calls the reference aborted (and equals the abort when one cycle is asked).

    python tests/fuzz_reference.py [seed] [seconds]      # CPU only; needs /root/reference or a built oracle/_ref

Test infrastructure (tests/test_reference_build.py runs a short fixed-seed slice); round-2 campaigns: seeds 2, 3 and
11-14 (240 + 120 + 4 x 600 s): 52,824 cases, 425,478 completed reference calls, 0 mismatches."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from quadrupedal_foothold_planner_b200 import _abi, synth  # noqa: E402


def one_case(oracle, rng):
    """-> (completed reference calls compared, description) ; raises AssertionError on a mismatch."""
    res = float(rng.choice([0.02, 0.01, 0.025, 0.04]))
    rows, cols = int(rng.integers(120, 260)), int(rng.integers(120, 260))
    if res == 0.01:
        rows += 120
        cols += 120
    start = (0, 0) if rng.random() < 0.4 else (int(rng.integers(0, rows)), int(rng.integers(0, cols)))
    m = synth.rough_map(rows, cols, res, seed=int(rng.integers(1, 1 << 30)), start=start)
    upstream = bool(rng.random() < 0.7)  # upstream's spiral (bug-compatible on a shifted start) or the repaired one
    kw = {}
    if rng.random() < 0.5:
        kw["step_length"] = float(np.float32(rng.uniform(0.10, 0.26)))
        kw["skew"] = float(np.float32(kw["step_length"] / 4))
    if rng.random() < 0.3:
        kw["rf_first"] = 1
    if rng.random() < 0.3:
        kw["thr_default"] = float(np.float32(rng.uniform(0.75, 0.95)))
        kw["thr_candidate"] = float(np.float32(rng.uniform(0.5, 0.74)))
    p = _abi.yaml_params(spiral_mode=_abi.FHP_SPIRAL_FAITHFUL if upstream else _abi.FHP_SPIRAL_WRAP_CORRECT, **kw)
    hx, hy = rows * res / 2, cols * res / 2
    q = synth.uniform_queries(10, (-hx + 0.5, hx - 1.7), (-hy + 0.4, hy - 0.4), seed=int(rng.integers(1, 1 << 30)))
    cycles = int(rng.integers(1, 7))
    what = f"res {res} map {rows}x{cols} start {start} upstream_spiral {upstream} {kw} cycles {cycles}"
    rn, rc, rs = oracle.ref_plan(m, p, q, cycles, wrap_correct=not upstream)
    gn, gc, gs = oracle.plan(m, p, q, cycles, centroid=True)
    done = rs["reserved"] == 0
    flag = (gs["reserved"] & 0x80) != 0
    assert np.all(~done[flag]), f"{what}: abort flag set on a call the reference completed"
    if cycles == 1:
        assert np.array_equal(flag, ~done), f"{what}: abort flag != reference abort (one cycle)"
    for f in ("success", "gait_cycles_succeed", "n_rows"):
        assert np.array_equal(gs[f][done], rs[f][done]), f"{what}: status.{f}"
    own = (np.arange(rn.shape[1])[None, :] < rs["n_rows"][:, None]) & done[:, None]
    for g, r, chain in ((gn, rn, "nominal"), (gc, rc, "centroid")):
        for f in ("foot_id", "gait_cycle_id", "x", "y", "z"):
            assert np.array_equal(g[f][own], r[f][own]), f"{what}: {chain}.{f}"
    return int(done.sum()), what


def run(oracle, seed, seconds=None, max_cases=None):
    rng = np.random.default_rng(seed)
    t0 = time.time()
    cases = calls = 0
    while (seconds is None or time.time() - t0 < seconds) and (max_cases is None or cases < max_cases):
        calls += one_case(oracle, rng)[0]
        cases += 1
    return cases, calls


if __name__ == "__main__":
    from tests import oracle_binding
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    t = time.time()
    c, k = run(oracle_binding, seed, seconds=secs)
    print(f"seed {seed}: {c} cases, {k} completed reference calls equal bit for bit, {time.time() - t:.0f} s")
