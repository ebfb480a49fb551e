"""Generates tests/golden/*.npz from the CPU oracle (oracle/).

The reference has NO tests, fixtures or golden vectors and cannot be built or imported here (ROS1 / grid_map /
Eigen / NLopt absent), so these vectors are produced by the oracle restatement and pin IT against regressions:
the CPU suite checks oracle == golden, the GPU suite checks libfhp == golden.  Inputs are regenerated from seeds
(synth.py); only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import binding as ob  # noqa: E402
from quadrupedal_foothold_planner_b200 import _abi, synth  # noqa: E402


def cases():
    """name -> (HostMap, params list, queries, cycles)"""
    out = {}
    m, q, n = synth.config_c1()
    out["c1_flat_trot"] = (m, [_abi.yaml_params()], q, n)
    m = synth.rough_map(400, 360, 0.02, seed=3, start=(123, 277))
    q = synth.uniform_queries(96, (-3.0, 0.8), (-2.6, 2.6), seed=4)
    out["rough_002_wrap"] = (m, [_abi.yaml_params()], q, 10)
    out["rough_002_faithful"] = (m, [_abi.yaml_params(spiral_mode=_abi.FHP_SPIRAL_FAITHFUL)], q, 10)
    m = synth.stairs_stones_map(1200, 600, 0.01, start=(7, 411))
    qa = synth.uniform_queries(24, (-5.4, -5.0), (-2.2, 2.2), seed=1)
    qb = synth.uniform_queries(24, (0.4, 0.8), (-2.2, 2.2), seed=2)
    out["stairs_001_walk"] = (m, [_abi.yaml_params(gait=_abi.FHP_GAIT_WALK)], np.concatenate([qa, qb]), 20)
    m = synth.rough_map(300, 300, 0.02, seed=8, start=(0, 0))
    ps = [_abi.yaml_params(), _abi.yaml_params(step_length=0.12, skew=0.03, rf_first=1),
          _abi.yaml_params(step_length=0.24, skew=0.06, gait=_abi.FHP_GAIT_WALK)]
    q = synth.uniform_queries(60, (-2.2, 0.2), (-2.0, 2.0), seed=9)
    q["param_id"] = np.arange(60) % 3
    out["multi_param"] = (m, ps, q, 8)
    return out


def member_case():
    m = synth.rough_map(300, 300, 0.01, seed=31, start=(41, 7))
    rng = np.random.default_rng(5)
    c = rng.uniform(-1.6, 1.6, size=(600, 2))
    rc = c + rng.uniform(-0.06, 0.06, size=c.shape)
    return m, _abi.yaml_params(), c, rc


def ref_cases():
    """Cases whose golden vectors come from the REFERENCE's own FootholdPlanner.cpp (oracle/_ref, built by
    oracle/ref_shim from /root/reference): name -> (HostMap, params, queries, cycles, upstream_spiral).
    Small maps only: the reference copies the whole GridMap by value for every spiral candidate."""
    out = {}
    m, q, n = synth.config_c1()
    out["ref_c1_flat_trot"] = (m, _abi.yaml_params(), q, n, True)
    m = synth.rough_map(300, 300, 0.02, seed=5, start=(0, 0))
    q = synth.uniform_queries(60, (-2.2, 0.8), (-2.2, 2.2), seed=6)
    out["ref_rough_002"] = (m, _abi.yaml_params(), q, 8, True)
    m = synth.rough_map(500, 500, 0.01, seed=7, start=(0, 0))
    q = synth.uniform_queries(60, (-1.7, 0.3), (-1.7, 1.7), seed=8)
    out["ref_rough_001"] = (m, _abi.yaml_params(), q, 8, True)
    m = synth.rough_map(300, 300, 0.02, seed=9, start=(33, 250))
    q = synth.uniform_queries(60, (-2.2, 0.8), (-2.2, 2.2), seed=10)
    out["ref_rough_002_start_upstream"] = (m, _abi.yaml_params(spiral_mode=_abi.FHP_SPIRAL_FAITHFUL), q, 8, True)
    out["ref_rough_002_start_repaired"] = (m, _abi.yaml_params(spiral_mode=_abi.FHP_SPIRAL_WRAP_CORRECT), q, 8, False)
    m = synth.stairs_stones_map(700, 400, 0.01, start=(0, 0))
    q = synth.uniform_queries(40, (-2.9, -1.2), (-1.2, 1.2), seed=11)
    out["ref_stairs_001_rf_first"] = (m, _abi.yaml_params(rf_first=1, step_length=0.16, skew=0.04), q, 8, True)
    # poses up to the map border, ONE gait cycle: the only abort the service can take is the opt chain's cycle-0 submap
    # failure (FP.cpp:931-934) — pins FHP_STATUS_OPT_SUBMAP_FAIL
    m = synth.rough_map(200, 200, 0.02, seed=13, start=(0, 0))
    q = synth.uniform_queries(80, (-2.8, 2.4), (-2.6, 2.6), seed=14)
    out["ref_border_002_one_cycle"] = (m, _abi.yaml_params(), q, 1, True)
    return out


def main():
    if ob.ref_available():
        for name, (m, p, q, n, upstream) in ref_cases().items():
            nom, cen, st, dfl, kpi = ob.ref_plan(m, p, q, n, wrap_correct=not upstream, extras=True)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), nominal=nom, centroid=cen, status=st, default=dfl,
                                kpi=kpi)
            done = st["reserved"] == 0
            print(name, "service calls completed", int(done.sum()), "of", len(q), "succeed",
                  int(st["gait_cycles_succeed"][done].sum()))
    for name, (m, ps, q, n) in cases().items():
        nom, cen, st = ob.plan(m, ps, q, n, centroid=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), nominal=nom, centroid=cen, status=st)
        print(name, "rows", int(st["n_rows"].sum()), "succeed", int(st["gait_cycles_succeed"].sum()))
    m, p, c, rc = member_case()
    fo, fv = ob.check_foothold(m, p, c, rc)
    co, cc = ob.check_centroid(m, p, c)
    np.savez_compressed(os.path.join(HERE, "members_001.npz"), foothold=fo, valid=fv, centroid=co, case=cc,
                        height=ob.mean_height(m, p, c), default=ob.check_default(m, p, c))
    print("members", int(fv.sum()), np.unique(cc, return_counts=True))


if __name__ == "__main__":
    main()
