"""GPU parity of the heading extension (fhp_params.heading, BASELINE config 4): libfhp's rotated-rectangle kernels vs
the CPU oracle on identical inputs — cell indices, ids and status bit-exact, xyz within 1e-5 m (tests/parity.py)."""
import math
import os

import numpy as np
import pytest

from quadrupedal_foothold_planner_b200 import _abi, synth
from tests.parity import XYZ_TOL, compare_plans

pytestmark = pytest.mark.gpu


def _run_both(fhp, oracle, hmap, params, q, cycles, centroid=True, threads=8):
    fhp.set_map(hmap)
    got = fhp.plan(params, q, cycles, centroid=centroid)
    exp = oracle.plan(hmap, params, q, cycles, centroid=centroid, n_threads=threads)
    return got, exp


def test_c4_full_sweep_vs_oracle(fhp, oracle):
    """BASELINE config 4 as written, at 64 poses per parameter set: {trot, walk} x 64 step lengths x 64 headings =
    8192 parameter sets x 64 poses = 524,288 queries x 8 cycles in ONE call, every query against the oracle."""
    m, params, q, N = synth.config_c4(n_poses=64, n_steps=64, rows=1500, cols=1500, n_headings=64)
    assert len(params) == 8192 and len(q) == 8192 * 64
    got, exp = _run_both(fhp, oracle, m, params, q, N, centroid=False, threads=os.cpu_count() or 8)
    assert compare_plans(got, exp, "C4 sweep") <= XYZ_TOL
    st = got[2]
    owned = np.arange(got[0].shape[1])[None, :] < st["n_rows"][:, None]
    kinds = got[0]["kind"][owned]
    assert (kinds == _abi.FHP_KIND_CANDIDATE).sum() > 10000 and (kinds == _abi.FHP_KIND_DEFAULT).sum() > 10000
    # the heading really steers the plan: last stance centre minus first, per heading, points along u
    ok = st["n_rows"] >= 20
    pid = q["param_id"]
    hd = np.array([p.heading for p in params])[pid]
    n = st["n_rows"]
    idx = np.arange(len(q))
    last_x = np.stack([got[0]["x"][idx, np.maximum(n - 4 + k, 0)] for k in range(4)], 1).mean(1)
    last_y = np.stack([got[0]["y"][idx, np.maximum(n - 4 + k, 0)] for k in range(4)], 1).mean(1)
    dx, dy = last_x - got[0]["x"][:, :4].mean(1), last_y - got[0]["y"][:, :4].mean(1)
    along = dx * np.cos(hd) + dy * np.sin(hd)
    # (not every plan: a centroid window whose rows are all "edges" leaves that chain's foot at (0,0,0), FP.cpp:810-813,
    # and the reference then searches around a stance centroid dragged towards the origin — kept bit for bit)
    assert ok.sum() > 1000 and np.median(along[ok]) > 0.5 and np.mean(along[ok] > 0.2) > 0.7


@pytest.mark.parametrize("res,start,mode", [(0.02, (0, 0), _abi.FHP_SPIRAL_WRAP_CORRECT),
                                            (0.02, (123, 277), _abi.FHP_SPIRAL_WRAP_CORRECT),
                                            (0.02, (123, 277), _abi.FHP_SPIRAL_FAITHFUL),
                                            (0.01, (70, 5), _abi.FHP_SPIRAL_WRAP_CORRECT)])
def test_heading_rough_maps_both_chains(fhp, oracle, res, start, mode):
    """Rough terrain with NaNs, wrapped start index, both spiral modes, 0.01 m (3x3 foot stencil), both chains' rows."""
    n = 600 if res == 0.02 else 900
    m = synth.rough_map(n, n, res, seed=3, start=start)
    half = n * res / 2
    params = [_abi.yaml_params(heading=h, spiral_mode=mode, gait=g)
              for h in (0.0, 0.4, math.pi / 2, 2.0, math.pi, 3.9, 4.712, 5.5) for g in (_abi.FHP_GAIT_TROT, _abi.FHP_GAIT_WALK)]
    poses = synth.uniform_queries(48, (-half + 2.6, half - 2.6), (-half + 2.6, half - 2.6), seed=4)
    q = np.tile(poses, len(params))
    q["param_id"] = np.repeat(np.arange(len(params), dtype=np.int32), len(poses))
    got, exp = _run_both(fhp, oracle, m, params, q, 10)
    assert compare_plans(got, exp, f"heading rough res={res} start={start} mode={mode}") <= XYZ_TOL
    st = got[2]
    kinds = got[0]["kind"][np.arange(got[0].shape[1])[None, :] < st["n_rows"][:, None]]
    assert (kinds == _abi.FHP_KIND_CANDIDATE).sum() > 100


def test_heading_border_and_tiny_maps(fhp, oracle):
    """Poses whose rotated plans leave the map, maps narrower than a window."""
    m = synth.rough_map(200, 260, 0.02, seed=21, start=(3, 250))
    params = [_abi.yaml_params(heading=h) for h in np.linspace(0.1, 6.2, 12)]
    poses = synth.uniform_queries(40, (-2.3, 2.3), (-2.9, 2.9), seed=8)
    q = np.tile(poses, len(params))
    q["param_id"] = np.repeat(np.arange(len(params), dtype=np.int32), len(poses))
    got, exp = _run_both(fhp, oracle, m, params, q, 8)
    assert compare_plans(got, exp, "heading border") <= XYZ_TOL
    rng = np.random.default_rng(5)
    t = rng.random((16, 24), dtype=np.float32) ** 0.2
    e = rng.random((16, 24), dtype=np.float32)
    tiny = _abi.HostMap(16, 24, 0.02, pos=(0.3, -0.2), start=(8, 23), layers={"traversability": t, "elevation": e})
    q2 = synth.uniform_queries(64, (-0.1, 0.7), (-0.7, 0.3), seed=3)
    got, exp = _run_both(fhp, oracle, tiny, _abi.yaml_params(heading=1.3), q2, 4)
    assert compare_plans(got, exp, "heading tiny") <= XYZ_TOL


def test_heading_members(fhp, oracle):
    """checkFoothold / checkCandidateFoothold / checkCirclePolygonFoothold with a rotated rectangle, incl. rectangle
    edges passing close to cell centres (axis-parallel headings: the rotated edge lands within rounding of them)."""
    m = synth.rough_map(400, 400, 0.02, seed=97, start=(133, 393))
    t = np.clip(m.layers["traversability"] + 0.25, 0, 1).astype(np.float32)
    t[::3, ::2] = 0.75
    m.set_layer("traversability", t)
    fhp.set_map(m)
    rng = np.random.default_rng(11)
    c = rng.uniform(-4.1, 4.1, size=(3000, 2))
    r = c + rng.normal(0, 0.04, size=c.shape)
    for theta in (0.35, math.pi / 2, math.pi, 3 * math.pi / 2, 5.1):
        p = _abi.yaml_params(heading=theta)
        got, gv = fhp.check_candidate(p, c, r)
        exp, ev = oracle.check_candidate(m, p, c, r)
        assert np.array_equal(gv, ev) and gv.sum() > 100 and (gv == 0).sum() > 20
        ok = gv == 1
        assert np.array_equal(got["row"][ok], exp["row"][ok]) and np.array_equal(got["col"][ok], exp["col"][ok])
        assert np.array_equal(fhp.check_circle_polygon(p, c, r), oracle.check_circle_polygon(m, p, c, r))
        g2, v2 = fhp.check_foothold(p, c, r)
        e2, w2 = oracle.check_foothold(m, p, c, r)
        assert np.array_equal(v2, w2) and np.array_equal(g2["row"], e2["row"]) and np.array_equal(g2["kind"], e2["kind"])
        assert np.max(np.abs(g2["z"] - e2["z"])) <= XYZ_TOL


def test_default_chain_with_heading(fhp, oracle):
    m = synth.rough_map(500, 500, 0.02, seed=31, start=(11, 401))
    fhp.set_map(m)
    q = synth.uniform_queries(64, (-2.0, 2.0), (-2.0, 2.0), seed=9)
    p = _abi.yaml_params(heading=2.2, gait=_abi.FHP_GAIT_WALK)
    nom, cen, dfl, st = fhp.plan_ex(p, q, 8)
    enom, ecen, edfl, _, est = oracle.plan_ex(m, p, q, 8)
    assert compare_plans((nom, cen, st), (enom, ecen, est), "heading ex") <= XYZ_TOL
    assert compare_plans((dfl, None, st), (edfl, None, est), "heading default chain") <= XYZ_TOL


def test_forced_heading_path_at_zero_is_bit_identical(oracle):
    """theta = 0 through the ROTATED kernels (FHP_FORCE_HEADING=1) returns byte-identical rows to the reference-shaped
    kernels, for every staging path: the extension reduces exactly to the reference."""
    from quadrupedal_foothold_planner_b200.planner import Fhp
    m = synth.rough_map(600, 600, 0.02, seed=3, start=(123, 277))
    q = synth.uniform_queries(512, (-5.0, 2.0), (-5.0, 5.0), seed=4)
    saved = {k: os.environ.get(k) for k in ("FHP_FORCE_HEADING", "FHP_STAGING")}
    try:
        for staging in ("mask", "tma", "ldg"):
            os.environ["FHP_STAGING"] = staging
            outs = []
            for force in ("0", "1"):
                os.environ["FHP_FORCE_HEADING"] = force
                ctx = Fhp(0)
                ctx.set_map(m)
                for p in (_abi.yaml_params(), _abi.yaml_params(gait=_abi.FHP_GAIT_WALK, spiral_mode=_abi.FHP_SPIRAL_FAITHFUL)):
                    nom, cen, dfl, st = ctx.plan_ex(p, q, 12)
                    owned = np.arange(nom.shape[1])[None, :] < st["n_rows"][:, None]
                    outs.append((force, nom[owned].tobytes(), cen[owned].tobytes(), dfl[owned].tobytes(), st.tobytes()))
                ctx.close()
            half = len(outs) // 2
            for a, b in zip(outs[:half], outs[half:]):
                assert a[0] == "0" and b[0] == "1" and a[1:] == b[1:], staging
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    exp = oracle.plan(m, _abi.yaml_params(), q, 12, n_threads=8)
    assert int(exp[2]["gait_cycles_succeed"].sum()) > 500
