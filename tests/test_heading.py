"""Heading extension (fhp_params.heading, BASELINE config 4): CPU-side checks of the oracle and the host geometry.

The reference has no heading (the body always advances along +x, FP.cpp:2199-2202), so nothing of the reference pins a
rotated plan; what pins the extension is (i) the rotated expressions reduce BIT-EXACTLY to the reference's at theta = 0,
(ii) rotation covariance where the terrain does not break it (flat map), (iii) the rotated rectangle's PNPOLY against a
brute-force half-plane test, and — on the GPU (tests/test_gpu_heading.py) — kernel == oracle on the full sweep."""
import ctypes as C
import math

import numpy as np
import pytest

from quadrupedal_foothold_planner_b200 import _abi, planner, synth


def _plan(oracle, m, p, q, n, force=False):
    oracle.lib().oracle_force_heading_path(C.c_int32(1 if force else 0))
    try:
        return oracle.plan(m, p, q, n, centroid=True)
    finally:
        oracle.lib().oracle_force_heading_path(C.c_int32(0))


@pytest.mark.parametrize("res,start", [(0.02, (0, 0)), (0.02, (41, 17)), (0.01, (0, 0))])
def test_heading_zero_reduces_bit_exactly(oracle, res, start):
    """theta = 0 through the ROTATED expressions (forced) == the restated reference path, byte for byte, both chains."""
    n = 300 if res == 0.02 else 500
    m = synth.rough_map(n, n, res, seed=11, start=start)
    half = n * res / 2
    q = synth.uniform_queries(96, (-half + 0.8, half - 2.6), (-half + 0.8, half - 0.8), seed=5)
    # (faithful spiral mode on a shifted start index drives some stance polygons degenerate — zero area, non-finite
    # centroid — which the rotated expressions must survive identically: exactly-zero sin terms are dropped)
    for p in (_abi.yaml_params(), _abi.yaml_params(gait=_abi.FHP_GAIT_WALK, step_length=0.21, rf_first=1),
              _abi.yaml_params(spiral_mode=_abi.FHP_SPIRAL_FAITHFUL)):
        a = _plan(oracle, m, p, q, 8, force=False)
        b = _plan(oracle, m, p, q, 8, force=True)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes()
        assert int(a[2]["gait_cycles_succeed"].sum()) > 0  # not vacuous
    # the default chain too
    oracle.lib().oracle_force_heading_path(C.c_int32(1))
    try:
        ex = oracle.plan_ex(m, _abi.yaml_params(), q[:16], 6)
    finally:
        oracle.lib().oracle_force_heading_path(C.c_int32(0))
    ex0 = oracle.plan_ex(m, _abi.yaml_params(), q[:16], 6)
    for x, y in zip(ex, ex0):
        assert x.tobytes() == y.tobytes()


@pytest.mark.parametrize("theta", [0.3, math.pi / 2, 2.5, math.pi, -1.1, 5.9])
def test_flat_map_rotation_covariance(oracle, theta):
    """On a flat map every foothold is its nominal position, so the plan at heading theta is the heading-0 plan rotated
    by theta about the initial pose (incl. the -7 mm/cycle drift, which acts along the rotated normal)."""
    m = synth.flat_map()
    pose = np.array([0.4, -0.3, 0.0])
    q = _abi.make_queries([pose])
    for gait in (_abi.FHP_GAIT_TROT, _abi.FHP_GAIT_WALK):
        a = oracle.plan(m, _abi.yaml_params(gait=gait), q, 8)
        b = oracle.plan(m, _abi.yaml_params(gait=gait, heading=theta), q, 8)
        assert int(a[2]["n_rows"][0]) == 36 and int(b[2]["n_rows"][0]) == 36
        c, s = math.cos(theta), math.sin(theta)
        for chain in (0, 1):
            dx, dy = a[chain][0]["x"] - pose[0], a[chain][0]["y"] - pose[1]
            ex, ey = pose[0] + dx * c - dy * s, pose[1] + dx * s + dy * c
            assert np.max(np.abs(b[chain][0]["x"] - ex)) < 1e-12 and np.max(np.abs(b[chain][0]["y"] - ey)) < 1e-12
            assert np.array_equal(a[chain][0]["foot_id"], b[chain][0]["foot_id"])
            assert np.array_equal(a[chain][0]["gait_cycle_id"], b[chain][0]["gait_cycle_id"])
            assert np.allclose(a[chain][0]["z"], b[chain][0]["z"])


def test_rotated_rectangle_pnpoly_vs_half_planes(oracle):
    """gm::Polygon::isInside on the rotated search rectangle == the four half-plane tests of the same rectangle for
    points away from its boundary; vertex order and offsets as getSearchPolygon (FP.cpp:2496-2517) rotated."""
    rng = np.random.default_rng(2)
    R = float(np.float32(0.1))
    for theta in (0.0, 0.2, 1.0, math.pi / 2, 2.2, 4.0, 6.0):
        p = _abi.yaml_params(heading=theta)
        pose = (1.0, -2.0, 0.0)
        _, first = planner.initial_stance(p, pose)
        _, nxt, poly = planner.search_geometry(p, pose, 0.0, first)
        c, s = math.cos(theta), math.sin(theta)
        for f in range(4):
            ctr = nxt[f]
            offs = np.array([[R, R / 2], [R, -R / 2], [-R, -R / 2], [-R, R / 2]])
            exp_v = ctr + offs @ np.array([[c, s], [-s, c]])
            assert np.allclose(poly[f], exp_v, atol=1e-15)
            pts = ctr + rng.uniform(-0.2, 0.2, size=(400, 2))
            local = (pts - ctr) @ np.array([[c, -s], [s, c]])  # into the rectangle's frame
            margin = np.minimum(R - np.abs(local[:, 0]), R / 2 - np.abs(local[:, 1]))
            for pt, mg in zip(pts, margin):
                if abs(mg) < 1e-9:
                    continue
                assert oracle.polygon_inside(poly[f], pt[0], pt[1]) == (mg > 0)


def test_host_geometry_members_with_heading_match_the_oracle(oracle):
    rng = np.random.default_rng(4)
    for theta in (0.7, -2.0, math.pi / 2):
        for p in (_abi.yaml_params(heading=theta), _abi.yaml_params(gait=_abi.FHP_GAIT_WALK, heading=theta, rf_first=1)):
            pose = (float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5)), float(rng.uniform(-1, 1)))
            init, first = planner.initial_stance(p, pose)
            o_init, o_bias, o_first = oracle.initial_geometry(p, pose)
            assert np.array_equal(init[:, :2], o_init) and np.array_equal(first[:, :2], o_first)
            for adj in (0.0, -0.007, -0.049):
                cur = first + rng.normal(0, 0.03, size=(4, 3))
                c, nxt, poly = planner.search_geometry(p, pose, adj, cur)
                oc, onxt, opoly = oracle.search_geometry(p, pose, adj, cur)
                assert np.array_equal(c[:2], oc) and np.array_equal(nxt, onxt) and np.array_equal(poly, opoly)


def test_heading_plans_differ_and_advance_along_heading(oracle):
    """On rough terrain the body really advances along the heading: displacement of the stance centre projected on u."""
    m = synth.rough_map(500, 500, 0.02, seed=5, start=(0, 0))
    q = _abi.make_queries([[0.0, 0.0, 0.0]])
    seen = 0
    for theta in np.linspace(0, 2 * math.pi, 16, endpoint=False):
        nom, _, st = oracle.plan(m, _abi.yaml_params(heading=float(theta)), q, 8, centroid=False)
        n = int(st["n_rows"][0])
        if n < 12:
            continue
        seen += 1
        c0 = np.array([nom[0][:4]["x"].mean(), nom[0][:4]["y"].mean()])
        c1 = np.array([nom[0][n - 4:n]["x"].mean(), nom[0][n - 4:n]["y"].mean()])
        d = c1 - c0
        u = np.array([math.cos(theta), math.sin(theta)])
        assert d @ u > 0.1 and abs(d @ np.array([-u[1], u[0]])) < 0.15
    assert seen >= 8


def test_config_c4_shapes():
    m, params, q, n = synth.config_c4(n_poses=4, n_steps=3, rows=400, cols=400, n_headings=5)
    assert len(params) == 2 * 3 * 5 and len(q) == len(params) * 4 and n == 8
    assert params[0].heading == 0.0 and abs(params[1].heading - 2 * math.pi / 5) < 1e-15
    assert params[5].heading == 0.0 and params[5].step_length > params[0].step_length
    assert params[15].gait == _abi.FHP_GAIT_WALK and q["param_id"][4] == 1
