"""GPU parity: libfhp (CUDA, through the C-ABI) vs the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest

from quadrupedal_foothold_planner_b200 import _abi, synth
from tests.parity import XYZ_TOL, compare_plans

pytestmark = pytest.mark.gpu


def _run_both(fhp, oracle, hmap, params, q, cycles, centroid=True):
    fhp.set_map(hmap)
    got = fhp.plan(params, q, cycles, centroid=centroid)
    exp = oracle.plan(hmap, params, q, cycles, centroid=centroid)
    return got, exp


def test_c1_flat_closed_form(fhp, oracle):
    """BASELINE config 1: trot, 8 cycles, one query, flat 500x500 @ 0.02 m — known answer."""
    m, q, N = synth.config_c1()
    p = _abi.yaml_params()
    got, exp = _run_both(fhp, oracle, m, p, q, N)
    assert compare_plans(got, exp, "C1") == 0.0
    nom, _, st = got
    assert st[0]["success"] == 1 and st[0]["gait_cycles_succeed"] == 8 and st[0]["n_rows"] == 36
    f32 = np.float32
    L = float(f32(0.4387)); W = float(f32(0.175) + f32(0.037) * f32(2)); s = float(f32(0.04))
    sl = float(f32(0.18)); hs = float(f32(0.18) / f32(2))
    bx = [0.5 * L - s, -0.5 * L + s, -0.5 * L - s, 0.5 * L + s]
    by = [-0.5 * W, -0.5 * W, 0.5 * W, 0.5 * W]
    for g in range(8):
        for f in range(4):
            r = nom[0][4 + 4 * g + f]
            assert abs(r["x"] - (-4 - hs + sl * (g + 1) + bx[f])) < 1e-12
            assert abs(r["y"] - (by[f] - 0.007 * g)) < 1e-12
            assert r["z"] == float(f32(0.01)) and r["foot_id"] == f and r["gait_cycle_id"] == g


@pytest.mark.parametrize("start", [(0, 0), (123, 277)])
@pytest.mark.parametrize("mode", [_abi.FHP_SPIRAL_WRAP_CORRECT, _abi.FHP_SPIRAL_FAITHFUL])
def test_rough_map_trot(fhp, oracle, start, mode):
    """C3-shaped: rough terrain @0.02 m with NaNs and a wrapped start index; both spiral modes."""
    m = synth.rough_map(600, 600, 0.02, seed=3, start=start)
    q = synth.uniform_queries(512, (-5.0, 2.0), (-5.0, 5.0), seed=4)
    p = _abi.yaml_params(spiral_mode=mode)
    got, exp = _run_both(fhp, oracle, m, p, q, 12)
    d = compare_plans(got, exp, f"rough start={start} mode={mode}")
    assert d <= XYZ_TOL
    st = got[2]
    # the case mix must actually exercise the spiral and failures
    kinds = got[0]["kind"][np.arange(got[0].shape[1])[None, :] < st["n_rows"][:, None]]
    assert (kinds == _abi.FHP_KIND_CANDIDATE).sum() > 100 and (kinds == _abi.FHP_KIND_DEFAULT).sum() > 100


def test_start_index_invariance(fhp):
    """plan(map with start s) == plan(same content re-laid-out with start 0) in wrap_correct mode."""
    m = synth.rough_map(500, 420, 0.02, seed=11, start=(77, 301))
    q = synth.uniform_queries(256, (-4.0, 1.5), (-3.5, 3.5), seed=5)
    p = _abi.yaml_params()
    fhp.set_map(m)
    a = fhp.plan(p, q, 10, centroid=True)
    fhp.set_map(m.relaid_start0())
    b = fhp.plan(p, q, 10, centroid=True)
    assert np.array_equal(a[2], b[2])
    keep = np.arange(a[0].shape[1])[None, :] < a[2]["n_rows"][:, None]  # rows beyond n_rows are unspecified
    for f in ("x", "y", "z", "foot_id", "gait_cycle_id", "kind"):
        assert np.array_equal(a[0][f][keep], b[0][f][keep]) and np.array_equal(a[1][f][keep], b[1][f][keep])


def test_c2_walk_stairs_stones(fhp, oracle):
    """BASELINE config 2 (reduced query count): walk preset, 32 cycles, stairs/stepping stones @0.01 m."""
    m, q, N = synth.config_c2(n_queries=128)
    p = _abi.yaml_params(gait=_abi.FHP_GAIT_WALK)
    got, exp = _run_both(fhp, oracle, m, p, q, N)
    assert compare_plans(got, exp, "C2") <= XYZ_TOL


def test_multi_param_sets(fhp, oracle):
    """Per-query parameter sets (the C4 sweep shape): step length / skew / gait vary, radii shared."""
    m = synth.rough_map(500, 500, 0.02, seed=7, start=(5, 9))
    ps = [_abi.yaml_params(),
          _abi.yaml_params(step_length=0.12, skew=0.03, rf_first=1),
          _abi.yaml_params(step_length=0.24, skew=0.06, gait=_abi.FHP_GAIT_WALK),
          _abi.yaml_params(gait=_abi.FHP_GAIT_CUSTOM, thr_default=0.8, thr_candidate=0.5,
                           bias=[[0.2, -0.1], [-0.2, -0.1], [-0.25, 0.1], [0.25, 0.1]])]
    q = synth.uniform_queries(400, (-4.0, 1.0), (-4.0, 4.0), seed=9)
    q["param_id"] = np.arange(400) % 4
    got, exp = _run_both(fhp, oracle, m, ps, q, 8)
    assert compare_plans(got, exp, "multi-param") <= XYZ_TOL


def test_edges_border_nan_blocked(fhp, oracle):
    """Queries at / beyond the map border, an all-NaN map, an all-blocked map, zero cycles, 255 cycles."""
    p = _abi.yaml_params()
    m = synth.rough_map(300, 260, 0.02, seed=21, start=(17, 250))
    hx, hy = 300 * 0.02 / 2, 260 * 0.02 / 2
    pts = [[-hx + 0.05, 0, 0], [hx - 0.3, 0, 0], [0, hy - 0.05, 0], [0, -hy + 0.02, 0], [hx + 1.0, 0, 0],
           [-hx - 0.4, hy + 0.3, 0], [hx - 0.9, -hy + 0.13, 0.25], [1e6, -1e6, 0]]
    q = _abi.make_queries(pts)
    got, exp = _run_both(fhp, oracle, m, p, q, 6)
    assert compare_plans(got, exp, "border") <= XYZ_TOL
    for fill_t, name in ((np.nan, "all-NaN"), (0.0, "all-blocked")):
        mm = synth.flat_map(200, 200, 0.02)
        mm.set_layer("traversability", np.full((200, 200), fill_t, np.float32))
        qq = synth.uniform_queries(64, (-1.0, 0.2), (-1.0, 1.0), seed=3)
        got, exp = _run_both(fhp, oracle, mm, p, qq, 5)
        assert compare_plans(got, exp, name) <= XYZ_TOL
    mm = synth.flat_map(200, 200, 0.02)
    qq = synth.uniform_queries(8, (-1.0, 0.2), (-1.0, 1.0), seed=3)
    got, exp = _run_both(fhp, oracle, mm, p, qq, 0)
    assert compare_plans(got, exp, "zero cycles") == 0.0 and np.all(got[2]["n_rows"] == 4)
    mm = synth.rough_map(3000, 200, 0.02, seed=2, start=(0, 0))
    qq = synth.uniform_queries(8, (-29.0, -28.0), (-1.0, 1.0), seed=3)
    got, exp = _run_both(fhp, oracle, mm, p, qq, 255)
    assert compare_plans(got, exp, "255 cycles") <= XYZ_TOL


def test_fine_grained_members(fhp, oracle):
    """checkFoothold / checkFootholdUseCentroidMethod / getFootholdMeanHeight / checkDefaultFoothold."""
    for res, n, kind in ((0.02, 300, "rough"), (0.01, 500, "rough"), (0.01, 1000, "stairs")):
        if kind == "rough":
            m = synth.rough_map(n, n, res, seed=31, start=(41, 7))
        else:
            m = synth.stairs_stones_map(n, n, res, start=(3, 999))
        fhp.set_map(m)
        p = _abi.yaml_params()
        half = n * res / 2
        rng = np.random.default_rng(5)
        c = rng.uniform(-half - 0.1, half + 0.1, size=(4000, 2))
        rc = c + rng.uniform(-0.06, 0.06, size=c.shape)
        g_out, g_valid = fhp.check_foothold(p, c, rc)
        e_out, e_valid = oracle.check_foothold(m, p, c, rc)
        assert np.array_equal(g_valid, e_valid)
        for f in ("row", "col", "kind"):
            assert np.array_equal(g_out[f], e_out[f]), f
        for f in ("x", "y", "z"):
            assert np.max(np.abs(g_out[f] - e_out[f])) <= XYZ_TOL, f
        g_out, g_case = fhp.check_centroid(p, c)
        e_out, e_case = oracle.check_centroid(m, p, c)
        assert np.array_equal(g_case, e_case)
        assert set(np.unique(e_case)) >= ({0, 1, 2, 3, 4, -1} if kind == "stairs" else {0, 1, 3})
        for f in ("row", "col"):
            assert np.array_equal(g_out[f], e_out[f]), f
        for f in ("x", "y", "z"):
            assert np.max(np.abs(g_out[f] - e_out[f])) <= XYZ_TOL, f
        assert np.array_equal(fhp.mean_height(p, c), oracle.mean_height(m, p, c))
        assert np.array_equal(fhp.check_default(p, c), oracle.check_default(m, p, c))


def test_errors_are_loud(fhp):
    from quadrupedal_foothold_planner_b200.planner import Fhp, FhpError
    p = _abi.yaml_params()
    q = _abi.make_queries([[0, 0, 0]])
    fresh = Fhp(0)
    with pytest.raises(FhpError) as e:
        fresh.plan(p, q, 4)
    assert e.value.code == _abi.FHP_ERR_NO_MAP
    m = synth.flat_map(100, 100, 0.02)
    del m.layers["elevation"]
    with pytest.raises(FhpError) as e:
        fresh.set_map(m)
    assert e.value.code == _abi.FHP_ERR_MISSING_LAYER
    fresh.set_map(synth.flat_map(100, 100, 0.02))
    with pytest.raises(FhpError) as e:
        fresh.plan(p, q, 256)
    assert e.value.code == _abi.FHP_ERR_INVALID_ARG
    with pytest.raises(FhpError) as e:  # window wider than the tables
        fresh.plan(_abi.yaml_params(search_radius=0.5), q, 2)
    assert e.value.code == _abi.FHP_ERR_UNSUPPORTED
    fresh.close()


def test_centroid_submap_one_cell_larger(fhp, oracle):
    """when a submap corner's lattice phase is within ~1e-7 cell of a boundary the window is 12x6 /
    11x7 instead of 11x6 and sticks out of the trimmed staged window -> the kernel's global-memory fallback."""
    m = synth.rough_map(400, 400, 0.02, seed=17, start=(9, 391))
    fhp.set_map(m)
    p = _abi.yaml_params()
    f32 = np.float32
    Lx, Ly = float(f32(0.1) * f32(2)), float(f32(0.1))
    pts = []
    for k in range(-150, 150, 7):
        for d in (-3e-9, -1e-9, 0.0, 1e-9, 3e-9):
            pts.append((k * 0.02 - 0.5 * Lx + d, 0.013 + 0.02 * (k % 11)))
            pts.append((0.011 + 0.02 * (k % 13), k * 0.02 - 0.5 * Ly + d))
            pts.append((k * 0.02 + 0.5 * Lx + d, k * 0.02 + 0.5 * Ly - d))
    pts = np.array(pts)
    sizes = {tuple(oracle.submap_info(m, x, y, Lx, Ly)[1][2:]) for x, y in pts}
    assert (12, 6) in sizes or (12, 7) in sizes, sizes
    assert (11, 7) in sizes or (12, 7) in sizes, sizes
    g_out, g_case = fhp.check_centroid(p, pts)
    e_out, e_case = oracle.check_centroid(m, p, pts)
    assert np.array_equal(g_case, e_case)
    for f in ("row", "col"):
        assert np.array_equal(g_out[f], e_out[f]), f
    for f in ("x", "y", "z"):
        assert np.max(np.abs(g_out[f] - e_out[f])) <= XYZ_TOL, f
    # and through the plan: stance centres chosen so the first cycle's windows hit the enlarged case
    q = _abi.make_queries([[x - 0.09, y, 0.0] for x, y in pts[:200]])
    got = fhp.plan(p, q, 3, centroid=True)
    exp = oracle.plan(m, p, q, 3, centroid=True)
    assert compare_plans(got, exp, "enlarged submap") <= XYZ_TOL


def test_default_chain_matches_oracle(fhp, oracle):
    """fhp_plan_ex: the "default" chain rows (FP.cpp:2253-2304, :1344-1348) vs the oracle, incl. a shifted start
    index and the walk preset (whose bias table is not x-symmetric, so the default chain drifts)."""
    m = synth.rough_map(400, 400, 0.02, seed=23, start=(11, 301))
    q = synth.uniform_queries(256, (-3.0, 0.5), (-3.0, 3.0), seed=24)
    for p in (_abi.yaml_params(), _abi.yaml_params(gait=_abi.FHP_GAIT_WALK, step_length=0.2)):
        fhp.set_map(m)
        nom, cen, dfl, st = fhp.plan_ex(p, q, 9)
        on, oc, od, ok, os_ = oracle.plan_ex(m, p, q, 9)
        assert compare_plans((nom, cen, st), (on, oc, os_), "plan_ex") <= XYZ_TOL
        assert compare_plans((dfl, None, st), (od, None, os_), "default chain") <= XYZ_TOL
        qi = int(np.argmax(st["n_rows"]))
        fd, cs = fhp.cycle_kpi(p, q[qi], nom[qi], int(st[qi]["n_rows"]))
        cyc = int(st[qi]["n_rows"]) // 4 - 1
        assert cyc >= 1 and np.allclose(fd, ok[qi, :cyc, 0:2], atol=1e-9) and np.allclose(cs, ok[qi, :cyc, 2:4], atol=1e-9)


def test_c3_full_size_bit_exact(fhp, oracle):
    """BASELINE config 3 at FULL size — the bench.py workload: 65,536 queries x 16 cycles on the 4000x4000 @0.02 m
    rough map with start index (1234,2777) — every row of every query against the oracle (all host threads)."""
    import os
    m, q, n = synth.config_c3()
    p = _abi.yaml_params()
    fhp.set_map(m)
    got = fhp.plan(p, q, n, centroid=True)
    exp = oracle.plan(m, p, q, n, n_threads=os.cpu_count() or 4, centroid=True)
    assert compare_plans(got, exp, "C3 full size") == 0.0
    st = got[2]
    assert 0.2 < st["success"].mean() < 0.95 and st["n_rows"].sum() > 2_000_000


def test_c2_full_size_bit_exact(fhp, oracle):
    """BASELINE config 2 at full size: walk preset, 1024 start poses x 32 cycles, 2000x2000 @0.01 m stairs/stones."""
    import os
    m, q, n = synth.config_c2()
    p = _abi.yaml_params(gait=_abi.FHP_GAIT_WALK)
    fhp.set_map(m)
    got = fhp.plan(p, q, n, centroid=True)
    exp = oracle.plan(m, p, q, n, n_threads=os.cpu_count() or 4, centroid=True)
    assert compare_plans(got, exp, "C2 full size") == 0.0


def test_row_pitch_not_multiple_of_16_bytes(fhp, oracle):
    """TMA needs the row pitch (rows*4 bytes) to be a multiple of 16: a 302-row map silently takes the LDG staging
    path on the default context too — results must not change."""
    m = synth.rough_map(302, 290, 0.02, seed=41, start=(5, 200))
    q = synth.uniform_queries(200, (-2.0, 0.5), (-2.0, 2.0), seed=42)
    p = _abi.yaml_params()
    fhp.set_map(m)
    got = fhp.plan(p, q, 8, centroid=True)
    exp = oracle.plan(m, p, q, 8, centroid=True)
    assert compare_plans(got, exp, "odd pitch") == 0.0


def test_large_map_fine_resolution(fhp, oracle):
    """C5-shaped stress (reduced): 6000x6000 @0.01 m (144 MB per layer), three layers (the extra `slope` layer is
    carried but unused), wrapped start index, 2048 queries x 32 cycles — the 32-row / 23-column box geometry."""
    import os
    m = synth.rough_map(6000, 6000, 0.01, seed=6, start=(4321, 1234), with_slope=True, pitch_m=0.35)
    q = synth.uniform_queries(2048, (-28.0, 22.0), (-28.0, 28.0), seed=7)
    p = _abi.yaml_params()
    fhp.set_map(m)
    got = fhp.plan(p, q, 32, centroid=True)
    exp = oracle.plan(m, p, q, 32, n_threads=os.cpu_count() or 4, centroid=True)
    assert compare_plans(got, exp, "large fine map") == 0.0


def test_mask_planes_lifecycle_and_fallbacks(fhp, oracle):
    """The cell-class mask planes are a cache of (map, thresholds): rebuilt on a threshold change and on a map update;
    maps with infinite cells and thrC > thrD (where a 2-bit class cannot represent the reference's raw compares) take
    the float staging paths.  Every combination must equal the oracle."""
    q = synth.uniform_queries(384, (-4.0, 1.0), (-4.0, 4.0), seed=13)
    ma = synth.rough_map(500, 500, 0.02, seed=31, start=(40, 9))
    mb = synth.rough_map(500, 500, 0.02, seed=32, start=(40, 9))
    pa, pb = _abi.yaml_params(), _abi.yaml_params(thr_default=0.8, thr_candidate=0.5)
    fhp.set_map(ma)
    for p, name in ((pa, "thr A"), (pb, "thr B"), (pa, "thr A again")):
        assert compare_plans(fhp.plan(p, q, 8), oracle.plan(ma, p, q, 8), name) <= XYZ_TOL
    fhp.set_map(mb)  # same geometry, new cell values
    assert compare_plans(fhp.plan(pa, q, 8), oracle.plan(mb, pa, q, 8), "map update") <= XYZ_TOL
    # thrC > thrD: the classes do not nest
    pc = _abi.yaml_params(thr_default=0.6, thr_candidate=0.85)
    assert compare_plans(fhp.plan(pc, q, 8), oracle.plan(mb, pc, q, 8), "thrC > thrD") <= XYZ_TOL
    # +-inf cells: -inf compares below every threshold but is not a valid (finite) cell
    mi = synth.rough_map(500, 500, 0.02, seed=33, start=(3, 471))
    t = mi.layers["traversability"].copy()
    rng = np.random.default_rng(5)
    idx = rng.integers(0, t.size, size=t.size // 25)
    t[np.unravel_index(idx[: len(idx) // 2], t.shape)] = -np.inf
    t[np.unravel_index(idx[len(idx) // 2:], t.shape)] = np.inf
    mi.set_layer("traversability", t)
    e = mi.layers["elevation"].copy()
    e[np.unravel_index(rng.integers(0, e.size, size=e.size // 50), e.shape)] = np.inf
    mi.set_layer("elevation", e)
    got, exp = _run_both(fhp, oracle, mi, pa, q, 8)
    assert compare_plans(got, exp, "inf cells") <= XYZ_TOL
    assert (got[2]["n_rows"] > 4).sum() > 20  # the case still plans something


def test_compact_response_equals_fixed_pitch_rows(fhp, oracle):
    """fhp_plan_compact returns exactly the rows fhp_plan produces, back to back, for one chunk and for several."""
    p = _abi.yaml_params()
    m = synth.rough_map(1200, 1200, 0.02, seed=71, start=(100, 900))
    fhp.set_map(m)
    for nq in (1, 37, 50000):  # 50,000 queries span more than one wave-sized chunk
        q = synth.uniform_queries(nq, (-10.0, 7.0), (-10.0, 10.0), seed=nq)
        nom, _, st = fhp.plan(p, q, 12)
        rows, offs, st2 = fhp.plan_compact(p, q, 12)
        assert np.array_equal(st, st2)
        assert offs[0] == 0 and np.array_equal(np.diff(offs), st["n_rows"]) and len(rows) == offs[-1]
        keep = np.arange(nom.shape[1])[None, :] < st["n_rows"][:, None]
        assert np.array_equal(rows, nom[keep])
    # against the oracle directly, and the capacity contract
    q = synth.uniform_queries(300, (-10.0, 7.0), (-10.0, 10.0), seed=5)
    exp = oracle.plan(m, p, q, 12)
    rows, offs, st = fhp.plan_compact(p, q, 12)
    for i in range(len(q)):
        got_i, exp_i = rows[offs[i]:offs[i + 1]], exp[0][i][: exp[2][i]["n_rows"]]
        assert len(got_i) == len(exp_i)
        for f in ("x", "y", "z"):
            assert np.max(np.abs(got_i[f] - exp_i[f]), initial=0.0) <= XYZ_TOL
        for f in ("row", "col", "foot_id", "gait_cycle_id", "kind"):
            assert np.array_equal(got_i[f], exp_i[f])
    from quadrupedal_foothold_planner_b200.planner import FhpError
    with pytest.raises(FhpError, match="capacity_rows"):
        fhp.plan_compact(p, q, 12, capacity_rows=int(offs[-1]) - 1)
    rows0, offs0, _ = fhp.plan_compact(p, q[:0], 12)
    assert len(rows0) == 0 and offs0.tolist() == [0]
    # 32-byte packed rows carry the same information
    pr, poffs, pst = fhp.plan_compact(p, q, 12, packed=True)
    assert np.array_equal(poffs, offs) and np.array_equal(pst, st) and pr.dtype.itemsize == 32
    for f in ("x", "y", "z", "foot_id", "gait_cycle_id", "kind"):
        assert np.array_equal(pr[f], rows[f])
    assert np.array_equal(pr["row"], (rows["row"] & 0xFFFF).astype(np.uint16))
    assert np.array_equal(pr["col"], (rows["col"] & 0xFFFF).astype(np.uint16))
    # 24-byte message rows: point (z as the float it is by construction), foot_id, gait_cycle_id, kind
    r24, o24, s24 = fhp.plan_compact(p, q, 12, fmt=24)
    assert np.array_equal(o24, offs) and np.array_equal(s24, st) and r24.dtype.itemsize == 24
    for f in ("x", "y", "foot_id", "gait_cycle_id", "kind"):
        assert np.array_equal(r24[f], rows[f])
    assert np.array_equal(r24["z"].astype(np.float64), rows["z"])  # lossless: every z is a float32 value
    big = synth.uniform_queries(50000, (-10.0, 7.0), (-10.0, 10.0), seed=77)  # several chunks
    rb, ob_, sb = fhp.plan_compact(p, big, 12)
    r24, o24, s24 = fhp.plan_compact(p, big, 12, fmt=24)
    assert np.array_equal(o24, ob_) and np.array_equal(s24, sb)
    assert np.array_equal(r24["x"], rb["x"]) and np.array_equal(r24["z"].astype(np.float64), rb["z"])


def test_mixed_radii_in_one_call(fhp, oracle):
    """Parameter sets with different foot / search radii in ONE fhp_plan call: planned per radii pair, rows scattered back
    in query order — every query equals the oracle's plan with its own parameter set."""
    m = synth.rough_map(600, 600, 0.02, seed=19, start=(41, 7))
    params = [_abi.yaml_params(), _abi.yaml_params(search_radius=0.14), _abi.yaml_params(foot_radius=0.031, gait=_abi.FHP_GAIT_WALK),
              _abi.yaml_params(step_length=0.2)]  # sets 0 and 3 share their radii
    q = synth.uniform_queries(400, (-4.5, 2.0), (-4.5, 4.5), seed=23)
    q["param_id"] = (np.arange(400) * 7) % 4
    exp = oracle.plan(m, params, q, 9)
    fhp.set_map(m)
    got = fhp.plan(params, q, 9, centroid=True)
    assert compare_plans(got, exp, "mixed radii") <= XYZ_TOL
    nom, cen, dfl, st = fhp.plan_ex(params, q, 9)
    assert compare_plans((nom, cen, st), exp, "mixed radii, plan_ex") <= XYZ_TOL
    from quadrupedal_foothold_planner_b200.planner import FhpError
    with pytest.raises(FhpError, match="share foot_radius"):
        fhp.plan_compact(params, q, 9)


def test_adopted_layer_rewritten_in_place_needs_map_changed(fhp, oracle):
    """fhp_set_map_device adopts caller-owned device buffers.  When the caller rewrites the traversability layer in place
    (e.g. re-broadcasts into the same tensors) it calls fhp_map_changed; the plan after that equals the oracle on the NEW
    content on every staging path (the cell-class planes of the default path are rebuilt)."""
    import torch
    m1 = synth.rough_map(500, 500, 0.02, seed=41, start=(17, 333))
    m2 = synth.rough_map(500, 500, 0.02, seed=42, start=(17, 333))
    names = ["traversability", "elevation"]
    layers = [torch.from_numpy(np.ascontiguousarray(m1.layers[n].T)).cuda() for n in names]
    fhp.set_map_device(m1, names, [t.data_ptr() for t in layers], keepalive=layers)
    q = synth.uniform_queries(256, (-4.0, 1.0), (-4.0, 4.0), seed=6)
    p = _abi.yaml_params()
    assert compare_plans(fhp.plan(p, q, 8, centroid=True), oracle.plan(m1, p, q, 8), "adopted, first content") <= XYZ_TOL
    for t, n in zip(layers, names):  # in place: same device buffers, new content
        t.copy_(torch.from_numpy(np.ascontiguousarray(m2.layers[n].T)))
    torch.cuda.synchronize()
    fhp.map_changed()
    got = fhp.plan(p, q, 8, centroid=True)
    exp2 = oracle.plan(m2, p, q, 8)
    assert compare_plans(got, exp2, "adopted, rewritten in place") <= XYZ_TOL
    assert not np.array_equal(exp2[2]["n_rows"], oracle.plan(m1, p, q, 8)[2]["n_rows"])  # the two maps really differ
    fhp.set_map(synth.flat_map(64, 64, 0.02))  # drop the adopted pointers before the tensors go away


def test_streamed_response_equals_plan_rows(fhp):
    """fhp_plan_stream (the kernel writes 20-byte points straight into pinned host memory, cycle-major blocks of 8
    queries with the successful ones packed to the front) carries exactly the rows fhp_plan returns; the masks give
    success and kind; bytes past a block's packed rows stay as the caller left them.  Partial last group, several waves,
    > 32 cycles (two mask words), a heading."""
    from quadrupedal_foothold_planner_b200.planner import Fhp, FhpError
    m = synth.rough_map(1200, 1200, 0.02, seed=71, start=(100, 900))
    fhp.set_map(m)
    for nq, cycles, p in ((1, 8, _abi.yaml_params()), (37, 12, _abi.yaml_params(gait=_abi.FHP_GAIT_WALK)),
                          (3000, 40, _abi.yaml_params()), (70000, 6, _abi.yaml_params(heading=0.7)),
                          (21, 0, _abi.yaml_params()), (50, 255, _abi.yaml_params())):
        q = synth.uniform_queries(nq, (-10.0, 3.0), (-10.0, 10.0), seed=nq)
        nom, _, st = fhp.plan(p, q, cycles)
        blocks, masks, st2 = fhp.plan_stream(p, q, cycles)
        assert np.array_equal(st, st2)
        assert blocks.shape == (cycles + 1, (nq + 7) // 8, 32) and masks.shape == (nq, max(1, (cycles + 31) // 32), 4)
        for qi in range(0, nq, max(1, nq // 257)):
            rows = nom[qi][: st[qi]["n_rows"]]
            pts, foot, cyc, kind = Fhp.stream_rows_of(blocks, masks, qi)
            assert len(pts) == len(rows)
            assert np.array_equal(pts["x"], rows["x"]) and np.array_equal(pts["y"], rows["y"])
            assert np.array_equal(pts["z"].astype(np.float64), rows["z"])
            assert np.array_equal(foot, rows["foot_id"]) and np.array_equal(cyc, rows["gait_cycle_id"])
            assert np.array_equal(kind[4:], rows["kind"][4:])
        # the host-side expander turns the blocks into exactly fhp_plan_compact24's response
        xr, xo = fhp.stream_expand(blocks, masks, st2, cycles)
        cr, co, _ = fhp.plan_compact(p, q, cycles, fmt=24)
        assert np.array_equal(xo, co) and xr.tobytes() == cr.tobytes()
        # whole batch: the four feet agree on success, popcount == successful cycles, unwritten row slots are still zero
        succ = masks["success"]
        assert np.all(succ == succ[:, :, :1])
        pop = np.array([bin(int(w)).count("1") for w in succ[:, :, 0].reshape(-1)]).reshape(nq, -1).sum(1)
        assert np.array_equal(pop, (st["n_rows"] - 4) // 4)
        groups = (nq + 7) // 8
        for g in range(cycles):
            ok = ((succ[:, g // 32, 0] >> np.uint32(g % 32)) & 1).astype(np.int64)
            per_group = np.bincount(np.arange(nq) // 8, weights=ok, minlength=groups).astype(np.int64)
            tail = np.arange(32)[None, :] >= (per_group * 4)[:, None]
            assert not blocks[g + 1][tail].view(np.uint8).any()
    q = synth.uniform_queries(8, (-10.0, 3.0), (-10.0, 10.0), seed=3)
    parr, npar = fhp._params_array(_abi.yaml_params())
    pageable = np.zeros(9 * 32 + 16, dtype=_abi.POINT20_DTYPE)
    with pytest.raises(FhpError, match="page-locked|aligned"):
        fhp.plan_stream_raw(parr, npar, q, 8, 8, pageable, np.zeros(64, np.uint32), np.zeros(8, _abi.STATUS_DTYPE))


@pytest.mark.parametrize("foot_radius", [0.031, 0.05, 0.085])
def test_wide_foot_discs(fhp, oracle, foot_radius):
    """Foot discs of 1.5 / 2.5 / 4.25 cells: the disc-membership fast path (k = 2, 3) and, for the widest, the generic
    window walk of CircleIterator (k = 5 exceeds the 8x8 membership mask)."""
    m = synth.rough_map(500, 500, 0.02, seed=81, start=(9, 333))
    # smoother acceptance so that wide discs still plan something
    t = np.clip(m.layers["traversability"] * 1.6, 0, 1).astype(np.float32)
    m.set_layer("traversability", t)
    q = synth.uniform_queries(256, (-4.0, 1.0), (-4.0, 4.0), seed=3)
    hx = 500 * 0.02 / 2
    q = np.concatenate([q, _abi.make_queries([[-hx + 0.03, 0.0, 0.0], [hx - 0.4, hx - 0.02, 0.0], [hx + 0.05, 0.0, 0.0]])])
    p = _abi.yaml_params(foot_radius=foot_radius, thr_default=0.5, thr_candidate=0.3)
    got, exp = _run_both(fhp, oracle, m, p, q, 8)
    assert compare_plans(got, exp, f"foot radius {foot_radius}") <= XYZ_TOL
    assert (got[2]["n_rows"] > 4).sum() > 10
    c = np.stack([q["x"][:64] + 0.2, q["y"][:64]], axis=1)
    assert np.array_equal(fhp.mean_height(p, c).view(np.uint32), oracle.mean_height(m, p, c).view(np.uint32))
    assert np.array_equal(fhp.check_default(p, c), oracle.check_default(m, p, c))


@pytest.mark.parametrize("nq", [5, 3000, 40000])  # 1, 1 and 8 queries per warp on a B200
def test_device_resident_plan_equals_host_plan(fhp, nq):
    """fhp_plan_device (device-resident queries and rows, caller's stream) writes exactly what fhp_plan returns."""
    import torch
    m = synth.rough_map(900, 900, 0.02, seed=91, start=(11, 800))
    fhp.set_map(m)
    p = _abi.yaml_params()
    q = synth.uniform_queries(nq, (-7.0, 4.0), (-7.0, 7.0), seed=nq)
    nom, cen, st = fhp.plan(p, q, 10, centroid=True)
    R = _abi.rows_per_query(10)
    d_q = torch.from_numpy(np.ascontiguousarray(q).view(np.uint8)).cuda()
    d_nom = torch.zeros(nq * R * 40, dtype=torch.uint8, device="cuda")
    d_cen = torch.zeros(nq * R * 40, dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(nq * 8, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()  # the uploads / fills above ran on torch's default stream; `stream` does not wait for it
    parr, npar = fhp._params_array(p)
    fhp.plan_device(parr, npar, d_q, nq, 10, d_nom, d_cen, d_st, stream=stream.cuda_stream)
    stream.synchronize()
    fhp.synchronize()
    st2 = d_st.cpu().numpy().view(_abi.STATUS_DTYPE)
    assert np.array_equal(st, st2)
    keep = np.arange(R)[None, :] < st["n_rows"][:, None]
    for got, exp in ((d_nom, nom), (d_cen, cen)):
        g = got.cpu().numpy().view(_abi.FOOTHOLD_DTYPE).reshape(nq, R)
        assert np.array_equal(g[keep], exp[keep])


def test_c4_parameter_sweep_heading_zero(fhp, oracle):
    """BASELINE config 4 at heading 0 (the slice the reference defines): 128 parameter sets ({trot, walk} x 64 step
    lengths, skew = stepLength/4) x 64 poses in ONE call, every query against the oracle."""
    m, params, q, N = synth.config_c4()
    assert len(params) == 128 and len(q) == 128 * 64
    got, exp = _run_both(fhp, oracle, m, params, q, N)
    assert compare_plans(got, exp, "C4 heading 0") <= XYZ_TOL
    # different step lengths really produce different plans
    first = got[0][:, 4]["x"].reshape(128, 64)
    assert np.unique(np.round(first[:64, 0] - first[0, 0], 9)).size > 32


def test_c5_sized_map(fhp, oracle):
    """BASELINE config 5's map size: 16000 x 16000 @ 0.01 m (2.56e8 cells per layer, 1 GB each) with a wrapped start
    index — index arithmetic, class planes (128 MB) and tensor map at the largest size the configs name."""
    rows = cols = 16000
    rng = np.random.default_rng(17)
    t = np.asfortranarray(rng.random((rows, cols), dtype=np.float32) ** 0.15)  # mostly traversable, some blocked
    t[::97, ::89] = np.nan
    e = np.zeros((rows, cols), dtype=np.float32, order="F")
    e[::2, :] = 0.125
    m = _abi.HostMap(rows, cols, 0.01, pos=(3.0, -2.0), start=(15999, 7777), layers={"traversability": t, "elevation": e})
    half = rows * 0.01 / 2
    pts = [[3.0 - half + 0.5, -2.0 + half - 0.5, 0.0], [3.0 + half - 3.0, -2.0 - half + 0.5, 0.1], [3.0, -2.0, 0.0],
           [3.0 + 70.0, -2.0 + 65.0, 0.0], [3.0 - 79.0, -2.0 - 79.5, 0.0]]
    q = np.concatenate([_abi.make_queries(pts), synth.uniform_queries(123, (3.0 - 75, 3.0 + 70), (-2.0 - 75, -2.0 + 75), seed=8)])
    p = _abi.yaml_params()
    got, exp = _run_both(fhp, oracle, m, p, q, 6)
    assert compare_plans(got, exp, "C5-sized map") <= XYZ_TOL
    assert (got[2]["n_rows"] > 4).sum() > 30
    fhp.set_map(synth.flat_map(64, 64, 0.02))  # release the 2 GB of layers for the tests that follow


@pytest.mark.parametrize("res", [0.02, 0.01])
def test_candidate_and_circle_polygon_members(fhp, oracle, res):
    """The two remaining public search members: checkCandidateFoothold (spiral alone) and
    checkCirclePolygonFoothold at arbitrary foot centres, including border, NaN-only and out-of-map discs."""
    n = 400 if res == 0.02 else 700
    m = synth.rough_map(n, n, res, seed=97, start=(n // 3, n - 7))
    t = m.layers["traversability"].copy()
    t[40:60, 40:70] = np.nan  # a NaN patch: discs that visit only invalid cells are VALID (FP.cpp:2159)
    m.set_layer("traversability", t)
    fhp.set_map(m)
    p = _abi.yaml_params()
    half = n * res / 2
    rng = np.random.default_rng(11)
    c = rng.uniform(-half - 0.1, half + 0.1, size=(3000, 2))
    c[:200] = m.pos + np.stack([half - res * rng.uniform(40, 60, 200), half - res * rng.uniform(40, 70, 200)], axis=1)
    r = c + rng.normal(0, 0.04, size=c.shape)
    got, gv = fhp.check_candidate(p, c, r)
    exp, ev = oracle.check_candidate(m, p, c, r)
    assert np.array_equal(gv, ev) and gv.sum() > 100 and (gv == 0).sum() > 20
    ok = gv == 1
    for f in ("row", "col"):
        assert np.array_equal(got[f][ok], exp[f][ok])
    for f in ("x", "y"):
        assert np.max(np.abs(got[f][ok] - exp[f][ok])) <= XYZ_TOL
    gp = fhp.check_circle_polygon(p, c, r)
    ep = oracle.check_circle_polygon(m, p, c, r)
    assert np.array_equal(gp, ep) and gp.sum() > 50 and (gp == 0).sum() > 20
    # cell-centred candidates: the member agrees with what the spiral search accepted
    cc = np.stack([got["x"][ok], got["y"][ok]], axis=1)
    assert np.all(fhp.check_circle_polygon(p, cc, r[ok]) == 1)


def test_search_rectangle_edges_on_cell_centres(fhp, oracle):
    """Rectangle edges that coincide EXACTLY with cell-centre coordinates (the half-open box: low edge inside, high
    edge outside) — the guard band of the interval estimate must hand these to the exact predicate."""
    m = synth.rough_map(400, 400, 0.02, seed=101, start=(0, 0))
    t = np.clip(m.layers["traversability"] + 0.3, 0, 1).astype(np.float32)  # many candidates accepted
    t[::3, ::2] = 0.75                                                       # default test fails often -> spiral runs
    m.set_layer("traversability", t)
    fhp.set_map(m)
    p = _abi.yaml_params()
    R = float(np.float32(0.1))
    half = 400 * 0.02 / 2
    fc = m.pos[0] + (half - 0.5 * 0.02)  # mapPosition + (0.5*length - 0.5*res)
    rects, centres = [], []
    for u in range(60, 340):
        X = fc + 0.02 * float(-u)
        for edge in (+R, -R, +0.5 * R, -0.5 * R):
            r = X - edge
            if r + edge == X:  # the edge lands exactly on the cell-centre coordinate
                for v in (100, 201, 333):
                    Y = fc + 0.02 * float(-v)
                    if abs(edge) == R:
                        rects.append((r, Y + 0.003))
                        centres.append((r + 0.013, Y - 0.004))
                    else:
                        rects.append((Y + 0.003, r))
                        centres.append((Y - 0.004, r + 0.007))
    rects, centres = np.array(rects), np.array(centres)
    assert len(rects) > 300
    got, gv = fhp.check_candidate(p, centres, rects)
    exp, ev = oracle.check_candidate(m, p, centres, rects)
    assert np.array_equal(gv, ev) and gv.sum() > 100
    ok = gv == 1
    assert np.array_equal(got["row"][ok], exp["row"][ok]) and np.array_equal(got["col"][ok], exp["col"][ok])
    assert np.array_equal(fhp.check_circle_polygon(p, centres, rects), oracle.check_circle_polygon(m, p, centres, rects))
    g2, v2 = fhp.check_foothold(p, centres, rects)
    e2, w2 = oracle.check_foothold(m, p, centres, rects)
    assert np.array_equal(v2, w2) and np.array_equal(g2["row"], e2["row"]) and np.array_equal(g2["kind"], e2["kind"])


@pytest.mark.parametrize("shape", [(8, 9, 0.02), (16, 24, 0.02), (33, 31, 0.02), (40, 64, 0.01), (5, 200, 0.02), (200, 7, 0.02)])
def test_maps_smaller_than_a_window(fhp, oracle, shape):
    """Maps narrower than the staged window / a 32-row word, with a wrapped start index and poses in and around them."""
    r, c, res = shape
    rng = np.random.default_rng(r * c)
    t = rng.random((r, c), dtype=np.float32) ** 0.2
    t[rng.random((r, c)) < 0.05] = np.nan
    e = rng.random((r, c), dtype=np.float32)
    m = _abi.HostMap(r, c, res, pos=(0.3, -0.2), start=(r // 2, c - 1), layers={"traversability": t, "elevation": e})
    hx, hy = r * res / 2, c * res / 2
    q = synth.uniform_queries(64, (0.3 - hx - 0.3, 0.3 + hx + 0.1), (-0.2 - hy - 0.2, -0.2 + hy + 0.2), seed=r + c)
    got, exp = _run_both(fhp, oracle, m, _abi.yaml_params(), q, 5)
    assert compare_plans(got, exp, f"tiny map {shape}") <= XYZ_TOL


@pytest.mark.parametrize("res", [0.02, 0.01])
def test_single_parameter_set_in_kernel_arguments_equals_table(oracle, res):
    """A call with ONE parameter set runs the plan_kernel<..., ONEPAR> variants (derived parameters as constant-bank
    operands); FHP_NO_ONEPAR=1 sends the same call through the variants that read the device table.  Same bytes from
    both, rows and streamed response, and the rows equal the oracle's — stuck and degenerate (dead) quads included: the
    rough map leaves centroid chains with two feet at (0,0,0)."""
    import os
    from quadrupedal_foothold_planner_b200.planner import Fhp
    n = 700 if res == 0.02 else 1100
    m = synth.rough_map(n, n, res, seed=17, start=(5, 9))
    half = n * res / 2
    q = synth.uniform_queries(4096, (-half + 0.8, half - 3.4), (-half + 0.8, half - 0.8), seed=3)
    p = _abi.yaml_params()
    N = 14
    out = {}
    for knob in ("0", "1"):
        old = os.environ.get("FHP_NO_ONEPAR")
        os.environ["FHP_NO_ONEPAR"] = knob
        try:
            ctx = Fhp(0)
        finally:
            if old is None:
                del os.environ["FHP_NO_ONEPAR"]
            else:
                os.environ["FHP_NO_ONEPAR"] = old
        try:
            ctx.set_map(m)
            nom, _, st = ctx.plan(p, q, N, centroid=False)
            blocks, masks, sst = ctx.plan_stream(p, q, N)
            rows, offs = ctx.stream_expand(blocks, masks, sst, N)
            out[knob] = (nom.copy(), st.copy(), rows.copy(), offs.copy(), sst.copy())
        finally:
            ctx.close()
    a, b = out["0"], out["1"]
    keep = np.arange(a[0].shape[1])[None, :] < a[1]["n_rows"][:, None]
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[4], b[4])
    assert a[0][keep].tobytes() == b[0][keep].tobytes()
    assert a[2].tobytes() == b[2].tobytes() and np.array_equal(a[3], b[3])
    assert int((a[1]["gait_cycles_succeed"] < N).sum()) > 100  # not only clean runs
    exp_nom, _, exp_st = oracle.plan(m, p, q, N, centroid=False)
    assert compare_plans((a[0], None, a[1]), (exp_nom, None, exp_st), "one parameter set") == 0.0  # bit-exact


@pytest.mark.parametrize("mode", [_abi.FHP_SPIRAL_WRAP_CORRECT, _abi.FHP_SPIRAL_FAITHFUL])
@pytest.mark.parametrize("start", [(0, 0), (37, 211)])
def test_non_finite_and_far_away_poses(fhp, oracle, start, mode):
    """Poses that are NaN, infinite, astronomically large or simply off the map, mixed with ordinary ones in the same
    warps: the index policy of SURVEY App. C (a non-finite / out-of-range coordinate converts to -2^30) makes every such
    foothold fail — except in the bug-compatible spiral mode on a shifted start index, where upstream wraps that index
    into the buffer — and the kernel's dead-quad shortcut must not change any of it."""
    m = synth.rough_map(400, 380, 0.02, seed=23, start=start)
    good = synth.uniform_queries(96, (-2.5, 0.5), (-2.5, 2.5), seed=8)
    odd = [[np.nan, 0.0, 0.0], [0.0, np.nan, 0.0], [np.inf, 0.0, 0.0], [0.0, -np.inf, 0.0], [1e300, 0.0, 0.0],
           [0.0, -1e300, 0.0], [1e12, 1e12, 0.0], [-3.0e7, 2.0, 0.0], [50.0, 0.0, 0.0], [0.0, -50.0, 0.0],
           [3.99, 0.0, 0.0], [-4.2, 3.7, 0.0], [0.0, 0.0, np.nan], [0.0, 0.0, np.inf]]
    q = np.concatenate([good[:40], _abi.make_queries(odd[:7]), good[40:70], _abi.make_queries(odd[7:]), good[70:]])
    p = _abi.yaml_params(spiral_mode=mode)
    fhp.set_map(m)
    got = fhp.plan(p, q, 9, centroid=True)
    exp = oracle.plan(m, p, q, 9)
    gs, es = got[2], exp[2]
    for f in ("success", "gait_cycles_succeed", "n_rows"):
        assert np.array_equal(gs[f], es[f]), f
    owned = np.arange(got[0].shape[1])[None, :] < es["n_rows"][:, None]
    for k in (0, 1):
        for f in ("row", "col", "foot_id", "gait_cycle_id", "kind"):
            assert np.array_equal(got[k][f][owned], exp[k][f][owned]), f
        for f in ("x", "y", "z"):  # NaN == NaN here: the stance rows of a NaN pose carry it
            assert np.array_equal(got[k][f][owned], exp[k][f][owned], equal_nan=True), f
    assert int(es["gait_cycles_succeed"][:40].max()) > 0  # the ordinary queries next to them still plan


def test_fuzz_kernel_vs_oracle(fhp, oracle):
    """Random geometry the fixed tests do not visit — resolutions 0.01 / 0.02 / 0.025 / 0.04 m (windows of 23 / 11 / 9 /
    7 columns: specialised and generic kernels), random map sizes and start indices, thresholds, step lengths, leg order,
    gait, spiral variant, foot radius — every query against the oracle, bit for bit.  Fixed seed: a regression test,
    not a campaign."""
    from quadrupedal_foothold_planner_b200.planner import FhpError
    import os
    n_cases = int(os.environ.get("FHP_FUZZ_CASES", "28"))  # (campaigns by hand: FHP_FUZZ_SEED / FHP_FUZZ_CASES)
    rng = np.random.default_rng(int(os.environ.get("FHP_FUZZ_SEED", "20261017")))
    ran = 0
    for case in range(n_cases):
        res = float(rng.choice([0.02, 0.01, 0.025, 0.04]))
        rows, cols = int(rng.integers(150, 420)), int(rng.integers(150, 420))
        if res == 0.01:
            rows += 200
            cols += 200
        start = (0, 0) if rng.random() < 0.3 else (int(rng.integers(0, rows)), int(rng.integers(0, cols)))
        m = synth.rough_map(rows, cols, res, seed=int(rng.integers(1, 1 << 30)), start=start)
        kw = {"spiral_mode": _abi.FHP_SPIRAL_FAITHFUL if rng.random() < 0.5 else _abi.FHP_SPIRAL_WRAP_CORRECT}
        if rng.random() < 0.5:
            kw["step_length"] = float(np.float32(rng.uniform(0.10, 0.26)))
            kw["skew"] = float(np.float32(kw["step_length"] / 4))
        if rng.random() < 0.3:
            kw["rf_first"] = 1
        if rng.random() < 0.3:
            kw["gait"] = _abi.FHP_GAIT_WALK
        if rng.random() < 0.3:
            kw["thr_default"] = float(np.float32(rng.uniform(0.75, 0.95)))
            kw["thr_candidate"] = float(np.float32(rng.uniform(0.5, 0.74)))
        if rng.random() < 0.25:
            kw["foot_radius"] = float(np.float32(rng.uniform(0.015, 0.045)))
        p = _abi.yaml_params(**kw)
        hx, hy = rows * res / 2, cols * res / 2
        q = synth.uniform_queries(96, (-hx + 0.3, hx - 1.2), (-hy + 0.3, hy - 0.3), seed=int(rng.integers(1, 1 << 30)))
        cycles = int(rng.integers(1, 9))
        fhp.set_map(m)
        try:
            got = fhp.plan(p, q, cycles, centroid=True)
        except FhpError as e:  # a foot-disc offset too close to the radius to classify: refused, never guessed
            assert e.code == _abi.FHP_ERR_UNSUPPORTED, e
            continue
        exp = oracle.plan(m, p, q, cycles)
        assert compare_plans(got, exp, f"fuzz case {case}: res {res} {rows}x{cols} start {start} {kw} cycles {cycles}") == 0.0
        ran += 1
    assert ran >= n_cases * 2 // 3
