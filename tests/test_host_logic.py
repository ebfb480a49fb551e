"""Host-side logic that needs neither GPU nor oracle: synthetic inputs, layouts, response assembly, sharding."""
import numpy as np

from quadrupedal_foothold_planner_b200 import _abi, dist, planner, synth


def test_synth_is_deterministic_and_layout_is_column_major():
    a = synth.rough_map(120, 90, 0.02, seed=3, start=(7, 11))
    b = synth.rough_map(120, 90, 0.02, seed=3, start=(7, 11))
    for k in a.layers:
        assert np.array_equal(a.layers[k], b.layers[k], equal_nan=True)
        assert a.layers[k].flags["F_CONTIGUOUS"] and a.layers[k].dtype == np.float32
    c = synth.rough_map(120, 90, 0.02, seed=4, start=(7, 11))
    assert not np.array_equal(a.layers["elevation"], c.layers["elevation"], equal_nan=True)
    # element (r,c) at c*rows + r
    flat = a.layers["elevation"].ravel(order="K")
    assert np.array_equal(flat[5 * 120 + 3], a.layers["elevation"][3, 5], equal_nan=True)
    t = a.layers["traversability"]
    assert 0.15 < np.nanmean(t < 0.9) < 0.40 and 0.03 < np.nanmean(t < 0.7) < 0.15 and 0.005 < np.isnan(t).mean() < 0.02


def test_start_index_relayout_roundtrip():
    a = synth.rough_map(60, 50, 0.02, seed=1, start=(13, 44))
    b = a.relaid_start0()
    assert b.start == (0, 0)
    # unwrapped (u0,u1) lives at buffer ((u0+s0)%rows, (u1+s1)%cols)
    for u0, u1 in ((0, 0), (5, 7), (59, 49), (47, 6)):
        assert np.array_equal(b.layers["elevation"][u0, u1], a.layers["elevation"][(u0 + 13) % 60, (u1 + 44) % 50],
                              equal_nan=True)


def test_configs_match_baseline_shapes():
    m, q, n = synth.config_c1()
    assert (m.rows, m.cols, m.resolution, len(q), n) == (500, 500, 0.02, 1, 8)
    m, q, n = synth.config_c2(64)
    assert (m.rows, m.cols, m.resolution, len(q), n) == (2000, 2000, 0.01, 64, 32)
    q = synth.uniform_queries(1000, (-39.0, 35.5), (-39.0, 39.0), seed=4)
    assert q["x"].min() >= -39.0 and q["x"].max() < 35.5 and len(np.unique(q["x"])) == 1000


def test_rows_to_msg_and_struct_views():
    rows = np.zeros(12, dtype=_abi.FOOTHOLD_DTYPE)
    rows["x"] = np.arange(12)
    rows["foot_id"] = np.arange(12) % 4
    rows["gait_cycle_id"] = np.arange(12) // 4
    st = np.zeros(1, dtype=_abi.STATUS_DTYPE)
    st[0] = (1, 5, 2, 0, 8)
    msg = planner.rows_to_msg(rows, st[0])
    assert msg.success and msg.gait_cycles == 5 and msg.gait_cycles_succeed == 2 and len(msg.footholds) == 8
    assert [f.foot_id for f in msg.footholds] == [0, 1, 2, 3, 0, 1, 2, 3] and msg.footholds[5].point.x == 5.0
    assert _abi.rows_per_query(16) == 68


def test_shard_bounds_partition():
    for n, w in ((65536, 8), (1000, 3), (5, 8), (0, 4), (7, 7)):
        spans = [dist.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and all(lo <= hi for lo, hi in spans)


def test_host_geometry_members_match_the_oracle(oracle):
    """fhp_initial_stance / fhp_search_geometry (initialize + setFirstGait, getPolygonCenter, getFootholdSearchGridMap,
    getDefaultFootholdNext, getSearchPolygon) are host-side pure functions of libfhp: bit-identical to the oracle."""
    import numpy as np
    from quadrupedal_foothold_planner_b200 import _abi, planner
    rng = np.random.default_rng(3)
    for p in (_abi.yaml_params(), _abi.yaml_params(rf_first=1, step_length=0.23, skew=0.055),
              _abi.yaml_params(gait=_abi.FHP_GAIT_WALK, search_radius=0.13)):
        pose = (float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5)), float(rng.uniform(-1, 1)))
        init, first = planner.initial_stance(p, pose)
        o_init, o_bias, o_first = oracle.initial_geometry(p, pose)
        assert np.array_equal(init[:, :2], o_init) and np.array_equal(first[:, :2], o_first)
        assert np.all(init[:, 2] == pose[2]) and np.all(first[:, 2] == pose[2])
        for adj in (0.0, -0.007, -0.049):
            cur = first + rng.normal(0, 0.03, size=(4, 3))
            c, nxt, poly = planner.search_geometry(p, pose, adj, cur)
            oc, onxt, opoly = oracle.search_geometry(p, pose, adj, cur)
            assert np.array_equal(c[:2], oc) and np.array_equal(nxt, onxt) and np.array_equal(poly, opoly)
            assert c[2] == (cur[0, 2] + cur[1, 2] + cur[2, 2] + cur[3, 2]) / 4.0
