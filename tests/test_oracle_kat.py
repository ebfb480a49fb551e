"""Known-answer and property tests of the CPU oracle (oracle/) — the restatement of the reference's CPU planner.

The reference ships no tests or golden vectors, so these pin the oracle against hand-derivable answers:
the flat-map closed form, grid_map_core iterator facts, and designed maps for every centroid case."""
import numpy as np
import pytest

from quadrupedal_foothold_planner_b200 import _abi, synth

f32 = np.float32
R = float(f32(0.1))   # searchRadius_ promoted
r = float(f32(0.02))  # footRadius_ promoted


def test_c1_closed_form(oracle):
    """flat map => foot f at cycle g = (-4 - 0.09f + 0.18f*(g+1) + bias_x[f], bias_y[f] - 0.007 g, 0.01f)."""
    m, q, N = synth.config_c1()
    nom, cen, st = oracle.plan(m, _abi.yaml_params(), q, N)
    assert tuple(st[0])[:3] == (1, 8, 8) and st[0]["n_rows"] == 36
    L = float(f32(0.4387)); W = float(f32(0.175) + f32(0.037) * f32(2)); s = float(f32(0.04))
    sl = float(f32(0.18)); hs = float(f32(0.18) / f32(2))
    bx = [0.5 * L - s, -0.5 * L + s, -0.5 * L - s, 0.5 * L + s]
    by = [-0.5 * W, -0.5 * W, 0.5 * W, 0.5 * W]
    sx = [0.5 * L, -0.5 * L, -0.5 * L, 0.5 * L]
    for f in range(4):  # stance rows, FP.cpp:685-699
        assert nom[0][f]["x"] == sx[f] - 4.0 and nom[0][f]["y"] == by[f] and nom[0][f]["kind"] == 0
        assert nom[0][f]["foot_id"] == f and nom[0][f]["gait_cycle_id"] == 0
    for g in range(8):
        for f in range(4):
            row = nom[0][4 + 4 * g + f]
            assert abs(row["x"] - (-4 - hs + sl * (g + 1) + bx[f])) < 1e-12
            assert abs(row["y"] - (by[f] - 0.007 * g)) < 1e-12
            assert row["z"] == float(f32(0.01))
            assert (row["foot_id"], row["gait_cycle_id"], row["kind"]) == (f, g, 1)
    # both chains coincide on a flat map
    for k in ("x", "y", "z"):
        assert np.array_equal(nom[k], cen[k])


def test_float_promotions(oracle):
    """the float-typed rosparams promote to these doubles."""
    init, bias, first = oracle.initial_geometry(_abi.yaml_params())
    assert init[0, 0] == 0.21934999525547028 and init[0, 1] == -0.12449999898672104
    assert bias[0, 0] == 0.21934999525547028 - 0.03999999910593033
    assert bias[3, 0] == 0.21934999525547028 + 0.03999999910593033
    assert first[0, 0] == init[0, 0] - float(f32(0.09))
    _, bias_rf, _ = oracle.initial_geometry(_abi.yaml_params(rf_first=1))
    assert bias_rf[0, 0] == 0.21934999525547028 + 0.03999999910593033  # RF_FIRST flips the skew sign, FP.cpp:403-411


@pytest.mark.parametrize("res,nrings,total,inner", [(0.02, 6, 145, 69), (0.01, 11, 437, 305)])
def test_spiral_ring_structure(oracle, res, nrings, total, inner):
    """nRings = ceil(0.1f/res); every ring d visits the lattice points with floor(|offset|) == d;
    totals incl. centre 145 / 437 before the distance filter; rings <= nRings-2 are never filtered (69 / 305)."""
    assert int(np.ceil(R / res)) == nrings
    pts = [(a, b) for a in range(-nrings - 1, nrings + 2) for b in range(-nrings - 1, nrings + 2)
           if int(np.sqrt(a * a + b * b)) <= nrings]
    assert len(pts) == total
    assert sum(1 for a, b in pts if int(np.sqrt(a * a + b * b)) <= nrings - 2) == inner
    n = 400
    m = synth.flat_map(n, n, res)
    # centre exactly on a cell centre: ring nRings-1 keeps the lattice points within R, ring nRings none
    ok, c = oracle.get_position(m, 200, 200)
    cells = oracle.spiral_cells(m, c[0], c[1], R)
    off = cells - np.array([200, 200])
    ring = np.floor(np.sqrt((off ** 2).sum(1))).astype(int)
    assert np.all(np.diff(ring) >= 0), "rings are visited in increasing order"
    assert np.array_equal(off[0], [0, 0])
    assert [tuple(o) for o in off[1:9]] == [(1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0)]
    assert (ring <= nrings - 2).sum() == inner
    expect = sum(1 for a, b in pts if int(np.sqrt(a * a + b * b)) <= nrings - 2 or
                 (a * a + b * b) * res * res <= R * R * (1 + 1e-12))
    assert len(cells) == expect and len({tuple(x) for x in cells}) == len(cells)


def test_spiral_modes_and_start_index(oracle):
    """faithful == wrap_correct when the start index is (0,0); wrap_correct is start-index invariant (App. C.3)."""
    m0 = synth.flat_map(300, 260, 0.02)
    rng = np.random.default_rng(1)
    for _ in range(20):
        cx, cy = rng.uniform(-2.5, 2.5), rng.uniform(-2.2, 2.2)
        a = oracle.spiral_cells(m0, cx, cy, R, wrap_correct=False)
        b = oracle.spiral_cells(m0, cx, cy, R, wrap_correct=True)
        assert np.array_equal(a, b)
        ms = synth.flat_map(300, 260, 0.02, start=(77, 201))
        c = oracle.spiral_cells(ms, cx, cy, R, wrap_correct=True)
        unwrapped = (c - np.array([77, 201])) % np.array([300, 260])
        assert np.array_equal(unwrapped, a)
    # near the border the spiral is clipped, never wrapped
    a = oracle.spiral_cells(m0, 2.99, 0.0, R, wrap_correct=True)
    assert a[:, 0].min() == 0 and len(a) < 81


def test_circle_iterator(oracle):
    """App. A.4: foot disc r = 0.02f around a cell centre is that single cell @0.02 m and the 3x3 block @0.01 m
    (the 0.02-distance neighbours are excluded because 0.0199999996 < 0.02); order is row-major (i1 fastest)."""
    m2 = synth.flat_map(100, 100, 0.02, start=(13, 91))
    ok, c = oracle.get_position(m2, 40, 50)
    assert [tuple(x) for x in oracle.circle_cells(m2, c[0], c[1], r)] == [(40, 50)]
    m1 = synth.flat_map(100, 100, 0.01)
    ok, c = oracle.get_position(m1, 40, 50)
    cells = [tuple(x) for x in oracle.circle_cells(m1, c[0], c[1], r)]
    assert cells == [(a, b) for a in (39, 40, 41) for b in (49, 50, 51)]
    # a disc completely off the map visits nothing
    assert len(oracle.circle_cells(m1, 5.0, 5.0, r)) == 0
    # continuous centre: all visited cells are within r, all unvisited neighbours are not
    rng = np.random.default_rng(3)
    for _ in range(50):
        cx, cy = rng.uniform(-0.4, 0.4, 2)
        cells = oracle.circle_cells(m1, cx, cy, r)
        for a in range(0, 100):
            pass
        pos = np.array([oracle.get_position(m1, int(a), int(b))[1] for a, b in cells])
        assert np.all(((pos - [cx, cy]) ** 2).sum(1) <= r * r)
        assert 9 <= len(cells) <= 16


def test_submap_window_sizes(oracle):
    """App. A.5: getSubmap(p, (0.2f, 0.1f)) is 11x6 cells @0.02 m and 21x11 @0.01 m for generic p."""
    for res, want in ((0.02, (11, 6)), (0.01, (21, 11))):
        m = synth.flat_map(400, 400, res, start=(5, 9))
        rng = np.random.default_rng(2)
        seen = set()
        for _ in range(200):
            px, py = rng.uniform(-1.0, 1.0, 2)
            ok, out, pl = oracle.submap_info(m, px, py, float(f32(0.1) * f32(2)), float(f32(0.1)))
            assert ok
            seen.add((int(out[2]), int(out[3])))
            assert abs(pl[2] - out[2] * res) < 1e-12
        assert want in seen and all(want[0] <= a <= want[0] + 1 and want[1] <= b <= want[1] + 1 for a, b in seen)
    # centre outside the map -> the requested position is not inside the bounded submap -> failure
    m = synth.flat_map(100, 100, 0.02)
    assert not oracle.submap_info(m, 1.2, 0.0, 0.2, 0.1)[0]


def test_index_position_roundtrip_and_frame(oracle):
    """App. A.2: x,y decrease with the index; position -> index -> position is the identity on cell centres; the
    start index only relabels storage."""
    m = synth.flat_map(50, 40, 0.02, pos=(1.0, -2.0), start=(7, 33))
    for i0, i1 in ((0, 0), (7, 33), (49, 39), (6, 32), (20, 5)):
        ok, p = oracle.get_position(m, i0, i1)
        assert ok
        ok2, idx = oracle.get_index(m, p[0], p[1])
        assert ok2 and idx == (i0, i1)
    ok, p_first = oracle.get_position(m, 7, 33)  # unwrapped (0,0) = top-left = max x, max y
    assert abs(p_first[0] - (1.0 + 0.5 - 0.01)) < 1e-12 and abs(p_first[1] - (-2.0 + 0.4 - 0.01)) < 1e-12
    assert not oracle.get_index(m, 1.0 + 0.5 + 1e-9, -2.0)[0]
    assert not oracle.get_position(m, 50, 0)[0]


def test_polygon_is_half_open_box(oracle):
    """App. A.7: PNPOLY on the reference's rectangle == x-R <= px < x+R and y-R/2 <= py < y+R/2."""
    x, y = 0.3, -0.2
    v = [(x + R, y + 0.5 * R), (x + R, y - 0.5 * R), (x - R, y - 0.5 * R), (x - R, y + 0.5 * R)]
    rng = np.random.default_rng(4)
    pts = rng.uniform(-0.2, 0.2, size=(2000, 2)) + [x, y]
    edge = [(x - R, y), (x + R, y), (x, y - 0.5 * R), (x, y + 0.5 * R), (x - R, y - 0.5 * R), (x + R, y + 0.5 * R)]
    for px, py in list(pts) + edge:
        want = (x - R <= px < x + R) and (y - 0.5 * R <= py < y + 0.5 * R)
        assert oracle.polygon_inside(v, px, py) == want


def _stripe_map(blocked_rows, n=200, res=0.02, value=0.0):
    """trav = 1 everywhere except whole map rows (x index) in blocked_rows."""
    m = synth.flat_map(n, n, res)
    t = np.ones((n, n), np.float32)
    for rr in blocked_rows:
        t[rr, :] = value
    m.set_layer("traversability", t)
    e = np.tile(np.arange(n, dtype=np.float32)[:, None] * 0.001, (1, n))
    m.set_layer("elevation", e)
    return m


def test_centroid_cases(oracle):
    """FP.cpp:1777 / :1843 / :1944: the three cases + whole-valid + no-case on designed stripe maps.
    Query point = centre of cell (100,100): the 2R x R window is rows 95..105 (11), cols 98..103 (6)
    (top-left = p + half length, and a corner exactly on a cell edge belongs to the lower-index cell)."""
    p = _abi.yaml_params()
    n = 200

    def run(rows):
        m = _stripe_map(rows)
        ok, c = oracle.get_position(m, 100, 100)
        out, case = oracle.check_centroid(m, p, [c])
        return m, c, out[0], int(case[0])

    m, c, out, case = run([])
    assert case == 0 and (out["x"], out["y"]) == c and out["z"] == float(f32(f32(0.1) + 0.01))  # elev(100)=0.1
    ok, info, _ = oracle.submap_info(m, c[0], c[1], float(f32(0.1) * f32(2)), float(f32(0.1)))
    top = int(info[0]); nrows = int(info[2]); ncols = int(info[3]); left = int(info[1])
    assert (nrows, ncols) in ((11, 6), (12, 6), (11, 7), (12, 7))
    bottom = nrows - 1
    # case 1: edge rows start at the top of the window -> newRow = floor((maxRow+bottom+1)/2), newCol = floor(ncols/2)
    m, c, out, case = run([top, top + 1, top + 2])
    assert case == 1
    new_row = (2 + bottom + 1) // 2
    assert (out["row"], out["col"]) == (top + new_row, left + ncols // 2)
    # case 3: edge rows end at the bottom -> newRow = ceil(minRow/2), newCol = floor((ncols-1)/2)
    m, c, out, case = run([top + bottom - 1, top + bottom])
    assert case == 3
    assert (out["row"], out["col"]) == (top + (bottom - 1 + 1) // 2, left + (ncols - 1) // 2)
    # case 2, upper band larger (minRow - top >= bottom - maxRow) -> ceil(minRow/2)
    m, c, out, case = run([top + 7, top + 8])
    assert case == 2 and out["row"] == top + 4
    # case 2, lower band larger -> floor((maxRow+bottom)/2)
    m, c, out, case = run([top + 2, top + 3])
    assert case == 2 and out["row"] == top + (3 + bottom) // 2
    # no case: edge rows at both the top and the bottom -> result stays (0,0,0)
    m, c, out, case = run([top, top + bottom])
    assert case == 4 and (out["x"], out["y"], out["z"]) == (0.0, 0.0, 0.0)
    # a row is an edge only if MORE than half of its cells are below the threshold
    m = _stripe_map([])
    t = m.layers["traversability"].copy()
    t[top + 5, left:left + ncols // 2] = 0.0  # exactly half (or less) -> not an edge -> "no edge rows" => case 1 with maxRow=0
    m.set_layer("traversability", t)
    out, case = oracle.check_centroid(m, p, [c])
    assert int(case[0]) == 1 and out[0]["row"] == top + (0 + bottom + 1) // 2
    # NaN cells never count as below the threshold
    t[:] = np.nan
    m.set_layer("traversability", t)
    out, case = oracle.check_centroid(m, p, [c])
    assert int(case[0]) == 0


def test_check_foothold_semantics(oracle):
    """FP.cpp:2001-2163: default disc accepted; else first valid cell in spiral order inside the rectangle; z is
    sampled at the CENTRE, not at the chosen cell; NaN cells are skipped; all-blocked fails."""
    p = _abi.yaml_params()
    m = synth.flat_map(200, 200, 0.02)
    e = np.tile(np.arange(200, dtype=np.float32)[:, None] * 0.001, (1, 200))
    m.set_layer("elevation", e)
    ok, c = oracle.get_position(m, 100, 100)
    out, valid = oracle.check_foothold(m, p, [c], [c])
    assert valid[0] == 1 and out[0]["kind"] == 1 and (out[0]["row"], out[0]["col"]) == (100, 100)
    # block the centre cell below the DEFAULT threshold only (0.8): candidate threshold 0.7 still accepts it
    t = np.ones((200, 200), np.float32)
    t[100, 100] = 0.8
    m.set_layer("traversability", t)
    out, valid = oracle.check_foothold(m, p, [c], [c])
    assert valid[0] == 1 and out[0]["kind"] == 2 and (out[0]["row"], out[0]["col"]) == (100, 100)
    # block it below both: the first ring-1 cell in spiral order wins: offset (+1,-1)
    t[100, 100] = 0.1
    m.set_layer("traversability", t)
    out, valid = oracle.check_foothold(m, p, [c], [c])
    assert valid[0] == 1 and (out[0]["row"], out[0]["col"]) == (101, 99)
    assert out[0]["z"] == float(f32(f32(0.1) + 0.01)), "height at the centre cell (100), not at the chosen cell (101)"
    # that cell outside the rectangle (rectangle shifted to -x by 2 cells... i.e. towards larger row index): still in;
    # shift the rectangle so only rows <= 100 are inside: candidates with row 101 are rejected
    rc = (c[0] + 0.1 + 0.011, c[1])  # box x in [c+0.011, c+0.211): rows <= 99
    out, valid = oracle.check_foothold(m, p, [c], [rc])
    assert valid[0] == 1 and out[0]["row"] <= 99
    # NaN centre cell: the default test skips it and passes (FP.cpp:2055, :2078)
    t[:] = 1.0
    t[100, 100] = np.nan
    m.set_layer("traversability", t)
    out, valid = oracle.check_foothold(m, p, [c], [c])
    assert valid[0] == 1 and out[0]["kind"] == 1
    t[:] = 0.0
    m.set_layer("traversability", t)
    out, valid = oracle.check_foothold(m, p, [c], [c])
    assert valid[0] == 0 and out[0]["z"] == 0.0 and out[0]["kind"] == 2


def test_failed_cycle_semantics(oracle):
    """FP.cpp:1571-1578: a failed cycle appends nothing, does not advance the chains, does not break; the lateral
    drift still accumulates; success = outcome of the LAST cycle; gait_cycles_succeed = index+1 of the last success."""
    p = _abi.yaml_params()
    m = synth.flat_map(500, 500, 0.02)
    t = np.ones((500, 500), np.float32)
    # a fully blocked band across the RF/LF path: x in [-3.45, -3.05] (first cycle front feet land at ~-3.65/-3.73)
    x, _ = synth.cell_centres(500, 500, 0.02)
    t[(x > -3.55) & (x < -3.15), :] = 0.0
    m.set_layer("traversability", t)
    q = _abi.make_queries([[-4.0, 0.0, 0.0]])
    nom, cen, st = oracle.plan(m, p, q, 6)
    s = st[0]
    assert s["gait_cycles"] == 6 and s["n_rows"] == 4 + 4 * s["gait_cycles_succeed"] or s["success"] == 0
    rows = nom[0][: s["n_rows"]]
    ids = rows["gait_cycle_id"][4:]
    assert np.all(np.diff(ids.astype(int)) >= 0)
    # once blocked the chains never advance: every later cycle fails too, success is false, succeed keeps the last index
    assert s["success"] == 0 and s["gait_cycles_succeed"] < 6
    assert s["n_rows"] == 4 + 4 * len(np.unique(ids))


def test_oracle_start_index_invariance(oracle):
    m = synth.rough_map(300, 300, 0.02, seed=5, start=(101, 17))
    q = synth.uniform_queries(64, (-2.0, 0.5), (-2.0, 2.0), seed=6)
    p = _abi.yaml_params()
    a = oracle.plan(m, p, q, 8)
    b = oracle.plan(m.relaid_start0(), p, q, 8)
    for k in ("x", "y", "z", "foot_id", "gait_cycle_id", "kind"):
        assert np.array_equal(a[0][k], b[0][k]) and np.array_equal(a[1][k], b[1][k])
    assert np.array_equal(a[2], b[2])
    # and faithful == wrap_correct on the start-0 layout
    c = oracle.plan(m.relaid_start0(), _abi.yaml_params(spiral_mode=_abi.FHP_SPIRAL_FAITHFUL), q, 8)
    assert np.array_equal(b[0], c[0]) and np.array_equal(b[2], c[2])


def test_oracle_threads_deterministic(oracle):
    m = synth.rough_map(300, 300, 0.02, seed=5, start=(0, 0))
    q = synth.uniform_queries(200, (-2.0, 0.5), (-2.0, 2.0), seed=6)
    p = _abi.yaml_params()
    a = oracle.plan(m, p, q, 6, n_threads=1)
    b = oracle.plan(m, p, q, 6, n_threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
