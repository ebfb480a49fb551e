"""Shared comparison helpers: product rows vs oracle rows."""
import numpy as np

XYZ_TOL = 1e-5  # metres — BASELINE.json north_star: "foothold xyz within 1e-5 m"


def compare_plans(got, exp, what=""):
    """got/exp = (nominal rows [nq,R], centroid rows or None, status [nq]).  Integer fields and status bit-exact,
    xyz within XYZ_TOL over the rows each query owns.  Returns the max |xyz| difference seen."""
    gn, gc, gs = got
    en, ec, es = exp
    for f in ("success", "gait_cycles", "gait_cycles_succeed", "n_rows"):
        bad = np.nonzero(gs[f] != es[f])[0]
        assert bad.size == 0, f"{what}: status.{f} differs for queries {bad[:8]} got {gs[f][bad[:8]]} exp {es[f][bad[:8]]}"
    assert np.all((gs["reserved"] & 0x7F) == 0), f"{what}: kernel error bits set"
    bad = np.nonzero(gs["reserved"] != es["reserved"])[0]  # FHP_STATUS_OPT_SUBMAP_FAIL (0x80)
    assert bad.size == 0, f"{what}: status flags differ for queries {bad[:8]} got {gs['reserved'][bad[:8]]} exp {es['reserved'][bad[:8]]}"
    R = gn.shape[1]
    owned = np.arange(R)[None, :] < es["n_rows"][:, None]
    maxd = 0.0
    for name, g, e in (("nominal", gn, en), ("centroid", gc, ec)):
        if g is None or e is None:
            continue
        for f in ("row", "col", "foot_id", "gait_cycle_id", "kind"):
            neq = (g[f] != e[f]) & owned
            if neq.any():
                qi, ri = np.argwhere(neq)[0]
                raise AssertionError(f"{what}: {name}.{f} differs at query {qi} row {ri}: got {g[qi, ri]} exp {e[qi, ri]}")
        for f in ("x", "y", "z"):
            d = np.abs(np.where(owned, g[f], 0.0) - np.where(owned, e[f], 0.0))  # rows beyond n_rows are unspecified
            assert not np.isnan(d).any(), f"{what}: NaN in {name}.{f}"
            m = float(d.max()) if d.size else 0.0
            if m > XYZ_TOL:
                qi, ri = np.unravel_index(np.argmax(d), d.shape)
                raise AssertionError(f"{what}: {name}.{f} differs by {m} at query {qi} row {ri}: got {g[qi, ri]} exp {e[qi, ri]}")
            maxd = max(maxd, m)
    return maxd
