"""Synthetic grid_maps and query sets for the BASELINE.json configs.

Everything is generated from a counter-based integer hash (splitmix64 of (seed, i, j)), so a map is a pure
function of its parameters: reproducible on any host, no files.  Content is defined on the UNWRAPPED cell grid
(cell (u0,u1) has centre x = pos_x + len_x/2 - res/2 - res*u0, same for y) and then rolled into the circular
buffer by the start index, exactly how grid_map::GridMap::move leaves a map laid out.
"""
from __future__ import annotations

import numpy as np

from ._abi import HostMap, make_queries

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def hash2(seed: int, i: np.ndarray, j: np.ndarray) -> np.ndarray:
    """uint64 hash of (seed, i, j); i, j broadcastable integer arrays."""
    with np.errstate(over="ignore"):
        a = splitmix64(np.asarray(i, dtype=np.int64).astype(np.uint64) + np.uint64(seed) * np.uint64(0x632BE59BD9B4E019))
        b = splitmix64(a ^ (np.asarray(j, dtype=np.int64).astype(np.uint64) * np.uint64(0xD1B54A32D192ED03)))
    return b


def u01(h: np.ndarray) -> np.ndarray:
    """uniform [0,1) float64 from the top 53 bits"""
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def cell_centres(rows, cols, res, pos=(0.0, 0.0)):
    """x[u0], y[u1] of the unwrapped grid (plain numpy arithmetic; used for map CONTENT only)."""
    x = pos[0] + rows * res / 2 - res / 2 - res * np.arange(rows)
    y = pos[1] + cols * res / 2 - res / 2 - res * np.arange(cols)
    return x, y


def _roll_to_buffer(a: np.ndarray, start) -> np.ndarray:
    # buffer[(u0+s0)%rows, (u1+s1)%cols] = unwrapped[u0,u1]
    return np.roll(a, shift=(start[0], start[1]), axis=(0, 1))


def _value_noise(seed, rows, cols, pitch_cells):
    """Bilinear value noise in [0,1): lattice values at every `pitch_cells` cells."""
    gi = np.arange(rows) / pitch_cells
    gj = np.arange(cols) / pitch_cells
    i0 = np.floor(gi).astype(np.int64)
    j0 = np.floor(gj).astype(np.int64)
    fi = (gi - i0)[:, None]
    fj = (gj - j0)[None, :]
    # smoothstep weights
    fi = fi * fi * (3 - 2 * fi)
    fj = fj * fj * (3 - 2 * fj)
    ni = int(i0.max()) + 2
    nj = int(j0.max()) + 2
    lat = u01(hash2(seed, np.arange(ni)[:, None], np.arange(nj)[None, :]))
    a = lat[i0][:, j0]
    b = lat[i0 + 1][:, j0]
    c = lat[i0][:, j0 + 1]
    d = lat[i0 + 1][:, j0 + 1]
    return (a * (1 - fi) + b * fi) * (1 - fj) + (c * (1 - fi) + d * fi) * fj


def _nan_mask(seed, rows, cols, frac):
    if frac <= 0:
        return None
    h = hash2(seed ^ 0x5EED, np.arange(rows)[:, None], np.arange(cols)[None, :])
    return u01(h) < frac


def flat_map(rows=500, cols=500, res=0.02, pos=(0.0, 0.0), start=(0, 0)) -> HostMap:
    """C1: trav = 1, elev = 0."""
    m = HostMap(rows, cols, res, pos, start)
    m.set_layer("traversability", np.ones((rows, cols), np.float32))
    m.set_layer("elevation", np.zeros((rows, cols), np.float32))
    return m


def stairs_stones_map(rows=2000, cols=2000, res=0.01, seed=20200802, nan_frac=0.005, pos=(0.0, 0.0),
                      start=(0, 0)) -> HostMap:
    """C2: x<0 stairs (0.30 m treads, 0.15 m risers, trav 0 within +-0.02 m of each riser);
    x>=0 stepping stones (0.25 m squares on a 0.40 m pitch: trav 1 / elev 0.05; gaps trav 0.2 / elev 0)."""
    x, y = cell_centres(rows, cols, res, pos)
    X = np.broadcast_to(x[:, None], (rows, cols))
    Y = np.broadcast_to(y[None, :], (rows, cols))
    half = rows * res / 2
    step = np.floor((X - pos[0] + half) / 0.30)
    elev_st = 0.15 * step
    dist_riser = np.abs((X - pos[0] + half) - 0.30 * np.round((X - pos[0] + half) / 0.30))
    trav_st = np.where(dist_riser <= 0.02, 0.0, 1.0)
    mx = np.mod(X - pos[0], 0.40)
    my = np.mod(Y - pos[1] + 100.0, 0.40)
    on = (mx < 0.25) & (my < 0.25)
    trav_ss = np.where(on, 1.0, 0.2)
    elev_ss = np.where(on, 0.05, 0.0)
    left = (X - pos[0]) < 0
    trav = np.where(left, trav_st, trav_ss).astype(np.float32)
    elev = np.where(left, elev_st, elev_ss).astype(np.float32)
    nm = _nan_mask(seed, rows, cols, nan_frac)
    if nm is not None:
        trav[nm] = np.nan
        elev[_nan_mask(seed + 1, rows, cols, nan_frac)] = np.nan
    m = HostMap(rows, cols, res, pos, start)
    m.set_layer("traversability", _roll_to_buffer(trav, start))
    m.set_layer("elevation", _roll_to_buffer(elev, start))
    return m


def rough_map(rows=4000, cols=4000, res=0.02, seed=3, nan_frac=0.01, pos=(0.0, 0.0), start=(1234, 2777),
              with_slope=False, pitch_m=0.5) -> HostMap:
    """C3/C4/C5: rough terrain from smooth value noise n1 in [0,1).  trav = clip(1.79 - 1.333*n1, 0, 1) puts about
    a quarter of the cells below the default threshold 0.9 and ~8 % below the candidate threshold 0.7, in patches
    of ~0.5 m; elev = 0.3*n2; `nan_frac` of the cells are NaN (independently per layer)."""
    pitch = max(2, int(round(pitch_m / res)))
    n1 = _value_noise(seed, rows, cols, pitch)
    n2 = _value_noise(seed + 101, rows, cols, pitch * 2)
    trav = np.clip(1.79 - 1.333 * n1, 0.0, 1.0).astype(np.float32)
    elev = (0.3 * n2).astype(np.float32)
    layers = {"traversability": trav, "elevation": elev}
    if with_slope:
        gx, gy = np.gradient(elev.astype(np.float64), res)
        layers["slope"] = np.hypot(gx, gy).astype(np.float32)
    nm = _nan_mask(seed, rows, cols, nan_frac)
    if nm is not None:
        trav[nm] = np.nan
        elev[_nan_mask(seed + 1, rows, cols, nan_frac)] = np.nan
    m = HostMap(rows, cols, res, pos, start)
    for k, v in layers.items():
        m.set_layer(k, _roll_to_buffer(v, start))
    return m


def uniform_queries(n, xlim, ylim, seed, z=0.0, param_id=0):
    """n start poses with x ~ U[xlim), y ~ U[ylim) from the counter hash."""
    i = np.arange(n)
    x = xlim[0] + (xlim[1] - xlim[0]) * u01(hash2(seed, i, np.zeros_like(i)))
    y = ylim[0] + (ylim[1] - ylim[0]) * u01(hash2(seed, i, np.ones_like(i)))
    return make_queries(np.stack([x, y, np.full(n, float(z))], axis=1), param_id)


# ---- the BASELINE.json configs ---------------------------------------------------------------------------------
def config_c1():
    """trot, 8 cycles, 1 query, flat 500x500 @ 0.02 m."""
    return flat_map(), make_queries([[-4.0, 0.0, 0.0]]), 8


def config_c2(n_queries=1024):
    """walk, 32 cycles, 1024 start poses on the stairs / stepping-stones map 2000x2000 @ 0.01 m."""
    m = stairs_stones_map()
    h = n_queries // 2
    a = uniform_queries(h, (-9.5, -9.0), (-8.0, 8.0), seed=1)
    b = uniform_queries(n_queries - h, (0.5, 1.0), (-8.0, 8.0), seed=2)
    return m, np.concatenate([a, b]), 32


def config_c3(n_queries=65536, rows=4000, cols=4000, seed_map=3, seed_q=4, start=(1234, 2777)):
    """trot, 65,536 queries x 16 cycles on a 4000x4000 @ 0.02 m rough map with a wrapped start index."""
    m = rough_map(rows, cols, 0.02, seed=seed_map, start=start)
    half = rows * 0.02 / 2
    # 16 cycles advance 2.9 m; keep every window >= 0.5 m inside the map
    q = uniform_queries(n_queries, (-half + 1.0, half - 4.5), (-half + 1.0, half - 1.0), seed=seed_q)
    return m, q, 16


def config_c5(n_queries=262144, rows=16000, cols=16000, seed_map=6, seed_q=7, start=(4321, 12345), with_map=True):
    """trot, 262,144 queries x 32 cycles on a 16000x16000 @ 0.01 m rough map with three layers (traversability,
    elevation, slope = |grad elevation|; 3 x 1.02 GB).  with_map=False returns the geometry only."""
    if with_map:
        m = rough_map(rows, cols, 0.01, seed=seed_map, start=start, with_slope=True)
    else:
        m = HostMap(rows, cols, 0.01, (0.0, 0.0), start)
    half = rows * 0.01 / 2
    # 32 cycles advance 5.8 m; keep every window >= 0.5 m inside the map
    q = uniform_queries(n_queries, (-half + 1.0, half - 7.5), (-half + 1.0, half - 1.0), seed=seed_q)
    return m, q, 32


def config_c4(n_poses=64, n_steps=64, rows=1500, cols=1500, seed_map=5, seed_q=6, n_headings=1, with_map=True):
    """BASELINE config 4: {trot, walk} x n_steps step lengths in [0.10, 0.26] (skew = stepLength/4, yaml:44) x
    n_headings headings in [0, 2 pi) x n_poses start poses, 8 cycles, on a C3-style rough map with start index (0,0).
    -> (map, params list [2*n_steps*n_headings], queries [len(params)*n_poses] with param_id, cycles); parameter set
    index = (gait*n_steps + step)*n_headings + heading.  n_headings = 1 is the heading-0 slice, which is the
    reference's own behaviour (the body always advances along +x, FP.cpp:2199-2202); the heading sweep is the
    extension defined by fhp_params.heading.  with_map=False returns the geometry only."""
    from ._abi import FHP_GAIT_TROT, FHP_GAIT_WALK, yaml_params
    if with_map:
        m = rough_map(rows, cols, 0.02, seed=seed_map, start=(0, 0))
    else:
        m = HostMap(rows, cols, 0.02, (0.0, 0.0), (0, 0))
    params = []
    for gait in (FHP_GAIT_TROT, FHP_GAIT_WALK):
        for sl in np.linspace(0.10, 0.26, n_steps):
            sl32 = float(np.float32(sl))
            for k in range(n_headings):
                params.append(yaml_params(gait=gait, step_length=sl32, skew=float(np.float32(sl32 / 4)),
                                          heading=2.0 * np.pi * k / n_headings))
    half = rows * 0.02 / 2
    # 8 cycles advance <= 2.1 m in the heading direction; keep every window >= 0.5 m inside the map
    lo, hi = (-half + 1.0, half - 3.5) if n_headings == 1 else (-half + 3.5, half - 3.5)
    ylo, yhi = (-half + 1.0, half - 1.0) if n_headings == 1 else (lo, hi)
    poses = uniform_queries(n_poses, (lo, hi), (ylo, yhi), seed=seed_q)
    q = np.tile(poses, len(params))
    q["param_id"] = np.repeat(np.arange(len(params), dtype=np.int32), n_poses)
    return m, params, q, 8
