"""ctypes binding of oracle/liboracle.so — ORACLE = TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module;
the product package (quadrupedal_foothold_planner_b200) never does."""
import ctypes as C
import os
import subprocess

import numpy as np

from quadrupedal_foothold_planner_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        src_newer = not os.path.exists(path) or any(
            os.path.getmtime(os.path.join(ROOT, "oracle", f)) > os.path.getmtime(path)
            for f in ("oracle_capi.cpp", "fp_oracle.hpp", "gm_core.hpp"))
        if src_newer:
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_spiral_cells.restype = C.c_int64
        _LIB.oracle_circle_cells.restype = C.c_int64
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def plan(hmap, params, queries, gait_cycles, n_threads=1, centroid=True):
    """-> (nominal rows [nq, R], centroid rows [nq, R] or None, status [nq])"""
    params = list(params) if isinstance(params, (list, tuple)) else [params]
    parr = (_abi.Params * len(params))(*params)
    d, ptrs, keep = hmap.desc()
    nq = len(queries)
    R = _abi.rows_per_query(gait_cycles)
    nom = np.zeros((nq, R), dtype=_abi.FOOTHOLD_DTYPE)
    cen = np.zeros((nq, R), dtype=_abi.FOOTHOLD_DTYPE) if centroid else None
    st = np.zeros(nq, dtype=_abi.STATUS_DTYPE)
    q = np.ascontiguousarray(queries)
    rc = lib().oracle_plan(C.byref(d), ptrs, parr, C.c_int32(len(params)), _p(q), C.c_int64(nq),
                           C.c_int32(gait_cycles), _p(nom), _p(cen) if centroid else None, _p(st),
                           C.c_int32(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle_plan failed rc={rc}")
    return nom, cen, st


def plan_ex(hmap, params, queries, gait_cycles):
    """-> (nominal, centroid, default rows [nq, R], kpi [nq, cycles, 8], status)"""
    params = list(params) if isinstance(params, (list, tuple)) else [params]
    parr = (_abi.Params * len(params))(*params)
    d, ptrs, keep = hmap.desc()
    nq = len(queries)
    R = _abi.rows_per_query(gait_cycles)
    nom = np.zeros((nq, R), dtype=_abi.FOOTHOLD_DTYPE)
    cen = np.zeros((nq, R), dtype=_abi.FOOTHOLD_DTYPE)
    dfl = np.zeros((nq, R), dtype=_abi.FOOTHOLD_DTYPE)
    kpi = np.zeros((nq, max(1, gait_cycles), 8), dtype=np.float64)
    st = np.zeros(nq, dtype=_abi.STATUS_DTYPE)
    q = np.ascontiguousarray(queries)
    rc = lib().oracle_plan_ex(C.byref(d), ptrs, parr, C.c_int32(len(params)), _p(q), C.c_int64(nq),
                              C.c_int32(gait_cycles), _p(nom), _p(cen), _p(dfl), _p(kpi), _p(st))
    if rc != 0:
        raise RuntimeError(f"oracle_plan_ex failed rc={rc}")
    return nom, cen, dfl, kpi, st


def plan_latency(hmap, params, query, gait_cycles, repeats=200, leg_threads=False):
    """-> (ms per call [repeats], status) of ONE plan on the CPU; leg_threads: the reference's 4 std::threads per cycle."""
    d, ptrs, keep = hmap.desc()
    q = np.ascontiguousarray(query).reshape(1)
    ms = np.zeros(repeats, dtype=np.float64)
    st = np.zeros(1, dtype=_abi.STATUS_DTYPE)
    rc = lib().oracle_plan_latency(C.byref(d), ptrs, C.byref(params), _p(q), C.c_int32(gait_cycles), C.c_int32(repeats),
                                   C.c_int32(1 if leg_threads else 0), _p(ms), _p(st))
    assert rc == 0, rc
    return ms, st[0]


def touch_bytes(hmap, params, queries, gait_cycles, with_centroid_height=True):
    params = list(params) if isinstance(params, (list, tuple)) else [params]
    parr = (_abi.Params * len(params))(*params)
    d, ptrs, keep = hmap.desc()
    nq = len(queries)
    out = np.zeros(nq, dtype=np.int64)
    sp = np.zeros(nq, dtype=np.int64)
    q = np.ascontiguousarray(queries)
    rc = lib().oracle_plan_touch_bytes(C.byref(d), ptrs, parr, C.c_int32(len(params)), _p(q), C.c_int64(nq),
                                       C.c_int32(gait_cycles), _p(out), _p(sp),
                                       C.c_int32(1 if with_centroid_height else 0))
    assert rc == 0, rc
    return out, sp


def spiral_rank_hist(hmap, params, queries, gait_cycles):
    """hist[0..255]: spiral searches that ended on their k-th visited cell; hist[256]: searches that found none."""
    params = list(params) if isinstance(params, (list, tuple)) else [params]
    parr = (_abi.Params * len(params))(*params)
    d, ptrs, keep = hmap.desc()
    hist = np.zeros(257, dtype=np.int64)
    q = np.ascontiguousarray(queries)
    rc = lib().oracle_spiral_rank_hist(C.byref(d), ptrs, parr, C.c_int32(len(params)), _p(q), C.c_int64(len(q)),
                                       C.c_int32(gait_cycles), _p(hist))
    assert rc == 0, rc
    return hist


def check_foothold(hmap, params, centres, rect_centres):
    d, ptrs, keep = hmap.desc()
    c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
    r = np.ascontiguousarray(rect_centres, dtype=np.float64).reshape(-1, 2)
    n = len(c)
    out = np.zeros(n, dtype=_abi.FOOTHOLD_DTYPE)
    valid = np.zeros(n, dtype=np.uint8)
    rc = lib().oracle_check_foothold(C.byref(d), ptrs, C.byref(params), _p(c), _p(r), C.c_int64(n), _p(out), _p(valid))
    assert rc == 0, rc
    return out, valid


def check_candidate(hmap, params, centres, rect_centres):
    d, ptrs, keep = hmap.desc()
    c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
    r = np.ascontiguousarray(rect_centres, dtype=np.float64).reshape(-1, 2)
    out = np.zeros(len(c), dtype=_abi.FOOTHOLD_DTYPE)
    valid = np.zeros(len(c), dtype=np.uint8)
    rc = lib().oracle_check_candidate(C.byref(d), ptrs, C.byref(params), _p(c), _p(r), C.c_int64(len(c)), _p(out),
                                      _p(valid))
    assert rc == 0, rc
    return out, valid


def check_circle_polygon(hmap, params, foot_centres, rect_centres):
    d, ptrs, keep = hmap.desc()
    c = np.ascontiguousarray(foot_centres, dtype=np.float64).reshape(-1, 2)
    r = np.ascontiguousarray(rect_centres, dtype=np.float64).reshape(-1, 2)
    valid = np.zeros(len(c), dtype=np.uint8)
    rc = lib().oracle_check_circle_polygon(C.byref(d), ptrs, C.byref(params), _p(c), _p(r), C.c_int64(len(c)), _p(valid))
    assert rc == 0, rc
    return valid


def search_geometry(params, query_xyz, adjusted_y, cur_xyz):
    q = _abi.make_queries([query_xyz])
    cur = np.ascontiguousarray(cur_xyz, dtype=np.float64).reshape(4, 3)
    c, n, p = np.zeros(2), np.zeros((4, 2)), np.zeros((4, 4, 2))
    rc = lib().oracle_search_geometry(C.byref(params), _p(q), C.c_double(adjusted_y), _p(cur), _p(c), _p(n), _p(p))
    assert rc == 0, rc
    return c, n, p


def check_centroid(hmap, params, points):
    d, ptrs, keep = hmap.desc()
    c = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    n = len(c)
    out = np.zeros(n, dtype=_abi.FOOTHOLD_DTYPE)
    case = np.zeros(n, dtype=np.int8)
    rc = lib().oracle_check_centroid(C.byref(d), ptrs, C.byref(params), _p(c), C.c_int64(n), _p(out), _p(case))
    assert rc == 0, rc
    return out, case


def mean_height(hmap, params, centres):
    d, ptrs, keep = hmap.desc()
    c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
    z = np.zeros(len(c), dtype=np.float32)
    rc = lib().oracle_mean_height(C.byref(d), ptrs, C.byref(params), _p(c), C.c_int64(len(c)), _p(z))
    assert rc == 0, rc
    return z


def check_default(hmap, params, centres):
    d, ptrs, keep = hmap.desc()
    c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
    v = np.zeros(len(c), dtype=np.uint8)
    rc = lib().oracle_check_default(C.byref(d), ptrs, C.byref(params), _p(c), C.c_int64(len(c)), _p(v))
    assert rc == 0, rc
    return v


def spiral_cells(hmap, cx, cy, radius, wrap_correct=False, max_out=4096):
    d, ptrs, keep = hmap.desc(names=[])
    out = np.zeros((max_out, 2), dtype=np.int32)
    n = lib().oracle_spiral_cells(C.byref(d), C.c_double(cx), C.c_double(cy), C.c_double(radius),
                                  C.c_int32(1 if wrap_correct else 0), _p(out), C.c_int64(max_out))
    return out[:n]


def circle_cells(hmap, cx, cy, radius, max_out=4096):
    d, ptrs, keep = hmap.desc(names=[])
    out = np.zeros((max_out, 2), dtype=np.int32)
    n = lib().oracle_circle_cells(C.byref(d), C.c_double(cx), C.c_double(cy), C.c_double(radius), _p(out),
                                  C.c_int64(max_out))
    return out[:n]


def submap_info(hmap, px, py, lx, ly):
    d, ptrs, keep = hmap.desc(names=[])
    out = np.zeros(4, dtype=np.int32)
    pl = np.zeros(4, dtype=np.float64)
    ok = lib().oracle_submap_info(C.byref(d), C.c_double(px), C.c_double(py), C.c_double(lx), C.c_double(ly), _p(out),
                                  _p(pl))
    return bool(ok), out, pl


def get_index(hmap, px, py):
    d, ptrs, keep = hmap.desc(names=[])
    out = np.zeros(2, dtype=np.int32)
    ok = lib().oracle_get_index(C.byref(d), C.c_double(px), C.c_double(py), _p(out))
    return bool(ok), (int(out[0]), int(out[1]))


def get_position(hmap, i0, i1):
    d, ptrs, keep = hmap.desc(names=[])
    out = np.zeros(2, dtype=np.float64)
    ok = lib().oracle_get_position(C.byref(d), C.c_int32(i0), C.c_int32(i1), _p(out))
    return bool(ok), (float(out[0]), float(out[1]))


def polygon_inside(verts, px, py):
    v = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 2)
    return bool(lib().oracle_polygon_inside(_p(v), C.c_int32(len(v)), C.c_double(px), C.c_double(py)))


def initial_geometry(params, query_xyz=(0.0, 0.0, 0.0)):
    q = _abi.make_queries([query_xyz])
    init = np.zeros((4, 2))
    bias = np.zeros((4, 2))
    first = np.zeros((4, 2))
    lib().oracle_initial_geometry(C.byref(params), _p(q), _p(init), _p(bias), _p(first))
    return init, bias, first


# ---- the reference's own FootholdPlanner.cpp compiled against shims (oracle/ref_shim -> oracle/_ref) -------------
_REF = None


def ref_available() -> bool:
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfhp_ref.so")) or os.path.isdir("/root/reference")


def ref_lib():
    """oracle/_ref/libfhp_ref.so; (re)built from /root/reference when that tree is present (this container only —
    the GPU box uses the prebuilt file that travels with the repo)."""
    global _REF
    if _REF is None:
        path = os.path.join(ROOT, "oracle", "_ref", "libfhp_ref.so")
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref_shim")])
        _REF = C.CDLL(path)
    return _REF


def ref_plan(hmap, params, queries, gait_cycles, wrap_correct=False, centroid=True, extras=False):
    """plan_global_footholds through the reference class, one service call per query.
    -> (nominal rows, centroid rows | None, status); status.reserved == 1 where the service returned false."""
    params = list(params) if isinstance(params, (list, tuple)) else [params]
    parr = (_abi.Params * len(params))(*params)
    d, ptrs, keep = hmap.desc()
    nq = len(queries)
    R = _abi.rows_per_query(gait_cycles)
    nom = np.zeros((nq, R), dtype=_abi.FOOTHOLD_DTYPE)
    cen = np.zeros((nq, R), dtype=_abi.FOOTHOLD_DTYPE) if centroid else None
    st = np.zeros(nq, dtype=_abi.STATUS_DTYPE)
    q = np.ascontiguousarray(queries)
    dfl = np.zeros((nq, R), dtype=_abi.FOOTHOLD_DTYPE) if extras else None
    kpi = np.zeros((nq, max(1, gait_cycles), 8), dtype=np.float64) if extras else None
    rc = ref_lib().ref_plan(C.byref(d), ptrs, parr, _p(q), C.c_int64(nq), C.c_int32(gait_cycles),
                            C.c_int32(1 if wrap_correct else 0), _p(nom), _p(cen) if centroid else None, _p(st),
                            _p(dfl) if extras else None, _p(kpi) if extras else None)
    if rc != 0:
        raise RuntimeError(f"ref_plan failed rc={rc}")
    if extras:
        return nom, cen, st, dfl, kpi
    return nom, cen, st
