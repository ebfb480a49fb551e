"""Host side of libfhp: a ctypes binding of include/fhp.h (`Fhp`) and a mirror of the reference's
`foothold_planner::FootholdPlanner` class interface (`FootholdPlanner`) on top of it.

The reference's host language is C++ (ROS1); a ROS node binds the same C-ABI from C++ (see INTEGRATION.md and
csrc/FootholdPlanner.hpp).  This Python mirror keeps the reference's member names, argument meaning and error
behaviour so the parity tests read like tests of the reference class.

There is NO CPU fallback: if libfhp.so is missing it is built with nvcc; if no sm_100 GPU is usable every compute
call raises FhpError(FHP_ERR_CUDA).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _abi
from ._abi import (FOOTHOLD_DTYPE, QUERY_DTYPE, STATUS_DTYPE, HostMap, MapDesc, Params, make_queries, rows_per_query,
                   yaml_params)

_LIB = None


class FhpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libfhp error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfhp.so")


def load_lib():
    """dlopen libfhp.so, building it first if the sources are newer (nvcc required for that)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    from . import build as _build
    path = lib_path()
    if os.environ.get("FHP_LIB"):  # experiments: an alternative build of the same library (scratch A/B runs)
        path = os.environ["FHP_LIB"]
    elif _build.needs_build():
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.fhp_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.fhp_destroy.argtypes = [vp]
    L.fhp_destroy.restype = None
    L.fhp_last_error.argtypes = [vp]
    L.fhp_last_error.restype = C.c_char_p
    L.fhp_default_params.argtypes = [C.POINTER(Params)]
    L.fhp_default_params.restype = None
    L.fhp_version.restype = C.c_char_p
    L.fhp_set_map.argtypes = [vp, C.POINTER(MapDesc), vp]
    L.fhp_set_map_device.argtypes = [vp, C.POINTER(MapDesc), vp]
    L.fhp_set_map_msg.argtypes = [vp, vp]
    L.fhp_set_map_wire.argtypes = [vp, vp, C.c_uint64]
    L.fhp_map_info.argtypes = [vp, C.POINTER(MapDesc)]
    L.fhp_map_layer_device.argtypes = [vp, C.c_char_p]
    L.fhp_map_layer_device.restype = vp
    L.fhp_map_changed.argtypes = [vp]
    L.fhp_plan.argtypes = [vp, vp, i32, vp, i64, i32, vp, vp, vp]
    L.fhp_plan_ex.argtypes = [vp, vp, i32, vp, i64, i32, vp]
    L.fhp_cycle_kpi.argtypes = [vp, vp, vp, i32, vp, vp]
    L.fhp_plan_compact.argtypes = [vp, vp, i32, vp, i64, i32, vp, i64, vp, vp]
    L.fhp_plan_compact_packed.argtypes = [vp, vp, i32, vp, i64, i32, vp, i64, vp, vp]
    L.fhp_plan_compact24.argtypes = [vp, vp, i32, vp, i64, i32, vp, i64, vp, vp]
    L.fhp_plan_stream.argtypes = [vp, vp, i32, vp, i64, i32, vp, vp, vp]
    L.fhp_stream_expand.argtypes = [vp, vp, vp, i64, i32, vp, i64, vp, i32]
    L.fhp_plan_device.argtypes = [vp, vp, i32, vp, i64, i32, vp, vp, vp, vp]
    L.fhp_synchronize.argtypes = [vp]
    L.fhp_check_foothold.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.fhp_check_centroid.argtypes = [vp, vp, vp, i64, vp, vp]
    L.fhp_mean_height.argtypes = [vp, vp, vp, i64, vp]
    L.fhp_check_default.argtypes = [vp, vp, vp, i64, vp]
    L.fhp_check_candidate.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.fhp_check_circle_polygon.argtypes = [vp, vp, vp, vp, i64, vp]
    L.fhp_initial_stance.argtypes = [vp, vp, vp, vp]
    L.fhp_search_geometry.argtypes = [vp, vp, C.c_double, vp, vp, vp, vp]
    L.fhp_kernel_launches.argtypes = [vp]
    L.fhp_kernel_launches.restype = i64
    L.fhp_broadcast_map.argtypes = [vp, C.POINTER(MapDesc), vp, C.c_int, C.c_int]
    _LIB = L
    return L


def _ptr(a) -> C.c_void_p:
    """address of a numpy array / torch tensor / raw int"""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


class PlanOutputs(C.Structure):  # fhp_plan_outputs
    _fields_ = [("nominal", C.c_void_p), ("centroid", C.c_void_p), ("default_chain", C.c_void_p),
                ("status", C.c_void_p)]


class Fhp:
    """One libfhp context = one device + one resident map.  Thin, explicit wrapper of the C-ABI."""

    def __init__(self, device: int = 0):
        self.L = load_lib()
        h = C.c_void_p()
        rc = self.L.fhp_create(int(device), C.byref(h))
        if rc != 0:
            raise FhpError(rc, (self.L.fhp_last_error(None) or b"").decode())
        self.h = h
        self.device = int(device)
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.L.fhp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise FhpError(rc, (self.L.fhp_last_error(self.h) or b"").decode())

    # -- map ingest (FootholdPlanner::gridmapCallback, FP.cpp:504-536)
    def set_map(self, hmap: HostMap):
        d, ptrs, keep = hmap.desc()
        self._check(self.L.fhp_set_map(self.h, C.byref(d), C.cast(ptrs, C.c_void_p)))
        self.map_geometry = hmap

    def set_map_wire(self, wire: bytes):
        """fhp_set_map_wire: ROS1-serialised grid_map_msgs/GridMap bytes (gridmap_msg.serialize_gridmap)."""
        buf = np.frombuffer(wire, dtype=np.uint8)
        self._check(self.L.fhp_set_map_wire(self.h, C.c_void_p(buf.ctypes.data), len(buf)))

    def set_map_msg(self, msg):
        """fhp_set_map_msg: a deserialised message (gridmap_msg.GridMapMessage)."""
        cmsg, keep = msg.as_c()
        self._check(self.L.fhp_set_map_msg(self.h, C.cast(C.pointer(cmsg), C.c_void_p)))

    def map_info(self) -> MapDesc:
        d = MapDesc()
        self._check(self.L.fhp_map_info(self.h, C.byref(d)))
        return d

    def set_map_device(self, geometry: HostMap, names: Sequence[str], device_ptrs: Sequence[int], keepalive=None):
        """Adopt device-resident layers (no copy).  `geometry` supplies rows/cols/resolution/pos/start only."""
        d, _, keep = geometry.desc(names=[])
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        d.n_layers = len(names)
        d.layer_names = C.cast(arr, C.POINTER(C.c_char_p))
        ptrs = (C.c_void_p * len(names))(*[int(p) for p in device_ptrs])
        self._check(self.L.fhp_set_map_device(self.h, C.byref(d), C.cast(ptrs, C.c_void_p)))
        self._keep = (keepalive, arr, ptrs)
        self.map_geometry = geometry

    def map_changed(self):
        """fhp_map_changed: an adopted / broadcast layer was rewritten in place."""
        self._check(self.L.fhp_map_changed(self.h))

    def layer_device_ptr(self, name: str) -> int:
        return int(self.L.fhp_map_layer_device(self.h, name.encode()) or 0)

    # -- plan (FootholdPlanner::globalFootholdPlan, FP.cpp:539-1602)
    @staticmethod
    def _params_array(params):
        params = list(params) if isinstance(params, (list, tuple)) else [params]
        return (Params * len(params))(*params), len(params)

    def plan(self, params, queries: np.ndarray, gait_cycles: int, centroid: bool = False, out=None):
        """Host buffers in, host buffers out.  -> (nominal [nq,R], centroid [nq,R] | None, status [nq])"""
        parr, npar = self._params_array(params)
        q = np.ascontiguousarray(queries, dtype=QUERY_DTYPE)
        nq = len(q)
        R = rows_per_query(max(0, min(255, int(gait_cycles))))
        if out is None:
            nom = np.zeros((nq, R), dtype=FOOTHOLD_DTYPE)
            cen = np.zeros((nq, R), dtype=FOOTHOLD_DTYPE) if centroid else None
            st = np.zeros(nq, dtype=STATUS_DTYPE)
        else:
            nom, cen, st = out
        self._check(self.L.fhp_plan(self.h, C.cast(parr, C.c_void_p), npar, _ptr(q), nq, int(gait_cycles), _ptr(nom),
                                    _ptr(cen) if centroid else None, _ptr(st)))
        return nom, cen, st

    def plan_ex(self, params, queries: np.ndarray, gait_cycles: int):
        """fhp_plan_ex: -> (nominal, centroid, default-chain rows [nq,R], status [nq])"""
        parr, npar = self._params_array(params)
        q = np.ascontiguousarray(queries, dtype=QUERY_DTYPE)
        nq = len(q)
        R = rows_per_query(max(0, min(255, int(gait_cycles))))
        nom = np.zeros((nq, R), dtype=FOOTHOLD_DTYPE)
        cen = np.zeros((nq, R), dtype=FOOTHOLD_DTYPE)
        dfl = np.zeros((nq, R), dtype=FOOTHOLD_DTYPE)
        st = np.zeros(nq, dtype=STATUS_DTYPE)
        out = PlanOutputs(nom.ctypes.data, cen.ctypes.data, dfl.ctypes.data, st.ctypes.data)
        self._check(self.L.fhp_plan_ex(self.h, C.cast(parr, C.c_void_p), npar, _ptr(q), nq, int(gait_cycles),
                                       C.cast(C.pointer(out), C.c_void_p)))
        return nom, cen, dfl, st

    def _compact_fn(self, packed=False, fmt=None):
        fmt = fmt or (32 if packed else 40)
        return {40: (self.L.fhp_plan_compact, FOOTHOLD_DTYPE), 32: (self.L.fhp_plan_compact_packed, _abi.FOOTHOLD_PACKED_DTYPE),
                24: (self.L.fhp_plan_compact24, _abi.FOOTHOLD24_DTYPE)}[fmt]

    def plan_compact(self, params, queries: np.ndarray, gait_cycles: int, capacity_rows: Optional[int] = None,
                     packed: bool = False, fmt: Optional[int] = None):
        """fhp_plan_compact / _packed / 24: -> (rows [total], row_offsets [nq+1], status [nq]); query q owns
        rows[row_offsets[q]:row_offsets[q+1]] (the service's variable-length footholds[]).  packed: 32-byte rows;
        fmt = 24: the 24-byte message rows."""
        parr, npar = self._params_array(params)
        q = np.ascontiguousarray(queries, dtype=QUERY_DTYPE)
        nq = len(q)
        cap = nq * rows_per_query(max(0, min(255, int(gait_cycles)))) if capacity_rows is None else int(capacity_rows)
        fn, dt = self._compact_fn(packed, fmt)
        rows = np.zeros(max(cap, 1), dtype=dt)
        offs = np.zeros(nq + 1, dtype=np.int64)
        st = np.zeros(nq, dtype=STATUS_DTYPE)
        self._check(fn(self.h, C.cast(parr, C.c_void_p), npar, _ptr(q), nq, int(gait_cycles),
                                            _ptr(rows), cap, _ptr(offs), _ptr(st)))
        return rows[:offs[nq]], offs, st

    def plan_compact_raw(self, parr, npar, q_ptr, nq, gait_cycles, rows_ptr, capacity_rows, offs_ptr, st_ptr,
                         packed: bool = False, fmt: Optional[int] = None):
        """fhp_plan_compact[_packed|24] on raw host addresses (pinned torch buffers in bench.py)."""
        fn, _ = self._compact_fn(packed, fmt)
        self._check(fn(self.h, C.cast(parr, C.c_void_p), npar, _ptr(q_ptr), nq, int(gait_cycles),
                                            _ptr(rows_ptr), int(capacity_rows), _ptr(offs_ptr), _ptr(st_ptr)))

    def plan_stream_raw(self, parr, npar, q_ptr, nq, gait_cycles, blocks_ptr, mask_ptr, st_ptr):
        """fhp_plan_stream on raw PINNED host addresses (see include/fhp.h for the block format)."""
        self._check(self.L.fhp_plan_stream(self.h, C.cast(parr, C.c_void_p), npar, _ptr(q_ptr), nq, int(gait_cycles),
                                           _ptr(blocks_ptr), _ptr(mask_ptr), _ptr(st_ptr)))

    def plan_stream(self, params, queries: np.ndarray, gait_cycles: int):
        """fhp_plan_stream with pinned buffers allocated here (torch).  -> (blocks [cycles+1, groups, 32] of POINT20_DTYPE,
        masks [nq, words, 4] of STREAM_MASK_DTYPE, status [nq]).  stream_rows_of() extracts one query's rows."""
        import torch
        parr, npar = self._params_array(params)
        q = np.ascontiguousarray(queries, dtype=QUERY_DTYPE)
        nq = len(q)
        cycles = int(gait_cycles)
        words = max(1, (cycles + 31) // 32)
        groups = max(1, (nq + 7) // 8)
        raw = torch.zeros((cycles + 1) * groups * 640 + 128, dtype=torch.uint8).pin_memory()
        off = (-raw.data_ptr()) % 128
        blocks = raw[off: off + (cycles + 1) * groups * 640]
        mask = torch.zeros(max(nq, 1) * words * 4 * 2, dtype=torch.int32).pin_memory()
        st = torch.zeros(max(nq, 1) * STATUS_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        self.plan_stream_raw(parr, npar, q, nq, cycles, blocks, mask, st)
        self._stream_keep = raw
        return (blocks.numpy().view(_abi.POINT20_DTYPE).reshape(cycles + 1, groups, 32),
                mask.numpy().view(_abi.STREAM_MASK_DTYPE).reshape(max(nq, 1), words, 4)[:nq], st.numpy().view(STATUS_DTYPE)[:nq])

    def stream_expand(self, blocks, masks, status, gait_cycles: int, n_threads: int = 0):
        """fhp_stream_expand: streamed blocks -> (rows [total] of FOOTHOLD24_DTYPE, row_offsets [nq+1])."""
        nq = len(status)
        rows = np.zeros(max(1, int(status["n_rows"].sum())), dtype=_abi.FOOTHOLD24_DTYPE)
        offs = np.zeros(nq + 1, dtype=np.int64)
        b = np.ascontiguousarray(blocks)
        m = np.ascontiguousarray(masks)
        self._check(self.L.fhp_stream_expand(_ptr(b), _ptr(m), _ptr(np.ascontiguousarray(status)), nq, int(gait_cycles), _ptr(rows),
                                             len(rows), _ptr(offs), int(n_threads)))
        return rows[: offs[nq]], offs

    @staticmethod
    def stream_rows_of(blocks, masks, q: int):
        """One query's response rows from the streamed blocks: -> (points [n_rows] POINT20, foot_id, gait_cycle_id, kind)."""
        cycles = blocks.shape[0] - 1
        G, lo = q >> 3, (q >> 3) << 3
        pts, foot, cyc, kind = [blocks[0, G, (q & 7) * 4:(q & 7) * 4 + 4]], [0, 1, 2, 3], [0] * 4, [0] * 4
        for g in range(cycles):
            w, bit = g >> 5, np.uint32(1 << (g & 31))
            if not (masks[q, w, 0]["success"] & bit):
                continue
            rank = int(sum(1 for p in range(lo, q) if masks[p, w, 0]["success"] & bit))
            pts.append(blocks[g + 1, G, rank * 4: rank * 4 + 4])
            foot += [0, 1, 2, 3]
            cyc += [g] * 4
            kind += [2 if (masks[q, w, f]["candidate"] & bit) else 1 for f in range(4)]
        return np.concatenate(pts), np.array(foot), np.array(cyc), np.array(kind)

    def cycle_kpi(self, params: Params, query, rows: np.ndarray, n_rows: int):
        """fhp_cycle_kpi: (feet_distance [cycles,2], cog_speed [cycles,2]) of one query's rows (FP.cpp:2571-2623)."""
        q = np.ascontiguousarray(query, dtype=QUERY_DTYPE).reshape(1)
        r = np.ascontiguousarray(rows[:n_rows])
        cycles = n_rows // 4 - 1
        fd = np.zeros((max(cycles, 1), 2))
        cs = np.zeros((max(cycles, 1), 2))
        self._check(self.L.fhp_cycle_kpi(C.cast(C.pointer(params), C.c_void_p), _ptr(q), _ptr(r), int(n_rows), _ptr(fd),
                                         _ptr(cs)))
        return fd[:cycles], cs[:cycles]

    def plan_raw(self, parr, npar, q_ptr, nq, gait_cycles, nom_ptr, cen_ptr, st_ptr):
        """fhp_plan on raw host addresses (pinned torch buffers in bench.py)."""
        self._check(self.L.fhp_plan(self.h, C.cast(parr, C.c_void_p), npar, _ptr(q_ptr), nq, int(gait_cycles),
                                    _ptr(nom_ptr), _ptr(cen_ptr), _ptr(st_ptr)))

    def plan_device(self, parr, npar, d_queries, nq, gait_cycles, d_nom, d_cen, d_st, stream=None):
        """Device-resident buffers (torch CUDA tensors or raw device addresses); asynchronous."""
        self._check(self.L.fhp_plan_device(self.h, C.cast(parr, C.c_void_p), npar, _ptr(d_queries), nq,
                                           int(gait_cycles), _ptr(d_nom), _ptr(d_cen), _ptr(d_st), _ptr(stream)))

    def synchronize(self):
        self._check(self.L.fhp_synchronize(self.h))

    def kernel_launches(self) -> int:
        return int(self.L.fhp_kernel_launches(self.h))

    # -- fine-grained members
    def check_foothold(self, params: Params, centres, rect_centres):
        c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
        r = np.ascontiguousarray(rect_centres, dtype=np.float64).reshape(-1, 2)
        out = np.zeros(len(c), dtype=FOOTHOLD_DTYPE)
        valid = np.zeros(len(c), dtype=np.uint8)
        self._check(self.L.fhp_check_foothold(self.h, C.cast(C.pointer(params), C.c_void_p), _ptr(c), _ptr(r), len(c),
                                              _ptr(out), _ptr(valid)))
        return out, valid

    def check_centroid(self, params: Params, points):
        c = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        out = np.zeros(len(c), dtype=FOOTHOLD_DTYPE)
        case = np.zeros(len(c), dtype=np.int8)
        self._check(self.L.fhp_check_centroid(self.h, C.cast(C.pointer(params), C.c_void_p), _ptr(c), len(c), _ptr(out),
                                              _ptr(case)))
        return out, case

    def mean_height(self, params: Params, centres):
        c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
        z = np.zeros(len(c), dtype=np.float32)
        self._check(self.L.fhp_mean_height(self.h, C.cast(C.pointer(params), C.c_void_p), _ptr(c), len(c), _ptr(z)))
        return z

    def check_candidate(self, params: Params, centres, rect_centres):
        """checkCandidateFoothold (FP.cpp:2085-2114): the spiral search alone -> (rows, valid)."""
        c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
        r = np.ascontiguousarray(rect_centres, dtype=np.float64).reshape(-1, 2)
        out = np.zeros(len(c), dtype=FOOTHOLD_DTYPE)
        valid = np.zeros(len(c), dtype=np.uint8)
        self._check(self.L.fhp_check_candidate(self.h, C.cast(C.pointer(params), C.c_void_p), _ptr(c), _ptr(r), len(c),
                                               _ptr(out), _ptr(valid)))
        return out, valid

    def check_circle_polygon(self, params: Params, foot_centres, rect_centres):
        """checkCirclePolygonFoothold (FP.cpp:2117-2163) at arbitrary foot centres -> valid."""
        c = np.ascontiguousarray(foot_centres, dtype=np.float64).reshape(-1, 2)
        r = np.ascontiguousarray(rect_centres, dtype=np.float64).reshape(-1, 2)
        valid = np.zeros(len(c), dtype=np.uint8)
        self._check(self.L.fhp_check_circle_polygon(self.h, C.cast(C.pointer(params), C.c_void_p), _ptr(c), _ptr(r),
                                                    len(c), _ptr(valid)))
        return valid

    def check_default(self, params: Params, centres):
        c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
        v = np.zeros(len(c), dtype=np.uint8)
        self._check(self.L.fhp_check_default(self.h, C.cast(C.pointer(params), C.c_void_p), _ptr(c), len(c), _ptr(v)))
        return v


def initial_stance(params: Params, query_xyz=(0.0, 0.0, 0.0)):
    """fhp_initial_stance -> (initial [4,3], first_gait [4,3]); no device needed."""
    L = load_lib()
    q = make_queries([query_xyz])
    a, b = np.zeros((4, 3)), np.zeros((4, 3))
    rc = L.fhp_initial_stance(C.cast(C.pointer(params), C.c_void_p), _ptr(q), _ptr(a), _ptr(b))
    if rc != 0:
        raise FhpError(rc, "fhp_initial_stance")
    return a, b


def search_geometry(params: Params, query_xyz, adjusted_y: float, cur_xyz):
    """fhp_search_geometry -> (center [3], next [4,2], polygons [4,4,2]); no device needed."""
    L = load_lib()
    q = make_queries([query_xyz])
    cur = np.ascontiguousarray(cur_xyz, dtype=np.float64).reshape(4, 3)
    c, n, p = np.zeros(3), np.zeros((4, 2)), np.zeros((4, 4, 2))
    rc = L.fhp_search_geometry(C.cast(C.pointer(params), C.c_void_p), _ptr(q), float(adjusted_y), _ptr(cur), _ptr(c),
                               _ptr(n), _ptr(p))
    if rc != 0:
        raise FhpError(rc, "fhp_search_geometry")
    return c, n, p


# ---------------------------------------------------------------------------------------------------------------------
# Mirror of the reference's service / message types and class interface
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class Point:  # geometry_msgs/Point
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0


@dataclass
class Foothold:  # foothold_planner_msgs/Foothold (msg/Foothold.msg:1-3)
    point: Point = field(default_factory=Point)
    foot_id: int = 0
    gait_cycle_id: int = 0


@dataclass
class GlobalFootholds:  # foothold_planner_msgs/GlobalFootholds (msg/GlobalFootholds.msg:1-5)
    success: bool = False
    gait_cycles: int = 0
    gait_cycles_succeed: int = 0
    footholds: List[Foothold] = field(default_factory=list)


@dataclass
class GlobalFootholdPlanRequest:  # srv/GlobalFootholdPlan.srv:1-5 (the four foot fields are ignored, FP.cpp:546-549)
    gait_cycles: int = 0
    FR_current_foothold: Point = field(default_factory=Point)
    RR_current_foothold: Point = field(default_factory=Point)
    RL_current_foothold: Point = field(default_factory=Point)
    FL_current_foothold: Point = field(default_factory=Point)


@dataclass
class GlobalFootholdPlanResponse:  # srv/GlobalFootholdPlan.srv:7
    footholds: GlobalFootholds = field(default_factory=GlobalFootholds)


def rows_to_msg(rows: np.ndarray, status) -> GlobalFootholds:
    """Response assembly (FP.cpp:1378-1396, :1588): the first n_rows rows of one query -> GlobalFootholds."""
    msg = GlobalFootholds(bool(status["success"]), int(status["gait_cycles"]), int(status["gait_cycles_succeed"]))
    for r in rows[: int(status["n_rows"])]:
        msg.footholds.append(Foothold(Point(float(r["x"]), float(r["y"]), float(r["z"])), int(r["foot_id"]),
                                      int(r["gait_cycle_id"])))
    return msg


class FootholdPlanner:
    """Drop-in shaped like foothold_planner::FootholdPlanner (FootholdPlanner.hpp:38-757) for the hot path.

    Constructor parameters are the rosparams the path reads (FP.cpp:249-311); `initial_position` is
    initial_position/{x,y,z}.  gridmapCallback replaces the topic subscription (FP.cpp:188, :504);
    globalFootholdPlan is the `plan_global_footholds` service body (FP.cpp:539)."""

    def __init__(self, params: Optional[Params] = None, initial_position=(0.0, 0.0, 0.0), device: int = 0):
        # like the reference class: bug-compatible SpiralIterator (it ignores the map's start index)
        self.params = params if params is not None else yaml_params(spiral_mode=_abi.FHP_SPIRAL_FAITHFUL)
        self.initialPose_ = tuple(float(v) for v in initial_position)
        self.fhp = Fhp(device)
        self.nominalGlobalFootholdsMsg_ = GlobalFootholds()
        self.centroidGlobalFootholdsMsg_ = GlobalFootholds()

    # FP.cpp:504-536 — accepts what a subscriber can hold: the message (gridmap_msg.GridMapMessage), its ROS1 wire
    # bytes, or an already converted map (HostMap = grid_map::GridMap geometry + layers)
    def gridmapCallback(self, msg):
        if isinstance(msg, (bytes, bytearray, memoryview)):
            self.fhp.set_map_wire(bytes(msg))
        elif isinstance(msg, HostMap):
            self.fhp.set_map(msg)
        else:
            self.fhp.set_map_msg(msg)
        d = self.fhp.map_info()  # gridMapInfo_ (FP.cpp:508-515)
        self.gridMapInfo_ = {"x": d.pos_x, "y": d.pos_y, "length": d.length_x, "width": d.length_y, "row": d.rows,
                             "col": d.cols, "resolution": d.resolution}

    # FP.cpp:539-1602; returns the service's bool, fills `response`
    def globalFootholdPlan(self, request: GlobalFootholdPlanRequest, response: GlobalFootholdPlanResponse) -> bool:
        cycles = int(request.gait_cycles) & 0xFF  # uint8 in the .srv
        q = make_queries([self.initialPose_])
        nom, cen, dfl, st = self.fhp.plan_ex(self.params, q, cycles)
        # the reference aborts the call when its opt chain's gait-cycle submap does not fit the map (`return false`,
        # FP.cpp:931-934) with only the stance rows recorded; libfhp reports that predicate for cycle 0
        self.optSubmapWouldFail_ = bool(st[0]["reserved"] & _abi.FHP_STATUS_OPT_SUBMAP_FAIL)
        if self.optSubmapWouldFail_:
            st = st.copy()
            st[0]["n_rows"], st[0]["success"], st[0]["gait_cycles_succeed"] = 4, 0, 0
            self.nominalGlobalFootholdsMsg_ = rows_to_msg(nom[0], st[0])
            return False
        self.nominalGlobalFootholdsMsg_ = rows_to_msg(nom[0], st[0])
        self.centroidGlobalFootholdsMsg_ = rows_to_msg(cen[0], st[0])
        # globalFootholdsResult_.defaultFootholds (FP.cpp:666-671, :1344-1348) and footholdsKPI_ (:1357-1368, :1422-1433)
        self.defaultGlobalFootholdsMsg_ = rows_to_msg(dfl[0], st[0])
        n = int(st[0]["n_rows"])
        self.feetDistance_nominal, self.cogSpeed_nominal = self.fhp.cycle_kpi(self.params, q[0], nom[0], n)
        self.feetDistance_centroid, self.cogSpeed_centroid = self.fhp.cycle_kpi(self.params, q[0], cen[0], n)
        response.footholds = self.nominalGlobalFootholdsMsg_  # FP.cpp:1588
        return True

    # batched extension of the same service: many initial poses at once
    def globalFootholdPlanBatch(self, initial_positions, gait_cycles: int, params=None, centroid=False):
        q = initial_positions if isinstance(initial_positions, np.ndarray) and initial_positions.dtype == QUERY_DTYPE \
            else make_queries(initial_positions)
        return self.fhp.plan(self.params if params is None else params, q, int(gait_cycles), centroid=centroid)

    # FP.cpp:2001-2036 — returns (footholdResult Point, footholdValidation)
    def checkFoothold(self, center, polygon_center):
        out, valid = self.fhp.check_foothold(self.params, [center], [polygon_center])
        return Point(float(out[0]["x"]), float(out[0]["y"]), float(out[0]["z"])), bool(valid[0])

    # FP.cpp:2039-2082
    def checkDefaultFoothold(self, center) -> bool:
        return bool(self.fhp.check_default(self.params, [center])[0])

    # FP.cpp:1605-1997 — returns (resultFoothold Point, case)
    def checkFootholdUseCentroidMethod(self, defaultFoothold):
        out, case = self.fhp.check_centroid(self.params, [defaultFoothold])
        return Point(float(out[0]["x"]), float(out[0]["y"]), float(out[0]["z"])), int(case[0])

    # FP.cpp:2520-2554
    def getFootholdMeanHeight(self, center) -> float:
        return float(self.fhp.mean_height(self.params, [center])[0])

    # FP.cpp:2085-2114 -> (valid, (x, y) of the first valid cell in spiral order or None)
    def checkCandidateFoothold(self, center, polygon_center):
        out, valid = self.fhp.check_candidate(self.params, [center], [polygon_center])
        return bool(valid[0]), ((float(out[0]["x"]), float(out[0]["y"])) if valid[0] else None)

    # FP.cpp:2117-2163
    def checkCirclePolygonFoothold(self, center, polygon_center) -> bool:
        return bool(self.fhp.check_circle_polygon(self.params, [center], [polygon_center])[0])

    # FP.cpp:350-378 + :2679-2699 -> (initialFootholds [4,3], currentFootholds [4,3]), RF,RH,LH,LF
    def setFirstGait(self):
        return initial_stance(self.params, self.initialPose_)

    # FP.cpp:2166-2250 -> (polygonCenter [3], nextFootholds [4,2], searchPolygons [4,4,2])
    def getFootholdSearchGridMap(self, currentFootholds, ajustedPoseY: float = 0.0):
        return search_geometry(self.params, self.initialPose_, ajustedPoseY, currentFootholds)

    # FP.cpp:2421-2463
    def getPolygonCenter(self, RF, RH, LH, LF):
        return search_geometry(self.params, self.initialPose_, 0.0, [RF, RH, LH, LF])[0]
