"""ctypes mirror of include/fhp.h (struct layouts + enums).  Shared by the product binding (planner.py) and by the
tests' oracle binding, so both sides are driven with byte-identical inputs."""
import ctypes as C

import numpy as np

FHP_OK = 0
FHP_ERR_INVALID_ARG = 1
FHP_ERR_NO_MAP = 2
FHP_ERR_MISSING_LAYER = 3
FHP_ERR_CUDA = 4
FHP_ERR_UNSUPPORTED = 5
FHP_ERR_NCCL = 6

FHP_RF, FHP_RH, FHP_LH, FHP_LF = 0, 1, 2, 3
FHP_KIND_STANCE, FHP_KIND_DEFAULT, FHP_KIND_CANDIDATE = 0, 1, 2
FHP_GAIT_TROT, FHP_GAIT_WALK, FHP_GAIT_CUSTOM = 0, 1, 2
FHP_SPIRAL_WRAP_CORRECT, FHP_SPIRAL_FAITHFUL = 0, 1


class MapDesc(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("resolution", C.c_double),
        ("length_x", C.c_double), ("length_y", C.c_double),
        ("pos_x", C.c_double), ("pos_y", C.c_double),
        ("start_row", C.c_int32), ("start_col", C.c_int32),
        ("n_layers", C.c_int32),
        ("layer_names", C.POINTER(C.c_char_p)),
    ]


class Params(C.Structure):
    _fields_ = [
        ("foot_radius", C.c_float), ("search_radius", C.c_float), ("step_length", C.c_float), ("skew", C.c_float),
        ("thr_default", C.c_float), ("thr_candidate", C.c_float),
        ("kin_length", C.c_float), ("kin_width", C.c_float), ("kin_l1", C.c_float),
        ("rf_first", C.c_int32), ("gait", C.c_int32), ("spiral_mode", C.c_int32),
        ("h", C.c_double), ("drift_per_cycle", C.c_double),
        ("bias", (C.c_double * 2) * 4),
        ("heading", C.c_double),
    ]


class Str(C.Structure):  # fhp_str
    _fields_ = [("p", C.c_char_p), ("n", C.c_uint32)]


class MultiArray(C.Structure):  # fhp_multiarray (std_msgs/Float32MultiArray)
    _fields_ = [("dim0_label", Str), ("dim0_size", C.c_uint32), ("dim1_label", Str), ("dim1_size", C.c_uint32),
                ("n_dims", C.c_uint32), ("reserved", C.c_uint32), ("data", C.c_void_p), ("n", C.c_uint64)]


class GridMapMsg(C.Structure):  # fhp_gridmap_msg (deserialised grid_map_msgs/GridMap)
    _fields_ = [("resolution", C.c_double), ("length_x", C.c_double), ("length_y", C.c_double),
                ("pos_x", C.c_double), ("pos_y", C.c_double), ("n_layers", C.c_int32), ("n_data", C.c_int32),
                ("layers", C.POINTER(Str)), ("data", C.POINTER(MultiArray)),
                ("outer_start_index", C.c_uint16), ("inner_start_index", C.c_uint16), ("reserved", C.c_uint32)]


# numpy views of the array-of-struct buffers (zero-copy between numpy, ctypes and pinned torch tensors)
QUERY_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("param_id", "<i4"), ("reserved", "<i4")])
FOOTHOLD_DTYPE = np.dtype([
    ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("row", "<i4"), ("col", "<i4"),
    ("foot_id", "u1"), ("gait_cycle_id", "u1"), ("kind", "u1"), ("reserved0", "u1"), ("reserved1", "<u4"),
])
FOOTHOLD_PACKED_DTYPE = np.dtype([  # fhp_foothold_packed (32 bytes)
    ("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("row", "<u2"), ("col", "<u2"),
    ("foot_id", "u1"), ("gait_cycle_id", "u1"), ("kind", "u1"), ("reserved", "u1"),
])
FOOTHOLD24_DTYPE = np.dtype([  # fhp_foothold24 (24 bytes): the Foothold message fields
    ("x", "<f8"), ("y", "<f8"), ("z", "<f4"), ("foot_id", "u1"), ("gait_cycle_id", "u1"), ("kind", "u1"), ("reserved", "u1"),
])
POINT20_DTYPE = np.dtype({"names": ["x", "y", "z"], "formats": ["<f8", "<f8", "<f4"], "offsets": [0, 8, 16],
                          "itemsize": 20})  # fhp_point20 (packed)
STREAM_MASK_DTYPE = np.dtype([("success", "<u4"), ("candidate", "<u4")])  # fhp_stream_mask
STATUS_DTYPE = np.dtype([("success", "u1"), ("gait_cycles", "u1"), ("gait_cycles_succeed", "u1"), ("reserved", "u1"),
                         ("n_rows", "<i4")])
assert QUERY_DTYPE.itemsize == 32 and FOOTHOLD_DTYPE.itemsize == 40 and STATUS_DTYPE.itemsize == 8
assert FOOTHOLD_PACKED_DTYPE.itemsize == 32 and FOOTHOLD24_DTYPE.itemsize == 24
# fhp_query_status.reserved: low bits = kernel error bits (any set bit fails the call), 0x80 = FHP_STATUS_OPT_SUBMAP_FAIL
FHP_STATUS_ERR_MASK = 0x7F
FHP_STATUS_OPT_SUBMAP_FAIL = 0x80


def rows_per_query(gait_cycles: int) -> int:
    return 4 + 4 * int(gait_cycles)


def yaml_params(**over) -> Params:
    """foothold_planner/config/foothold_planner.yaml:17-44 (trot, RF_FIRST false); h_/drift from FP.cpp:336,:1578."""
    p = Params()
    p.foot_radius = 0.02
    p.search_radius = 0.1
    p.step_length = 0.18
    p.skew = 0.04
    p.thr_default = 0.9
    p.thr_candidate = 0.7
    p.kin_length = 0.4387
    p.kin_width = 0.175
    p.kin_l1 = 0.037
    p.rf_first = 0
    p.gait = FHP_GAIT_TROT
    p.spiral_mode = FHP_SPIRAL_WRAP_CORRECT
    p.h = 0.01
    p.drift_per_cycle = -0.007
    p.heading = 0.0
    for k, v in over.items():
        if k == "bias":
            for f in range(4):
                p.bias[f][0] = float(v[f][0])
                p.bias[f][1] = float(v[f][1])
        else:
            setattr(p, k, v)
    return p


class HostMap:
    """A grid_map::GridMap on the host: geometry + named column-major float32 layers.

    layers[name] is a numpy array of shape (rows, cols), dtype float32, Fortran order (Eigen::MatrixXf layout:
    element (r,c) at c*rows + r)."""

    def __init__(self, rows, cols, resolution, pos=(0.0, 0.0), start=(0, 0), layers=None):
        self.rows, self.cols = int(rows), int(cols)
        self.resolution = float(resolution)
        # GridMap::setGeometry: length = size * resolution
        self.length = (self.rows * self.resolution, self.cols * self.resolution)
        self.pos = (float(pos[0]), float(pos[1]))
        self.start = (int(start[0]), int(start[1]))
        self.layers = {}
        for k, v in (layers or {}).items():
            self.set_layer(k, v)

    def set_layer(self, name, arr):
        a = np.asfortranarray(arr, dtype=np.float32)
        assert a.shape == (self.rows, self.cols), (a.shape, self.rows, self.cols)
        self.layers[name] = a

    def desc(self, names=None):
        names = list(self.layers.keys()) if names is None else list(names)
        d = MapDesc()
        d.rows, d.cols = self.rows, self.cols
        d.resolution = self.resolution
        d.length_x, d.length_y = self.length
        d.pos_x, d.pos_y = self.pos
        d.start_row, d.start_col = self.start
        d.n_layers = len(names)
        arr = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
        d.layer_names = C.cast(arr, C.POINTER(C.c_char_p))
        ptrs = (C.POINTER(C.c_float) * max(1, len(names)))()
        for i, n in enumerate(names):
            ptrs[i] = self.layers[n].ctypes.data_as(C.POINTER(C.c_float))
        keep = (arr, ptrs, [self.layers[n] for n in names])
        return d, ptrs, keep

    def relaid_start0(self):
        """The same map content re-laid-out with start index (0,0): unwrapped cell (u0,u1) = buffer
        ((u0+s0)%rows, (u1+s1)%cols)."""
        out = HostMap(self.rows, self.cols, self.resolution, self.pos, (0, 0))
        for k, v in self.layers.items():
            out.set_layer(k, np.roll(v, shift=(-self.start[0], -self.start[1]), axis=(0, 1)))
        return out


def make_queries(xyz, param_id=0):
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    q = np.zeros(xyz.shape[0], dtype=QUERY_DTYPE)
    q["x"], q["y"], q["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    q["param_id"] = param_id
    return q
