"""grid_map_msgs/GridMap on the host: the message as roscpp/rospy hold it, and its ROS1 wire form.

The reference's map input is a `grid_map_msgs/GridMap` (FootholdPlanner::gridmapCallback, FP.cpp:504-506).  Neither
ROS nor grid_map is in this image, so this module restates the message definition (grid_map_msgs 1.6.x:
GridMapInfo info, string[] layers, string[] basic_layers, std_msgs/Float32MultiArray[] data, uint16
outer_start_index, uint16 inner_start_index) and ROS1's serialisation rules (little-endian scalars, uint32 length
prefix on strings and variable arrays, fields in definition order) for batch clients, rosbag readers and the tests
of `fhp_set_map_wire` / `fhp_set_map_msg`.  `GridMapRosConverter::toMessage` is mirrored by `from_host_map`:
column-major layers are announced as layout.dim[0] = ("column_index", cols, rows*cols), dim[1] = ("row_index",
rows, rows).
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from ._abi import GridMapMsg, HostMap, MultiArray, Str

COLUMN_LABEL = "column_index"  # grid_map_ros storageIndexNames[StorageIndices::Column]
ROW_LABEL = "row_index"


@dataclass
class MultiArrayDimension:  # std_msgs/MultiArrayDimension
    label: str = ""
    size: int = 0
    stride: int = 0


@dataclass
class Float32MultiArray:  # std_msgs/Float32MultiArray
    dim: List[MultiArrayDimension] = field(default_factory=list)
    data_offset: int = 0
    data: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))


@dataclass
class GridMapMessage:  # grid_map_msgs/GridMap
    seq: int = 0
    stamp: Tuple[int, int] = (0, 0)
    frame_id: str = "odom"
    resolution: float = 0.0
    length_x: float = 0.0
    length_y: float = 0.0
    position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    orientation: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 1.0)
    layers: List[str] = field(default_factory=list)
    basic_layers: List[str] = field(default_factory=list)
    data: List[Float32MultiArray] = field(default_factory=list)
    outer_start_index: int = 0
    inner_start_index: int = 0

    # -- GridMapRosConverter::toMessage
    @staticmethod
    def from_host_map(m: HostMap, frame_id: str = "odom", basic_layers=()) -> "GridMapMessage":
        msg = GridMapMessage(frame_id=frame_id, resolution=m.resolution, length_x=m.length[0], length_y=m.length[1],
                             position=(m.pos[0], m.pos[1], 0.0), layers=list(m.layers.keys()),
                             basic_layers=list(basic_layers), outer_start_index=m.start[0],
                             inner_start_index=m.start[1])
        for name in msg.layers:
            a = m.layers[name]  # (rows, cols) Fortran order == Eigen column-major storage
            arr = Float32MultiArray(
                dim=[MultiArrayDimension(COLUMN_LABEL, m.cols, m.rows * m.cols),
                     MultiArrayDimension(ROW_LABEL, m.rows, m.rows)],
                data_offset=0, data=np.ravel(a, order="F"))
            msg.data.append(arr)
        return msg

    # -- ROS1 serialisation
    def serialize(self) -> bytes:
        out = bytearray()

        def s(x: str):
            b = x.encode()
            out.extend(struct.pack("<I", len(b)))
            out.extend(b)

        out.extend(struct.pack("<III", self.seq, self.stamp[0], self.stamp[1]))
        s(self.frame_id)
        out.extend(struct.pack("<ddd", self.resolution, self.length_x, self.length_y))
        out.extend(struct.pack("<ddd", *self.position))
        out.extend(struct.pack("<dddd", *self.orientation))
        for names in (self.layers, self.basic_layers):
            out.extend(struct.pack("<I", len(names)))
            for n in names:
                s(n)
        out.extend(struct.pack("<I", len(self.data)))
        for a in self.data:
            out.extend(struct.pack("<I", len(a.dim)))
            for d in a.dim:
                s(d.label)
                out.extend(struct.pack("<II", d.size, d.stride))
            out.extend(struct.pack("<I", a.data_offset))
            flat = np.ascontiguousarray(a.data, dtype="<f4")
            out.extend(struct.pack("<I", flat.size))
            out.extend(flat.tobytes())
        out.extend(struct.pack("<HH", self.outer_start_index, self.inner_start_index))
        return bytes(out)

    @staticmethod
    def deserialize(buf: bytes) -> "GridMapMessage":
        o = 0

        def take(fmt):
            nonlocal o
            v = struct.unpack_from(fmt, buf, o)
            o += struct.calcsize(fmt)
            return v

        def s():
            nonlocal o
            (n,) = take("<I")
            v = buf[o:o + n].decode()
            o += n
            return v

        m = GridMapMessage()
        m.seq, a, b = take("<III")
        m.stamp = (a, b)
        m.frame_id = s()
        m.resolution, m.length_x, m.length_y = take("<ddd")
        m.position = take("<ddd")
        m.orientation = take("<dddd")
        m.layers = [s() for _ in range(take("<I")[0])]
        m.basic_layers = [s() for _ in range(take("<I")[0])]
        for _ in range(take("<I")[0]):
            arr = Float32MultiArray()
            for _ in range(take("<I")[0]):
                label = s()
                size, stride = take("<II")
                arr.dim.append(MultiArrayDimension(label, size, stride))
            (arr.data_offset,) = take("<I")
            (n,) = take("<I")
            arr.data = np.frombuffer(buf, dtype="<f4", count=n, offset=o).copy()
            o += 4 * n
            m.data.append(arr)
        m.outer_start_index, m.inner_start_index = take("<HH")
        assert o == len(buf), "trailing bytes"
        return m

    # -- GridMapRosConverter::fromMessage (host-side mirror, used to build the oracle's input)
    def to_host_map(self) -> HostMap:
        rows = int(round(self.length_x / self.resolution))
        cols = int(round(self.length_y / self.resolution))
        if len(self.layers) != len(self.data):
            raise ValueError("Different number of layers and data in grid map message.")
        m = HostMap(rows, cols, self.resolution, pos=(self.position[0], self.position[1]),
                    start=(self.outer_start_index, self.inner_start_index))
        for name, a in zip(self.layers, self.data):
            if a.dim[0].label != COLUMN_LABEL:
                raise ValueError("multiArrayMessageToMatrixEigen() failed because the storage order is not compatible.")
            if a.dim[0].size != cols or a.dim[1].size != rows:
                raise ValueError("layer size differs from the geometry")
            m.set_layer(name, np.asarray(a.data, dtype=np.float32).reshape((rows, cols), order="F"))
        return m

    # -- fhp_gridmap_msg view (pointers into this object's buffers; keep the returned tuple alive)
    def as_c(self):
        keep = []
        names = (Str * max(1, len(self.layers)))()
        for i, n in enumerate(self.layers):
            b = n.encode()
            keep.append(b)
            names[i] = Str(b, len(b))
        arrs = (MultiArray * max(1, len(self.data)))()
        for i, a in enumerate(self.data):
            flat = np.ascontiguousarray(a.data, dtype=np.float32)
            keep.append(flat)
            ma = MultiArray()
            labels = [d.label.encode() for d in a.dim[:2]] + [b"", b""]
            keep.extend(labels)
            ma.dim0_label = Str(labels[0], len(labels[0]))
            ma.dim1_label = Str(labels[1], len(labels[1]))
            ma.dim0_size = a.dim[0].size if len(a.dim) > 0 else 0
            ma.dim1_size = a.dim[1].size if len(a.dim) > 1 else 0
            ma.n_dims = len(a.dim)
            ma.data = flat.ctypes.data
            ma.n = flat.size
            arrs[i] = ma
        c = GridMapMsg()
        c.resolution, c.length_x, c.length_y = self.resolution, self.length_x, self.length_y
        c.pos_x, c.pos_y = self.position[0], self.position[1]
        c.n_layers, c.n_data = len(self.layers), len(self.data)
        c.layers = C.cast(names, C.POINTER(Str))
        c.data = C.cast(arrs, C.POINTER(MultiArray))
        c.outer_start_index, c.inner_start_index = self.outer_start_index, self.inner_start_index
        keep.extend([names, arrs])
        return c, keep


def serialize_gridmap(m: HostMap, **kw) -> bytes:
    return GridMapMessage.from_host_map(m, **kw).serialize()
