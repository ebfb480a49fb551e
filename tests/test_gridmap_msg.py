"""grid_map_msgs/GridMap host tooling (message mirror + ROS1 wire form): round trips, a hand-checked known-answer
byte string, the fromMessage geometry rules, and the ctypes views of the ingest structs.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from quadrupedal_foothold_planner_b200 import _abi, synth
from quadrupedal_foothold_planner_b200.gridmap_msg import (COLUMN_LABEL, ROW_LABEL, Float32MultiArray, GridMapMessage,
                                                           MultiArrayDimension, serialize_gridmap)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# 2x1 map, res 0.5, position (1,-2), start (1,0), one layer "t" = [1,2], frame "map", seq 7, stamp (3,4), basic_layers
# ["t"].  Checked field by field against the message definitions and ROS1's serialisation rules (little-endian,
# uint32 length prefixes, definition order).
KAT_HEX = (
    "070000000300000004000000" "03000000" "6d6170"                      # header: seq, stamp, frame_id
    "000000000000e03f" "000000000000f03f" "000000000000e03f"            # resolution 0.5, length_x 1.0, length_y 0.5
    "000000000000f03f" "00000000000000c0" "0000000000000000"            # pose.position 1, -2, 0
    "0000000000000000" "0000000000000000" "0000000000000000" "000000000000f03f"  # pose.orientation 0,0,0,1
    "01000000" "01000000" "74"                                          # layers = ["t"]
    "01000000" "01000000" "74"                                          # basic_layers = ["t"]
    "01000000"                                                          # data[] count
    "02000000"                                                          #   layout.dim[] count
    "0c000000" "636f6c756d6e5f696e646578" "01000000" "02000000"         #   "column_index", size 1 (cols), stride 2
    "09000000" "726f775f696e646578" "02000000" "02000000"               #   "row_index", size 2 (rows), stride 2
    "00000000"                                                          #   data_offset
    "02000000" "0000803f" "00000040"                                    #   data = [1.0, 2.0]
    "0100" "0000")                                                      # outer_start_index 1, inner_start_index 0


def _kat_message():
    m = _abi.HostMap(2, 1, 0.5, pos=(1.0, -2.0), start=(1, 0), layers={"t": np.array([[1.0], [2.0]], np.float32)})
    msg = GridMapMessage.from_host_map(m, frame_id="map", basic_layers=["t"])
    msg.seq, msg.stamp = 7, (3, 4)
    return m, msg


def test_known_answer_bytes():
    _, msg = _kat_message()
    assert msg.serialize().hex() == KAT_HEX
    back = GridMapMessage.deserialize(bytes.fromhex(KAT_HEX))
    assert back.frame_id == "map" and back.layers == ["t"] and back.basic_layers == ["t"]
    assert [(d.label, d.size, d.stride) for d in back.data[0].dim] == [(COLUMN_LABEL, 1, 2), (ROW_LABEL, 2, 2)]
    assert back.data[0].data.tolist() == [1.0, 2.0] and (back.outer_start_index, back.inner_start_index) == (1, 0)


def _hand_assembled_wire():
    """A second known answer, assembled HERE field by field from the message definitions with struct.pack — it shares no
    code with gridmap_msg.py's serialiser.  3x2 map, two layers in the order (elevation, traversability), empty
    frame_id, basic_layers != layers, non-zero strides / data_offset / start indices.
      std_msgs/Header            uint32 seq, time stamp (uint32 secs, uint32 nsecs), string frame_id
      grid_map_msgs/GridMapInfo  Header header, float64 resolution, float64 length_x, float64 length_y, geometry_msgs/Pose pose
      geometry_msgs/Pose         Point position (3 x float64), Quaternion orientation (4 x float64)
      grid_map_msgs/GridMap      GridMapInfo info, string[] layers, string[] basic_layers,
                                 std_msgs/Float32MultiArray[] data, uint16 outer_start_index, uint16 inner_start_index
      std_msgs/Float32MultiArray MultiArrayLayout layout {MultiArrayDimension[] dim {string label, uint32 size,
                                 uint32 stride}, uint32 data_offset}, float32[] data
    ROS1 wire rules: little-endian, fields in definition order, uint32 length before strings and variable arrays."""
    import struct

    def string(t):
        b = t.encode()
        return struct.pack("<I", len(b)) + b

    rows, cols, res = 3, 2, 0.25
    elev = np.array([[10.0, 13.0], [11.0, 14.0], [12.0, 15.0]], np.float32)   # (row, col); column-major on the wire
    trav = np.array([[0.5, 0.8], [0.6, np.nan], [0.7, 1.0]], np.float32)
    w = struct.pack("<III", 42, 1600000000, 123456789) + string("")            # header
    w += struct.pack("<ddd", res, rows * res, cols * res)                      # resolution, length_x, length_y
    w += struct.pack("<ddd", -1.5, 2.25, 9.0) + struct.pack("<dddd", 0.0, 0.0, 0.0, 1.0)  # pose
    w += struct.pack("<I", 2) + string("elevation") + string("traversability")  # layers
    w += struct.pack("<I", 1) + string("elevation")                              # basic_layers
    w += struct.pack("<I", 2)                                                    # data[]
    for layer in (elev, trav):
        w += struct.pack("<I", 2)                                                # layout.dim[]
        w += string("column_index") + struct.pack("<II", cols, rows * cols)
        w += string("row_index") + struct.pack("<II", rows, rows)
        w += struct.pack("<I", 0)                                                # data_offset
        flat = [float(layer[r, c]) for c in range(cols) for r in range(rows)]    # Eigen column-major: rows contiguous
        w += struct.pack("<I", len(flat)) + struct.pack("<%df" % len(flat), *flat)
    w += struct.pack("<HH", 2, 1)                                                # outer / inner start index
    expect = _abi.HostMap(rows, cols, res, pos=(-1.5, 2.25), start=(2, 1), layers={"elevation": elev, "traversability": trav})
    return w, expect


def test_hand_assembled_wire_is_parsed_field_by_field():
    wire, exp = _hand_assembled_wire()
    msg = GridMapMessage.deserialize(wire)
    assert (msg.seq, msg.stamp, msg.frame_id) == (42, (1600000000, 123456789), "")
    assert (msg.resolution, msg.length_x, msg.length_y) == (0.25, 0.75, 0.5)
    assert msg.position == (-1.5, 2.25, 9.0) and msg.orientation == (0.0, 0.0, 0.0, 1.0)
    assert msg.layers == ["elevation", "traversability"] and msg.basic_layers == ["elevation"]
    assert (msg.outer_start_index, msg.inner_start_index) == (2, 1)
    got = msg.to_host_map()
    assert (got.rows, got.cols, got.start, got.pos, got.resolution) == (3, 2, (2, 1), (-1.5, 2.25), 0.25)
    for k in ("elevation", "traversability"):
        assert np.array_equal(got.layers[k].view(np.uint32), exp.layers[k].view(np.uint32))
    assert msg.serialize() == wire  # and the serialiser reproduces the hand-assembled bytes exactly


def test_round_trip_keeps_every_bit():
    m = synth.rough_map(37, 53, 0.02, seed=9, start=(5, 50))
    wire = serialize_gridmap(m, frame_id="odom")
    back = GridMapMessage.deserialize(wire).to_host_map()
    assert (back.rows, back.cols, back.start, back.pos) == (m.rows, m.cols, m.start, m.pos)
    assert back.resolution == m.resolution and back.length == m.length
    for k in m.layers:
        assert np.array_equal(back.layers[k].view(np.uint32), m.layers[k].view(np.uint32))  # NaN payloads included
    assert back.layers["traversability"].flags["F_CONTIGUOUS"]


def test_from_message_geometry_rules():
    """GridMap::setGeometry: size = round(length/resolution); length is re-derived as size*resolution."""
    msg = GridMapMessage(resolution=0.02, length_x=1.009, length_y=0.611, layers=[], data=[])
    m = msg.to_host_map()
    assert (m.rows, m.cols) == (50, 31) and m.length == (50 * 0.02, 31 * 0.02)
    bad = GridMapMessage(resolution=0.5, length_x=1.0, length_y=0.5, layers=["t"],
                         data=[Float32MultiArray(dim=[MultiArrayDimension(ROW_LABEL, 2, 2),
                                                      MultiArrayDimension(COLUMN_LABEL, 1, 1)],
                                                 data=np.zeros(2, np.float32))])
    with pytest.raises(ValueError, match="storage order"):
        bad.to_host_map()
    with pytest.raises(ValueError, match="Different number"):
        GridMapMessage(resolution=0.5, length_x=1.0, length_y=0.5, layers=["a", "b"], data=[]).to_host_map()


def test_ingest_struct_layouts_match_header(tmp_path):
    c = tmp_path / "layout.c"
    c.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "fhp.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(fhp_str), sizeof(fhp_multiarray), sizeof(fhp_gridmap_msg));
  printf("%zu %zu %zu %zu\n", offsetof(fhp_multiarray, dim1_label), offsetof(fhp_multiarray, data),
         offsetof(fhp_gridmap_msg, layers), offsetof(fhp_gridmap_msg, outer_start_index));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    v = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert v[0] == C.sizeof(_abi.Str) and v[1] == C.sizeof(_abi.MultiArray) and v[2] == C.sizeof(_abi.GridMapMsg)
    assert v[3] == _abi.MultiArray.dim1_label.offset and v[4] == _abi.MultiArray.data.offset
    assert v[5] == _abi.GridMapMsg.layers.offset and v[6] == _abi.GridMapMsg.outer_start_index.offset
