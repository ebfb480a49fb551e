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
