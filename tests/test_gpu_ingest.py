"""Map ingest through the message entry points: fhp_set_map_wire / fhp_set_map_msg must leave the
device in exactly the state fhp_set_map does (same plans, same geometry), follow fromMessage's acceptance rules, and
keep the previous map when a message is rejected."""
import struct

import numpy as np
import pytest

from quadrupedal_foothold_planner_b200 import _abi, synth
from quadrupedal_foothold_planner_b200.gridmap_msg import (COLUMN_LABEL, ROW_LABEL, GridMapMessage, serialize_gridmap)
from quadrupedal_foothold_planner_b200.planner import FhpError
from tests.parity import XYZ_TOL, compare_plans

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert np.array_equal(a[2], b[2])
    keep = np.arange(a[0].shape[1])[None, :] < a[2]["n_rows"][:, None]  # rows beyond n_rows are unspecified
    for k in (0, 1):
        for f in ("x", "y", "z", "row", "col", "foot_id", "gait_cycle_id", "kind"):
            assert np.array_equal(a[k][f][keep], b[k][f][keep]), f


@pytest.mark.parametrize("shape", [(300, 260, (17, 250)), (2000, 1100, (1999, 3))])  # 0.3 MB layers; 8.8 MB (two chunks)
def test_wire_and_msg_ingest_equal_set_map(fhp, oracle, shape):
    rows, cols, start = shape
    m = synth.rough_map(rows, cols, 0.02, seed=41, start=start)
    hx, hy = rows * 0.02 / 2, cols * 0.02 / 2
    q = synth.uniform_queries(256, (-hx + 1.0, hx - 3.5), (-hy + 1.0, hy - 1.0), seed=2)
    p = _abi.yaml_params()
    fhp.set_map(m)
    ref = fhp.plan(p, q, 10, centroid=True)
    assert compare_plans(ref, oracle.plan(m, p, q, 10), "set_map") <= XYZ_TOL
    wire = serialize_gridmap(m, frame_id="odom", basic_layers=["elevation"])
    # unaligned float payloads: prepend nothing, but the frame_id length already shifts them off 4-byte alignment
    fhp.set_map(synth.flat_map(64, 64, 0.02))  # something else in between
    fhp.set_map_wire(wire)
    _same(fhp.plan(p, q, 10, centroid=True), ref)
    d = fhp.map_info()
    assert (d.rows, d.cols, d.start_row, d.start_col, d.n_layers) == (rows, cols, start[0], start[1], len(m.layers))
    assert d.resolution == m.resolution and (d.length_x, d.length_y) == m.length and (d.pos_x, d.pos_y) == m.pos
    fhp.set_map(synth.flat_map(64, 64, 0.02))
    fhp.set_map_msg(GridMapMessage.from_host_map(m))
    _same(fhp.plan(p, q, 10, centroid=True), ref)


def test_periodic_updates_reuse_buffers(fhp, oracle):
    """Same-size map updates (the robot's steady state) overwrite the device buffers in place."""
    p = _abi.yaml_params()
    q = synth.uniform_queries(128, (-4.0, 1.0), (-4.0, 4.0), seed=6)
    ptrs = set()
    for seed in (51, 52, 53):
        m = synth.rough_map(600, 600, 0.02, seed=seed, start=(seed, 2 * seed))
        fhp.set_map_wire(serialize_gridmap(m))
        ptrs.add(fhp.layer_device_ptr("traversability"))
        assert compare_plans(fhp.plan(p, q, 6, centroid=True), oracle.plan(m, p, q, 6), f"update {seed}") <= XYZ_TOL
    assert len(ptrs) == 1


def test_hand_assembled_wire_bytes_reach_the_device(fhp, oracle):
    """The wire KAT assembled field by field in tests/test_gridmap_msg.py (no shared code with the serialiser): after
    fhp_set_map_wire the device holds exactly that map — geometry, start index, layer order by NAME, NaN cell."""
    from tests.test_gridmap_msg import _hand_assembled_wire
    wire, exp = _hand_assembled_wire()
    fhp.set_map_wire(wire)
    d = fhp.map_info()
    assert (d.rows, d.cols, d.start_row, d.start_col, d.n_layers) == (3, 2, 2, 1, 2)
    assert (d.resolution, d.length_x, d.length_y, d.pos_x, d.pos_y) == (0.25, 0.75, 0.5, -1.5, 2.25)
    p = _abi.yaml_params(foot_radius=0.05, search_radius=0.3)
    pts = np.array([oracle.get_position(exp, r, c)[1] for r in range(3) for c in range(2)])
    assert np.array_equal(fhp.mean_height(p, pts), oracle.mean_height(exp, p, pts))
    assert set(np.round(fhp.mean_height(p, pts) - 0.01, 4).tolist()) == {10.0, 11.0, 12.0, 13.0, 14.0, 15.0}
    assert np.array_equal(fhp.check_default(p, pts), oracle.check_default(exp, p, pts))
    assert np.array_equal(fhp.check_default(_abi.yaml_params(foot_radius=0.05, search_radius=0.3, thr_default=0.65), pts),
                          oracle.check_default(exp, _abi.yaml_params(foot_radius=0.05, search_radius=0.3, thr_default=0.65), pts))


def test_unknown_layout_label_is_read_as_column_major_like_upstream(fhp):
    """grid_map_ros's isRowMajor() logs an error for a label that is neither "column_index" nor "row_index" and answers
    false: the array is then copied as column-major.  Same here (FHP_OK; the log text is left in fhp_last_error)."""
    m = synth.rough_map(120, 90, 0.02, seed=5, start=(3, 4))
    q = synth.uniform_queries(16, (-0.5, 0.0), (-0.4, 0.4), seed=1)
    p = _abi.yaml_params()
    fhp.set_map(m)
    ref = fhp.plan(p, q, 3, centroid=True)
    msg = GridMapMessage.from_host_map(m)
    msg.data[1].dim[0].label = "cols"
    fhp.set_map(synth.flat_map(64, 64, 0.02))
    fhp.set_map_wire(msg.serialize())
    assert b"layout label is not set correctly" in fhp.L.fhp_last_error(fhp.h)
    _same(fhp.plan(p, q, 3, centroid=True), ref)


def test_rejected_messages_keep_the_previous_map(fhp):
    p = _abi.yaml_params()
    good = synth.rough_map(200, 180, 0.02, seed=61, start=(7, 11))
    q = synth.uniform_queries(32, (-1.0, 0.0), (-1.0, 1.0), seed=1)
    fhp.set_map(good)
    ref = fhp.plan(p, q, 4, centroid=True)
    other = synth.rough_map(200, 180, 0.02, seed=62, start=(0, 0))

    def rejected(wire, match):
        with pytest.raises(FhpError, match=match) as e:
            fhp.set_map_wire(wire)
        assert e.value.code == _abi.FHP_ERR_INVALID_ARG
        _same(fhp.plan(p, q, 4, centroid=True), ref)

    msg = GridMapMessage.from_host_map(other)
    msg.data[0].dim[0].label, msg.data[0].dim[1].label = ROW_LABEL, COLUMN_LABEL  # row-major announcement
    rejected(msg.serialize(), "storage order is not compatible")
    msg = GridMapMessage.from_host_map(other)
    msg.layers.append("extra")
    rejected(msg.serialize(), "Different number of layers and data")
    msg = GridMapMessage.from_host_map(other)
    msg.length_x += 0.02  # geometry says 201 rows, the arrays hold 200
    rejected(msg.serialize(), "does not have the size of the map geometry")
    msg = GridMapMessage.from_host_map(other)
    msg.outer_start_index = 200
    rejected(msg.serialize(), "start index outside the map")
    wire = serialize_gridmap(other)
    rejected(wire[:-3], "truncated or inconsistent")
    rejected(wire[: len(wire) // 2], "truncated or inconsistent")
    rejected(wire + b"\0", "trailing bytes")
    huge = bytearray(wire)
    off = 12 + 4 + len("odom") + 8 * 10
    huge[off:off + 4] = struct.pack("<I", 0xFFFFFFF0)  # layers[] count far beyond the buffer
    rejected(bytes(huge), "truncated or inconsistent")
    # a message without the layers the planner needs is accepted by fromMessage and refused like GridMap::at would
    lone = _abi.HostMap(50, 50, 0.02, layers={"slope": np.zeros((50, 50), np.float32)})
    with pytest.raises(FhpError) as e:
        fhp.set_map_wire(serialize_gridmap(lone))
    assert e.value.code == _abi.FHP_ERR_MISSING_LAYER


def test_wire_parser_survives_mutations(fhp):
    """Random truncations, byte flips and length-field corruptions of a valid message: every call returns (accepted
    or rejected with INVALID_ARG / MISSING_LAYER) — no crash, no out-of-bounds read — and the planner stays usable."""
    m = synth.rough_map(48, 40, 0.02, seed=5, start=(3, 7))
    wire = bytearray(serialize_gridmap(m, frame_id="odom", basic_layers=["elevation"]))
    rng = np.random.default_rng(2)
    header_end = 12 + 4 + 4 + 80 + 200  # header, geometry, names and the first array's layout live in these bytes
    outcomes = {"ok": 0, "rejected": 0}
    for trial in range(400):
        b = bytearray(wire)
        kind = trial % 4
        if kind == 0:
            b = b[: int(rng.integers(0, len(b)))]
        elif kind == 1:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, header_end))] = int(rng.integers(0, 256))
        elif kind == 2:
            off = int(rng.integers(0, header_end - 4))
            b[off:off + 4] = struct.pack("<I", int(rng.choice([0, 1, 0x7FFFFFFF, 0xFFFFFFFF, len(b), len(b) * 4])))
        else:
            b += bytes(int(rng.integers(1, 9)))
        try:
            fhp.set_map_wire(bytes(b))
            outcomes["ok"] += 1
        except FhpError as e:
            assert e.code in (_abi.FHP_ERR_INVALID_ARG, _abi.FHP_ERR_MISSING_LAYER), e
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 250
    fhp.set_map_wire(bytes(wire))
    q = synth.uniform_queries(16, (-0.2, 0.0), (-0.2, 0.2), seed=1)
    st = fhp.plan(_abi.yaml_params(), q, 2)[2]
    assert np.all((st["reserved"] & 0x7F) == 0)
