"""The C++ host shim a ROS node would link (csrc/FootholdPlannerGpu.hpp): compiled with g++ against libfhp.so and
run as a separate process on the GPU; its response must equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

from quadrupedal_foothold_planner_b200 import _abi, planner, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, ros_mode=False):
    """ros_mode: -DFHP_WITH_ROS against tests/cpp/mock_ros (headers with the real ROS / grid_map type and field names)."""
    planner.load_lib()
    exe = str(tmp_path / ("shim_main_ros" if ros_mode else "shim_main"))
    pkg = os.path.join(ROOT, "quadrupedal_foothold_planner_b200")
    extra = ["-DFHP_WITH_ROS", "-I", os.path.join(ROOT, "tests", "cpp", "mock_ros")] if ros_mode else []
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-Werror"] + extra +
                          ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "shim_main.cpp"), "-o", exe, "-L", pkg, "-lfhp",
                           "-Wl,-rpath," + pkg])
    return exe


@pytest.mark.parametrize("ros_mode", [False, True])
def test_cpp_shim_links(tmp_path, ros_mode):
    assert os.path.exists(_build(tmp_path, ros_mode))


@pytest.mark.gpu
@pytest.mark.parametrize("ros_mode", [False, True])
def test_cpp_shim_matches_oracle(tmp_path, oracle, ros_mode):
    exe = _build(tmp_path, ros_mode)
    m = synth.rough_map(400, 300, 0.02, seed=13, start=(31, 250))
    path = str(tmp_path / "map.bin")
    with open(path, "wb") as f:
        f.write(np.asfortranarray(m.layers["traversability"]).tobytes(order="F"))
        f.write(np.asfortranarray(m.layers["elevation"]).tobytes(order="F"))
    for pose, extra in (([-2.5, 0.3, 0.0], []), ([-1.0, -1.2, 0.125], ["msg"])):  # second run: message callback
        out = subprocess.check_output([exe, path, "400", "300", "0.02", "31", "250", repr(pose[0]), repr(pose[1]),
                                       repr(pose[2]), "9"] + extra, text=True).strip().splitlines()
        p0 = _abi.yaml_params(spiral_mode=_abi.FHP_SPIRAL_FAITHFUL)  # the shim is the reference class: upstream's spiral
        nom, _, st = oracle.plan(m, p0, _abi.make_queries([pose]), 9)
        s, gc, gs, n = (int(v) for v in out[0].split())
        assert (s, gc, gs, n) == (st[0]["success"], 9, st[0]["gait_cycles_succeed"], st[0]["n_rows"])
        mem = out[1 + n].split()
        assert mem[0] == "members"
        o_init, _, o_first = oracle.initial_geometry(p0, pose)
        cur = np.concatenate([o_first, np.full((4, 1), pose[2])], axis=1)
        oc, onxt, opoly = oracle.search_geometry(p0, pose, -0.007, cur)
        vals = [float(v) for v in mem[1:9]]
        assert vals == [o_first[0, 0], o_first[3, 1], oc[0], oc[1], onxt[2, 0], onxt[2, 1], opoly[3, 0, 0], opoly[3, 2, 1]]
        ec, ev = oracle.check_candidate(m, p0, [onxt[0]], [onxt[0] + np.array([0.01, -0.02])])
        assert int(mem[9]) == ev[0] and (ev[0] == 0 or (float(mem[10]), float(mem[11])) == (ec[0]["x"], ec[0]["y"]))
        assert int(mem[12]) == oracle.check_circle_polygon(m, p0, [onxt[1]], [onxt[1]])[0]
        for line, row in zip(out[1:1 + n], nom[0][:n]):
            x, y, z, foot, cyc = line.split()
            assert abs(float(x) - row["x"]) <= 1e-5 and abs(float(y) - row["y"]) <= 1e-5 and abs(float(z) - row["z"]) <= 1e-5
            assert (int(foot), int(cyc)) == (row["foot_id"], row["gait_cycle_id"])
        # marker payloads (FP.cpp:2731-2790): per chain and foot, the points of that foot's planned rows in cycle order
        enom, ecen, edfl, _, est = oracle.plan_ex(m, p0, _abi.make_queries([pose]), 9)
        for chain, rows in enumerate((enom, ecen, edfl)):
            tok = out[2 + n + chain].split()
            assert tok[0] == "markers" and int(tok[1]) == chain
            k = 2
            for foot in range(4):
                exp = [r for r in rows[0][4:n] if r["foot_id"] == foot]
                cnt = int(tok[k])
                k += 1
                assert cnt == len(exp) == (n - 4) // 4
                for r in exp:
                    got = [float(v) for v in tok[k:k + 3]]
                    k += 3
                    assert max(abs(got[0] - r["x"]), abs(got[1] - r["y"]), abs(got[2] - r["z"])) <= 1e-5
            assert k == len(tok)
        # the reference-signature overloads agree with the short forms; a foreign polygon is refused
        assert out[5 + n].split() == ["refsig", "1", "1"]


@pytest.mark.gpu
def test_cpp_shim_reports_the_reference_abort(tmp_path, oracle):
    """A pose whose opt-chain gait submap cannot fit the map: the reference's service returns false at cycle 0
    (FP.cpp:931-934) with only the stance rows recorded; so does the shim (FHP_STATUS_OPT_SUBMAP_FAIL)."""
    exe = _build(tmp_path)
    m = synth.rough_map(200, 200, 0.02, seed=13, start=(0, 0))
    path = str(tmp_path / "map.bin")
    with open(path, "wb") as f:
        f.write(np.asfortranarray(m.layers["traversability"]).tobytes(order="F"))
        f.write(np.asfortranarray(m.layers["elevation"]).tobytes(order="F"))
    pose = [2.3, 0.4, 0.0]  # the next stance centre lies outside the 4 m x 4 m map
    st = oracle.plan(m, _abi.yaml_params(), _abi.make_queries([pose]), 3, centroid=False)[2]
    assert st[0]["reserved"] & 0x80
    out = subprocess.check_output([exe, path, "200", "200", "0.02", "0", "0", repr(pose[0]), repr(pose[1]), repr(pose[2]),
                                   "3"], text=True).strip().splitlines()
    assert out[0].split() == ["aborted", "1", "4"]
