"""The C-ABI library: builds, loads, exports every symbol include/fhp.h declares, struct layouts match the ctypes
mirror, and WITHOUT a GPU every entry point fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from quadrupedal_foothold_planner_b200 import _abi, planner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fhp.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fhp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = planner.load_lib()
    names = declared_functions()
    assert len(names) >= 15 and "fhp_plan" in names and "fhp_set_map" in names
    for n in names:
        assert hasattr(lib, n), f"libfhp.so does not export {n}"
    assert b"sm_100a" in lib.fhp_version()


def test_struct_layouts_match_header(tmp_path):
    """Compile a tiny C program against include/fhp.h and compare sizeof/offsetof with the ctypes/numpy mirror."""
    c = tmp_path / "layout.c"
    c.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "fhp.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(fhp_map_desc), sizeof(fhp_params), sizeof(fhp_query), sizeof(fhp_foothold),
         sizeof(fhp_query_status));
  if (sizeof(fhp_foothold_packed) != 32 || offsetof(fhp_foothold_packed, row) != 24 ||
      offsetof(fhp_foothold_packed, foot_id) != 28 || offsetof(fhp_foothold_packed, kind) != 30) return 1;
  if (sizeof(fhp_foothold24) != 24 || offsetof(fhp_foothold24, z) != 16 || offsetof(fhp_foothold24, foot_id) != 20 ||
      offsetof(fhp_foothold24, kind) != 22) return 2;
  if (sizeof(fhp_point20) != 20 || offsetof(fhp_point20, y) != 8 || offsetof(fhp_point20, z) != 16) return 3;
  if (sizeof(fhp_stream_mask) != 8 || offsetof(fhp_stream_mask, candidate) != 4) return 4;
  if (offsetof(fhp_params, heading) != offsetof(fhp_params, bias) + 64 || sizeof(fhp_params) != offsetof(fhp_params, heading) + 8) return 5;
  if (FHP_STREAM_BLOCKS(65536, 16) != 17 * 8192 || FHP_STREAM_MASK_WORDS(33) != 2 || FHP_STATUS_OPT_SUBMAP_FAIL != 0x80) return 6;
  printf("%zu %zu %zu %zu %zu\n", offsetof(fhp_params, h), offsetof(fhp_params, bias), offsetof(fhp_foothold, row),
         offsetof(fhp_foothold, foot_id), offsetof(fhp_query_status, n_rows));
  printf("%zu %zu %zu\n", offsetof(fhp_map_desc, resolution), offsetof(fhp_map_desc, start_row),
         offsetof(fhp_map_desc, layer_names));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = [int(x) for x in out]
    assert sizes[0] == C.sizeof(_abi.MapDesc)
    assert sizes[1] == C.sizeof(_abi.Params)
    assert sizes[2] == _abi.QUERY_DTYPE.itemsize == 32
    assert sizes[3] == _abi.FOOTHOLD_DTYPE.itemsize == 40
    assert sizes[4] == _abi.STATUS_DTYPE.itemsize == 8
    assert sizes[5] == _abi.Params.h.offset and sizes[6] == _abi.Params.bias.offset
    assert sizes[7] == _abi.FOOTHOLD_DTYPE.fields["row"][1]
    assert sizes[8] == _abi.FOOTHOLD_DTYPE.fields["foot_id"][1]
    assert sizes[9] == _abi.STATUS_DTYPE.fields["n_rows"][1]
    assert sizes[10] == _abi.MapDesc.resolution.offset and sizes[11] == _abi.MapDesc.start_row.offset
    assert sizes[12] == _abi.MapDesc.layer_names.offset


def test_new_struct_views_match():
    assert _abi.FOOTHOLD24_DTYPE.itemsize == 24 and _abi.FOOTHOLD24_DTYPE.fields["foot_id"][1] == 20
    assert _abi.POINT20_DTYPE.itemsize == 20 and _abi.POINT20_DTYPE.fields["z"][1] == 16
    assert _abi.STREAM_MASK_DTYPE.itemsize == 8
    assert _abi.Params.heading.offset == _abi.Params.bias.offset + 64 and C.sizeof(_abi.Params) == _abi.Params.heading.offset + 8


def test_stream_row_and_expander_on_hand_built_blocks():
    """fhp_stream_row / fhp_stream_expand are pure host functions: a hand-built streamed response (10 queries, 3 cycles,
    known success pattern) is located and expanded exactly (no GPU)."""
    import numpy as np
    lib = planner.load_lib()
    nq, cycles = 10, 3
    groups, words = 2, 1
    ok = np.zeros((nq, cycles), bool)
    ok[0, 0] = ok[2, 0] = ok[2, 1] = ok[7, 0] = ok[9, 2] = ok[8, 2] = ok[3, 1] = True
    blocks = np.zeros((cycles + 1, groups, 32), dtype=_abi.POINT20_DTYPE)
    masks = np.zeros((nq, words, 4), dtype=_abi.STREAM_MASK_DTYPE)
    st = np.zeros(nq, dtype=_abi.STATUS_DTYPE)
    for q in range(nq):
        for f in range(4):
            blocks[0, q // 8, (q % 8) * 4 + f] = (q, f, -1.0)
        st[q]["n_rows"] = 4 + 4 * int(ok[q].sum())
    for g in range(cycles):
        for G in range(groups):
            rank = 0
            for q in range(G * 8, min(nq, G * 8 + 8)):
                if ok[q, g]:
                    for f in range(4):
                        blocks[g + 1, G, rank * 4 + f] = (100 * g + q, f, float(g))
                        masks[q, 0, f]["success"] |= 1 << g
                        if (q + f + g) % 2:
                            masks[q, 0, f]["candidate"] |= 1 << g
                    rank += 1
    lib.fhp_stream_row.restype = C.c_void_p
    lib.fhp_stream_row.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32]
    base = blocks.ctypes.data
    for q in range(nq):
        for g in range(-1, cycles):
            for f in range(4):
                ptr = lib.fhp_stream_row(base, masks.ctypes.data, nq, cycles, q, g, f)
                if g >= 0 and not ok[q, g]:
                    assert not ptr
                    continue
                row = blocks.reshape(-1)[(ptr - base) // 20]
                assert (row["x"], row["y"]) == ((q, f) if g < 0 else (100 * g + q, f))
    rows = np.zeros(int(st["n_rows"].sum()), dtype=_abi.FOOTHOLD24_DTYPE)
    offs = np.zeros(nq + 1, dtype=np.int64)
    rc = lib.fhp_stream_expand(C.c_void_p(base), C.c_void_p(masks.ctypes.data), C.c_void_p(st.ctypes.data), C.c_int64(nq),
                               C.c_int32(cycles), C.c_void_p(rows.ctypes.data), C.c_int64(len(rows)),
                               C.c_void_p(offs.ctypes.data), C.c_int32(1))
    assert rc == 0 and offs[-1] == len(rows)
    for q in range(nq):
        r = rows[offs[q]:offs[q + 1]]
        exp_cycles = [0] * 4 + [g for g in range(cycles) if ok[q, g] for _ in range(4)]
        assert r["gait_cycle_id"].tolist() == exp_cycles and r["foot_id"].tolist() == [0, 1, 2, 3] * (len(r) // 4)
        assert r["x"][:4].tolist() == [q] * 4 and r["kind"][:4].tolist() == [0] * 4
        for i in range(4, len(r)):
            g, f = int(r["gait_cycle_id"][i]), int(r["foot_id"][i])
            assert r["x"][i] == 100 * g + q and r["z"][i] == g and r["kind"][i] == (2 if (q + f + g) % 2 else 1)
    assert lib.fhp_stream_expand(C.c_void_p(base), C.c_void_p(masks.ctypes.data), C.c_void_p(st.ctypes.data), C.c_int64(nq),
                                 C.c_int32(cycles), C.c_void_p(rows.ctypes.data), C.c_int64(len(rows) - 1),
                                 C.c_void_p(offs.ctypes.data), C.c_int32(1)) == _abi.FHP_ERR_INVALID_ARG


def test_default_params_are_the_yaml():
    lib = planner.load_lib()
    p = _abi.Params()
    lib.fhp_default_params(C.byref(p))
    y = _abi.yaml_params()
    for name, _ in _abi.Params._fields_:
        if name == "bias":
            continue
        assert getattr(p, name) == getattr(y, name), name


def test_cpp_shim_compiles():
    """The FootholdPlanner-shaped C++ host shim (what a ROS node links) compiles against the C-ABI header."""
    hpp = os.path.join(ROOT, "quadrupedal_foothold_planner_b200", "csrc", "FootholdPlannerGpu.hpp")
    src = '#include "%s"\nint main(){ foothold_planner::FootholdPlannerGpu* p = nullptr; (void)p; return 0; }\n' % hpp
    subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"],
                   input=src, text=True, check=True)


def test_no_gpu_means_loud_failure_not_fallback():
    """On a box without a usable CUDA device fhp_create must fail with FHP_ERR_CUDA (there is no CPU path).
    On a GPU box this test only checks that creation works."""
    import torch
    if torch.cuda.is_available():
        ctx = planner.Fhp(0)
        ctx.close()
        return
    with pytest.raises(planner.FhpError) as e:
        planner.Fhp(0)
    assert e.value.code == _abi.FHP_ERR_CUDA
    assert "CUDA" in str(e.value) or "device" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the product package may import, include, link or call it."""
    pkg = os.path.join(ROOT, "quadrupedal_foothold_planner_b200")
    bad = re.compile(r"(from\s+oracle|import\s+oracle|liboracle|oracle_[a-z]+\s*\(|#include\s+[\"<].*oracle|oracle[/\\])")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(txt), f"{f} references the oracle: {bad.search(txt).group(0)}"
