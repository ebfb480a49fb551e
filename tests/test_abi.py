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
