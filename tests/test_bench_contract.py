"""bench.py prints ONE JSON line with the keys the driver reads — checked for the CPU reference arm here and for the
GPU arm (on the small C1 workload) under -m gpu."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample", "256", "--workload", "c2",
              "--no-as-written"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "footholds/s" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["gpu_launches"] == 0


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--workload", "c1", "--steps", "3", "--warmup", "3"], 900)
    assert (BASE_KEYS | {"clocks", "roofline"}) <= set(d) and "impl" not in d
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 3 and d["gpu_launches"] == 3
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
