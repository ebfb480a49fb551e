import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_binding
    oracle_binding.lib()
    return oracle_binding


@pytest.fixture(scope="session", params=["mask", "mask-generic", "tma", "ldg"])
def fhp(request):
    """One libfhp context on cuda:0 per window-staging path (GPU tests only): "mask" = precomputed D/C/F mask planes
    (default; the yaml geometries run on the geometry-specialised kernels), "mask-generic" = the same path with the
    runtime-geometry kernels only (FHP_NO_SPEC=1), "tma" = 2-D TMA box loads of the float layer, "ldg" = pipelined
    coalesced loads (also serves maps TMA cannot describe)."""
    from quadrupedal_foothold_planner_b200.planner import Fhp
    old = {k: os.environ.get(k) for k in ("FHP_STAGING", "FHP_NO_SPEC")}
    os.environ["FHP_STAGING"] = request.param.split("-")[0]
    os.environ["FHP_NO_SPEC"] = "1" if request.param.endswith("generic") else "0"
    ctx = Fhp(0)
    for k, v in old.items():
        if v is None:
            del os.environ[k]
        else:
            os.environ[k] = v
    yield ctx
    ctx.close()
