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


@pytest.fixture(scope="session", params=["mask", "tma", "ldg"])
def fhp(request):
    """One libfhp context on cuda:0 per window-staging path (GPU tests only): "mask" = precomputed D/C/F mask planes
    (default), "tma" = 2-D TMA box loads of the float layer, "ldg" = pipelined coalesced loads (also serves maps TMA
    cannot describe)."""
    from quadrupedal_foothold_planner_b200.planner import Fhp
    old = os.environ.get("FHP_STAGING")
    os.environ["FHP_STAGING"] = request.param
    ctx = Fhp(0)
    if old is None:
        del os.environ["FHP_STAGING"]
    else:
        os.environ["FHP_STAGING"] = old
    yield ctx
    ctx.close()
