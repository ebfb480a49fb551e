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


@pytest.fixture(scope="session", params=["tma", "ldg"])
def fhp(request):
    """One libfhp context on cuda:0 per window-staging path (GPU tests only): "tma" = 2-D TMA box loads (default),
    "ldg" = FHP_DISABLE_TMA=1, the pipelined LDG path that also serves maps TMA cannot describe."""
    from quadrupedal_foothold_planner_b200.planner import Fhp
    old = os.environ.get("FHP_DISABLE_TMA")
    os.environ["FHP_DISABLE_TMA"] = "1" if request.param == "ldg" else "0"
    ctx = Fhp(0)
    if old is None:
        del os.environ["FHP_DISABLE_TMA"]
    else:
        os.environ["FHP_DISABLE_TMA"] = old
    yield ctx
    ctx.close()
