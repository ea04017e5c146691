import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the product libraries and the oracle once per session (no-ops when up to date; on the
    GPU box the prebuilt .so files travel with the snapshot)."""
    from rs_pathtracing_b200 import build as rt_build
    try:
        rt_build.build_all()
    except Exception:
        # no nvcc on this box: the prebuilt libraries must already be there
        if not (os.path.exists(rt_build.CORE_SO) and os.path.exists(rt_build.HOST_SO)):
            raise
    from oracle import pyoracle
    try:
        pyoracle.build()
    except Exception:
        if not os.path.exists(pyoracle.SO):
            raise
    yield


SCENES = os.path.join(ROOT, "scenes")


def scene_path(name: str) -> str:
    return os.path.join(SCENES, name)
