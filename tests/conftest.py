import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def matfiles(tmp_path_factory):
    """Synthetic material files in the reference's on-disk format (material.cpp:86-114)."""
    from montecarlocpp_b200 import materials
    d = tmp_path_factory.mktemp("materials")
    files = materials.write_all(str(d), nw=1000)
    files["silicon_small"] = materials.write_silicon(str(tmp_path_factory.mktemp("si_small")), nw=64)
    return files


@pytest.fixture(scope="session")
def omats(matfiles):
    """Oracle Material objects."""
    from oracle import pyoracle as orc
    return {k: orc.Material(*v) for k, v in matfiles.items()}


@pytest.fixture(scope="session")
def gpu_ctx():
    from montecarlocpp_b200 import capi
    ctx = capi.Context(0)          # raises (no fallback) when there is no sm_100 device
    yield ctx
    ctx.close()
