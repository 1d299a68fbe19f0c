import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_world():
    import oracle_lib
    return oracle_lib.World.ireland(fast=True)


@pytest.fixture(scope="session")
def gpu_ctx():
    from eirgrid_b200 import _lib
    ctx = _lib.Context(0)
    ctx.map_load_dir(os.path.join(ROOT, "tests", "golden", "ireland_map"))
    yield ctx
    ctx.close()
