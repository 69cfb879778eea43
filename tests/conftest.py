import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def coracle():
    import oracle
    return oracle.COracle()


@pytest.fixture(scope="session", params=["dense", "tile-skip"])
def engine(request):
    """Every GPU test runs in both packed-affinity formats: the dense one (library default) and the tile-skip fast
    path; the process-wide default is switched too, so engines created inside the registration classes follow."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import roman_b200
    from roman_b200 import clipperpy
    from roman_b200.clipperpy import _core
    from roman_b200.align import object_registration
    from roman_b200.engine import Engine
    skip = request.param == "tile-skip"
    roman_b200.set_default_tile_skip(skip)
    _core._pool.clear()                     # pooled engines were created under the other default
    object_registration._arun_engine = None
    e = Engine(0)
    e.default_tile_skip = skip
    yield e
    roman_b200.set_default_tile_skip(False)
