import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def mano_model():
    from artiboost_b200 import assets
    return assets.make_synthetic_mano(seed=0)


@pytest.fixture(scope="session")
def objects():
    from artiboost_b200 import assets
    return assets.make_synthetic_objects(assets.HO3D_TRAIN_OBJS, seed=0)


@pytest.fixture(scope="session")
def lib_built():
    """Builds (if stale) and loads the C-ABI library; nvcc cross-compiles without a GPU."""
    from artiboost_b200 import build, lib
    build.build()
    return lib.load()
