"""pytest configuration: markers, repo root on sys.path, shared fixtures."""

import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device should fail loudly, not skip: the product has
    # no CPU path.  `-m "not gpu"` never reaches these items.
    pass


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        with np.load(GOLDEN / name) as z:
            return {k: z[k] for k in z.files}

    return _load
