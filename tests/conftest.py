import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    return orc.Oracle()


def golden_names(prefix="svn_"):
    """svn_* = SVN-ICP class fixtures, svgd_* = SVGD-ICP class fixtures (different keys)."""
    d = os.path.join(ROOT, "tests", "golden")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".npz") and f.startswith(prefix))


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
