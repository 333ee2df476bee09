import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "abstract_golden.npz"))


@pytest.fixture(scope="session")
def cuda_lib():
    """Builds (if needed) and loads libmrb200.so; GPU tests fail loudly if it is missing."""
    from multirobot_pathplanning_benchmark_b200 import build, _lib
    build.build()
    return _lib.load()


REFERENCE_SRC = "/root/reference/src"


@pytest.fixture(scope="session")
def reference():
    """Makes the (read-only) reference importable, with MagicMock stubs for its GUI / rai
    dependencies (recipe: SURVEY.md 8c).  Only exists in the build container; tests that need it
    are skipped elsewhere (nothing under -m gpu depends on it)."""
    if not os.path.isdir(REFERENCE_SRC):
        pytest.skip("reference checkout not available")
    from unittest.mock import MagicMock
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches", "matplotlib.collections",
              "mpl_toolkits", "mpl_toolkits.mplot3d", "robotic", "simple_parsing"):
        sys.modules.setdefault(n, MagicMock())
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import multi_robot_multi_goal_planning.problems as problems
    return problems
