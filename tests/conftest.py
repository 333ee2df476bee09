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


@pytest.fixture(scope="session")
def reference():
    """Makes the reference importable (offline install under baseline/_ref, else the build container's read-only
    checkout), with inert stand-ins for its GUI / rai dependencies (recipe: SURVEY.md 8c).  Tests that need it are
    skipped where neither exists."""
    from multirobot_pathplanning_benchmark_b200 import refimport
    if not refimport.ensure_reference():
        pytest.skip("reference package not available")
    import multi_robot_multi_goal_planning.problems as problems
    return problems
