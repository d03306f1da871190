import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lift_panda_osc():
    from robosuite_benchmark_b200.controllers import load_controller_config
    from robosuite_benchmark_b200.model.tasks import build_task
    return build_task("Lift", "Panda", load_controller_config(default_controller="OSC_POSE"), ignore_done=True)
