"""`import robosuite as suite` (util/rlkit_utils.py:19, scripts/rollout.py:3): suite.make / load_controller_config on the batched CUDA backend."""
from robosuite_benchmark_b200 import ALL_CONTROLLERS, load_controller_config, make  # noqa: F401
from . import controllers, wrappers  # noqa: F401

__version__ = "1.0.1+b200"
