"""`from robosuite.controllers import load_controller_config, ALL_CONTROLLERS` (util/rlkit_utils.py:22, scripts/rollout.py:5)."""
from robosuite_benchmark_b200.controllers import ALL_CONTROLLERS, load_controller_config  # noqa: F401
