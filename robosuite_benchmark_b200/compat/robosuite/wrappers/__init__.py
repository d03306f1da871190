"""`from robosuite.wrappers import GymWrapper` (util/rlkit_utils.py:20, scripts/rollout.py:4)."""
from robosuite_benchmark_b200.wrappers import Box, GymWrapper  # noqa: F401
