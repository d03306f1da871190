"""`from rlkit.core import logger, eval_util` (util/rlkit_custom.py:6, util/rlkit_utils.py:18)."""
from robosuite_benchmark_b200.rlkit_api import eval_util, logger  # noqa: F401
