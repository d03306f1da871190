from robosuite_benchmark_b200.rlkit_api import TanhMlpPolicy  # noqa: F401
from robosuite_benchmark_b200.sac import FlattenMlp  # noqa: F401
