from robosuite_benchmark_b200.algorithm import MdpPathCollector  # noqa: F401
from robosuite_benchmark_b200.rlkit_api import DataCollector, PathCollector  # noqa: F401
