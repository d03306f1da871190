from robosuite_benchmark_b200.rlkit_api import GaussianStrategy  # noqa: F401
