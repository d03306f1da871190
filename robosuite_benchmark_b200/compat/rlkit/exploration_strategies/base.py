from robosuite_benchmark_b200.rlkit_api import PolicyWrappedWithExplorationStrategy  # noqa: F401
