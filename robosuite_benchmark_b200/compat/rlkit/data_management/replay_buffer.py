from robosuite_benchmark_b200.rlkit_api import ReplayBuffer  # noqa: F401
