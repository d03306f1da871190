from robosuite_benchmark_b200.sac import EnvReplayBuffer  # noqa: F401
