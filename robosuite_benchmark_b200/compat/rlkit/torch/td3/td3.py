from robosuite_benchmark_b200.rlkit_api import TD3Trainer  # noqa: F401
