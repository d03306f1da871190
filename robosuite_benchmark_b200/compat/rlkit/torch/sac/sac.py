from robosuite_benchmark_b200.sac import SACTrainer  # noqa: F401
