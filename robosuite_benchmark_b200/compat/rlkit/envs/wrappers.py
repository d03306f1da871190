from robosuite_benchmark_b200.algorithm import NormalizedBoxEnv  # noqa: F401
