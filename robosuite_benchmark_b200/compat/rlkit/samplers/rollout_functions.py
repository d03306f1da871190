from robosuite_benchmark_b200.algorithm import rollout  # noqa: F401
