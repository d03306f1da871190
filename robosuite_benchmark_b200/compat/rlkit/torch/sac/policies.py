from robosuite_benchmark_b200.sac import MakeDeterministic, TanhGaussianPolicy  # noqa: F401
