from robosuite_benchmark_b200.rlkit_api import setup_logger  # noqa: F401
