from robosuite_benchmark_b200.rlkit_api import BaseRLAlgorithm, _get_epoch_timings  # noqa: F401
