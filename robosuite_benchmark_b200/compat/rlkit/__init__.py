"""rlkit's import paths as the reference uses them, re-exporting robosuite_benchmark_b200's CUDA-backed implementation."""
from robosuite_benchmark_b200.sac import register_safe_globals as _reg

_reg()          # snapshots written by this backend unpickle through the reference's plain torch.load(params.pkl)
