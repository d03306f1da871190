"""`import rlkit.torch.pytorch_util as ptu` (scripts/train.py:7,101; util/rlkit_utils.py:3,6,162): set_gpu_mode, device."""
from robosuite_benchmark_b200.rlkit_api import ptu as _ptu


def set_gpu_mode(mode, gpu_id=0):
    _ptu.set_gpu_mode(mode, gpu_id)


def gpu_enabled():
    return _ptu.gpu_enabled()


def __getattr__(name):
    if name == "device":
        return _ptu.device
    raise AttributeError(name)
