"""B200-native batched implementation of the robosuite-benchmark hot path (env.step + SAC update).

The names exported here mirror what the reference imports from robosuite (util/rlkit_utils.py:19-22):
`make`, `load_controller_config`, `ALL_CONTROLLERS`, `wrappers.GymWrapper`.
"""
from .controllers import ALL_CONTROLLERS, load_controller_config  # noqa: F401


def make(*args, **kwargs):
    from .environments import make as _make
    return _make(*args, **kwargs)


__all__ = ["make", "load_controller_config", "ALL_CONTROLLERS"]
