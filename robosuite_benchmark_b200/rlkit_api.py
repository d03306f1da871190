"""The pieces of rlkit / gtimer that the reference's own files import beside the networks, trainer, replay buffer and collectors
(SURVEY.md 8b "import surface"): the module-level `logger` (rlkit.core.logger), `setup_logger` (rlkit.launchers.launcher_util),
`pytorch_util` (`set_gpu_mode`, `device`), `eval_util`, `_get_epoch_timings`, `pythonplusplus.list_of_dicts__to__dict_of_lists`, the gtimer
calls of util/rlkit_custom.py (`timed_for`, `stamp`, `get_times`) and the TD3 names util/rlkit_utils.py:12-15 imports.

The thin packages under `robosuite_benchmark_b200/compat/` (`robosuite`, `rlkit`, `gtimer`) re-export these under the reference's import
paths, so that `scripts/train.py`, `util/rlkit_utils.py` and `util/rlkit_custom.py` run UNMODIFIED with
`PYTHONPATH=<repo>/robosuite_benchmark_b200/compat:<repo>:<reference>` -- on the CUDA kernels, not on rlkit / robosuite / mujoco.
"""
from __future__ import annotations

import datetime
import os
import time
from collections import OrderedDict

import numpy as np

from . import algorithm as _alg


# ----------------------------------------------------------------------------- rlkit.core.logger
class RlkitLogger:
    """rlkit.core.logger (module-level singleton there): tabular rows -> progress.csv, text -> debug.log, snapshots -> params.pkl."""

    def __init__(self):
        self._impl = None
        self._snapshot_dir = None

    def _need(self):
        if self._impl is None:
            raise RuntimeError("logger used before setup_logger(...)")
        return self._impl

    def setup(self, log_dir, variant=None):
        self._impl = _alg.Logger(log_dir, variant)
        self._snapshot_dir = log_dir

    def get_snapshot_dir(self):
        return self._snapshot_dir

    def log(self, s, with_prefix=True, with_timestamp=True):
        if self._impl is not None:
            self._impl.log(s)
        print(s)

    def record_dict(self, d, prefix=None):
        self._need().record_dict(d, prefix or "")

    def record_tabular(self, key, val):
        self._need().record_tabular(key, val)

    def dump_tabular(self, *args, **kwargs):
        if self._impl is not None and self._impl._row:
            self._impl.dump_tabular()

    def save_itr_params(self, itr, params):
        self._need().save_itr_params(itr, params)


logger = RlkitLogger()


def setup_logger(exp_prefix="default", variant=None, base_log_dir=None, **unused):
    """rlkit.launchers.launcher_util.setup_logger as scripts/train.py:100 calls it: `<base_log_dir>/<prefix>/<prefix>_<timestamp>_0000--s-0/`
    with variant.json, progress.csv, debug.log (SURVEY.md 8b "outputs")."""
    stamp = datetime.datetime.now().strftime("%Y_%m_%d_%H_%M_%S")
    base = base_log_dir or os.path.join(os.getcwd(), "data")
    log_dir = os.path.join(base, exp_prefix, f"{exp_prefix}_{stamp}_0000--s-0")
    logger.setup(log_dir, variant)
    return log_dir


# ----------------------------------------------------------------------------- rlkit.torch.pytorch_util
class _Ptu:
    """`ptu.set_gpu_mode(torch.cuda.is_available())`, `ptu.device` (scripts/train.py:101; util/rlkit_utils.py:162).  The networks and the
    replay ring of this backend always live on the GPU: gpu mode off is an error, not a CPU fallback."""
    _use_gpu = False
    device = None

    def set_gpu_mode(self, mode, gpu_id=0):
        import torch
        from .sac import default_device
        if not mode:
            raise _alg_error("this backend runs the env step and the SAC update on a CUDA device only: set_gpu_mode(False) is not supported")
        self._use_gpu = True
        self.device = torch.device(default_device() if gpu_id == 0 else f"cuda:{gpu_id}")

    def gpu_enabled(self):
        return self._use_gpu


def _alg_error(msg):
    from .backend import RsbError
    return RsbError(msg)


ptu = _Ptu()


# ----------------------------------------------------------------------------- gtimer
class _Stamps:
    def __init__(self):
        self.itrs = OrderedDict()
        self.cum = OrderedDict()


class _Times:
    def __init__(self):
        self.stamps = _Stamps()
        self.total = 0.0


class GTimer:
    """The four gtimer calls of util/rlkit_custom.py:57,138,211-239.  A stamp closes the interval since the previous stamp; inside
    `timed_for(..., save_itrs=True)` every name keeps one value per loop iteration (`stamps.itrs[name]`), repeated stamps of one
    iteration (unique=False) accumulate -- what rlkit's `_get_epoch_timings` reads."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._times = _Times()
        self._last = time.time()
        self._seen = set()

    def stamp(self, name, unique=True, **unused):
        now = time.time()
        dt, self._last = now - self._last, now
        st = self._times.stamps
        st.cum[name] = st.cum.get(name, 0.0) + dt
        if name in self._seen:
            st.itrs[name][-1] += dt
        else:
            st.itrs.setdefault(name, []).append(dt)
            self._seen.add(name)
        self._times.total += dt
        return dt

    def timed_for(self, iterable, save_itrs=True, **unused):
        for x in iterable:
            self._seen = set()
            yield x

    def get_times(self):
        return self._times


gt = GTimer()


def _get_epoch_timings():
    """rlkit.core.rl_algorithm._get_epoch_timings: the `time/* (s)` columns of progress.csv from the timer's last iteration."""
    itrs = gt.get_times().stamps.itrs
    times, epoch_time = OrderedDict(), 0.0
    for key in sorted(itrs):
        t = itrs[key][-1]
        epoch_time += t
        times["time/{} (s)".format(key)] = t
    times["time/epoch (s)"] = epoch_time
    times["time/total (s)"] = gt.get_times().total
    return times


# ----------------------------------------------------------------------------- small rlkit helpers
def list_of_dicts__to__dict_of_lists(lst):
    """rlkit.pythonplusplus: [{a: 1, b: 2}, {a: 3, b: 4}] -> {a: [1, 3], b: [2, 4]} (util/rlkit_custom.py:354)."""
    if len(lst) == 0:
        return {}
    keys = lst[0].keys()
    out = {k: [] for k in keys}
    for d in lst:
        assert set(d.keys()) == set(keys)
        for k in keys:
            out[k].append(d[k])
    return out


class eval_util:
    """rlkit.core.eval_util names used by util/rlkit_custom.py:112,132,327-377."""
    create_stats_ordered_dict = staticmethod(_alg.create_stats_ordered_dict)
    get_generic_path_information = staticmethod(_alg.get_generic_path_information)
    get_average_returns = staticmethod(_alg.get_average_returns)


class BaseRLAlgorithm:
    """rlkit.core.rl_algorithm.BaseRLAlgorithm: imported by util/rlkit_custom.py:8, never instantiated there (the file defines its own base)."""


class ReplayBuffer:
    """rlkit.data_management.replay_buffer.ReplayBuffer: a type annotation in util/rlkit_custom.py:32,174."""


class DataCollector:
    """rlkit.samplers.data_collector.DataCollector / PathCollector: type annotations in util/rlkit_custom.py:30-31,172-173."""


PathCollector = DataCollector


# ----------------------------------------------------------------------------- TD3 names (imported by util/rlkit_utils.py:12-15)
def _td3_unavailable(name):
    class _Stub:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name}: the TD3 branch of util/rlkit_utils.py:107-135 is not on the benchmark path (all committed runs use SAC); "
                                      "the name exists so that the reference's imports resolve")
    _Stub.__name__ = name
    return _Stub


TD3Trainer = _td3_unavailable("TD3Trainer")
TanhMlpPolicy = _td3_unavailable("TanhMlpPolicy")
PolicyWrappedWithExplorationStrategy = _td3_unavailable("PolicyWrappedWithExplorationStrategy")
GaussianStrategy = _td3_unavailable("GaussianStrategy")
