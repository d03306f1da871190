"""`robosuite.wrappers.GymWrapper` as the reference feeds it to rlkit (util/rlkit_utils.py:20,58-62; scripts/rollout.py:138).

`gym` is not a dependency of this package: `Box` below carries what rlkit reads (`low`, `high`, `shape`, `low.size`).
"""
from __future__ import annotations

import numpy as np

from .environments import BatchedEnv, RobosuiteEnv


class Box:
    """Minimal stand-in for gym.spaces.Box (attributes rlkit's NormalizedBoxEnv / EnvReplayBuffer / SACTrainer read)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box{self.shape}"


class GymWrapper:
    """Flattens the observation dict into one vector.

    keys=None -> robosuite v1.0 behaviour: iterate the observation dict in insertion order and keep the robot-state
    blocks followed by `object-state` (the order the committed 2020 policies were trained on).
    Explicit keys -> concatenate in the given order (robosuite >= 1.2 behaviour, what scripts/rollout.py:133-138 passes);
    both `robot{i}_robot-state` and `robot{i}_proprio-state` are accepted.
    """

    def __init__(self, env, keys=None):
        self.env = env
        self.batched = isinstance(env, BatchedEnv)
        names = list(env.obs_slices.keys())
        if keys is None:
            self.keys = names
            self._order_is_native = True
        else:
            keys = [k.replace("proprio-state", "robot-state") for k in keys]
            for k in keys:
                if k not in env.obs_slices:
                    raise KeyError(f"observation key {k!r} not available (have {names})")
            self.keys = keys
            self._order_is_native = keys == names
        self._index = np.concatenate([np.arange(env.obs_dim)[env.obs_slices[k]] for k in self.keys])
        self.obs_dim = int(self._index.size)
        high = np.inf * np.ones(self.obs_dim)
        self.observation_space = Box(-high, high)
        lo, hi = env.action_spec
        self.action_space = Box(lo, hi)
        self._torch_index = None

    def __getattr__(self, name):                    # attribute pass-through (camera_names, robots, _get_observations, ...)
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def seed(self, seed=None):
        if seed is not None:
            np.random.seed(seed)

    def _flatten_obs(self, obs_dict):
        return np.concatenate([np.asarray(obs_dict[k]).ravel() for k in self.keys])

    def _select(self, obs):
        if self._order_is_native:
            return obs
        import torch
        if self._torch_index is None:
            self._torch_index = torch.as_tensor(self._index, device=obs.device)
        return obs.index_select(1, self._torch_index)

    def reset(self, *a, **k):
        if self.batched:
            return self._select(self.env.reset(*a, **k))
        return self._flatten_obs(self.env.reset())

    def step(self, action):
        if self.batched:
            obs, r, d, info = self.env.step(action)
            return self._select(obs), r, d, info
        ob, r, d, info = self.env.step(action)
        return self._flatten_obs(ob), r, d, info
