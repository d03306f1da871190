"""`suite.make(...)` and the env objects it returns -- the robosuite API surface the reference uses
(util/rlkit_utils.py:49-56,189-196; scripts/rollout.py:114-130), backed by the batched CUDA library.

`make(..., num_envs=1)` returns a `RobosuiteEnv` with robosuite's single-env protocol (OrderedDict observations, float
reward, bool done, ValueError when stepped after done).  `make(..., num_envs=N>1)` returns a `BatchedEnv` whose
reset/step take and return torch tensors on the GPU; both run the same kernels.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Sequence

import numpy as np

from .backend import BatchSim
from .controllers import load_controller_config, validate
from .model.tasks import TASK_IDS, build_task

ROBOT_STATE_DIM = 32        # sin q, cos q, qd (3 x 7), eef_pos (3), eef_quat (4), and 2 x 2 gripper joint values with the two-finger grippers


def _obs_slices(task):
    """name -> slice into the flat observation row (robosuite v1.0 order: robot blocks, then object-state).  A robot without a gripper (TwoArmPegInHole) has a
    28-wide block: the width follows the robot's gripper dofs, as in rsb_dev.h obs_element."""
    sl, o = OrderedDict(), 0
    for i in range(task["nrobot"]):
        w = 28 + 2 * int(task["robot"][i]["grip_ndof"])
        sl[f"robot{i}_robot-state"] = slice(o, o + w)
        o += w
    sl["object-state"] = slice(o, task["obs_dim"])
    return sl


class _EnvBase:
    def __init__(self, env_name, robots, controller_configs=None, horizon=500, control_freq=20, reward_scale=1.0,
                 reward_shaping=False, hard_reset=True, ignore_done=False, has_renderer=False,
                 has_offscreen_renderer=False, use_object_obs=True, use_camera_obs=False,
                 env_configuration="single-arm-opposed", num_envs=1, device="cuda:0", seed=0, env_id_base=0,
                 ncon_max=0, nefc_max=0, gripper_types="default", initialization_noise="default", solver="fp32", **unsupported):
        if has_renderer or has_offscreen_renderer or use_camera_obs:
            raise NotImplementedError("rendering / camera observations are outside the batched hot path")
        if not use_object_obs:
            raise NotImplementedError("use_object_obs=False is not supported")
        from .model.tasks import GRIPPERS
        if gripper_types not in ("default", GRIPPERS.get(env_name, "default")):
            raise NotImplementedError(f"gripper_types={gripper_types!r} is not supported for {env_name} (the task's own default gripper only)")
        for k, v in unsupported.items():
            if k in ("render_camera", "render_collision_mesh", "render_visual_mesh", "camera_names", "camera_heights",
                     "camera_widths", "camera_depths", "placement_initializer", "use_indicator_object", "prehensile") and not v:
                continue
            if k in ("camera_names", "camera_heights", "camera_widths", "render_camera"):
                continue                      # accepted and ignored: only meaningful with a renderer
            raise NotImplementedError(f"suite.make option {k}={v!r} is not supported on the batched path")
        if env_name not in TASK_IDS:
            raise NotImplementedError(f"environment {env_name!r} is not implemented (have {sorted(TASK_IDS)})")
        if controller_configs is None:
            controller_configs = load_controller_config(default_controller="OSC_POSE")
        validate(controller_configs)
        self.env_name, self.robot_names = env_name, [robots] if isinstance(robots, str) else list(robots)
        self.horizon, self.control_freq, self.reward_scale = int(horizon), control_freq, reward_scale
        self.ignore_done, self.hard_reset, self.reward_shaping = bool(ignore_done), bool(hard_reset), bool(reward_shaping)
        self.use_object_obs, self.use_camera_obs, self.camera_names = True, False, []
        self.has_renderer = self.has_offscreen_renderer = False
        self.controller_configs = controller_configs
        self.model, self.task = build_task(env_name, self.robot_names, controller_configs, horizon=horizon,
                                           control_freq=control_freq, reward_scale=1.0 if reward_scale is None else reward_scale,
                                           reward_shaping=reward_shaping, ignore_done=ignore_done,
                                           env_configuration=env_configuration, solver=solver)
        self.sim = BatchSim(self.model, self.task, num_envs, device=device, seed=seed, env_id_base=env_id_base,
                            ncon_max=ncon_max, nefc_max=nefc_max)
        self.num_envs = int(num_envs)
        self.robots = [_RobotHandle(n, r) for n, r in zip(self.robot_names, self.task["robot"])]
        self.obs_slices = _obs_slices(self.task)
        self.action_dim, self.obs_dim = self.task["act_dim"], self.task["obs_dim"]
        self.control_timestep, self.model_timestep = 1.0 / control_freq, float(self.model.timestep)

    @property
    def action_spec(self):
        return -np.ones(self.action_dim), np.ones(self.action_dim)

    def close(self):
        self.sim.close()


class _RobotHandle:
    """What the reference reads off `env.robots[i]` (scripts/rollout.py:134 uses only len(env.robots))."""

    def __init__(self, name, desc):
        self.name = name
        self.action_dim = desc["control_dim"] + desc["grip_action_dim"]
        self.dof = self.action_dim
        self.controller_type = desc["ctrl_type"]


class RobosuiteEnv(_EnvBase):
    """Single-environment protocol of robosuite's MujocoEnv (host numpy in/out through the C-ABI's host entry points)."""

    def __init__(self, *a, **k):
        k["num_envs"] = 1
        super().__init__(*a, **k)
        self.done, self.timestep = False, 0

    def _dict(self, row):
        d = OrderedDict()
        row = row.astype(np.float64)
        for name, sl in self.obs_slices.items():
            d[name] = row[sl].copy()
        for i in range(len(self.robots)):            # robosuite >= 1.2 key name for the same block
            d[f"robot{i}_proprio-state"] = d[f"robot{i}_robot-state"]
        return d

    def reset(self):
        obs = self.sim.reset_host()
        self.done, self.timestep = False, 0
        return self._dict(obs[0])

    def step(self, action):
        if self.done:
            raise ValueError("executing action in terminated episode")
        action = np.asarray(action, dtype=np.float64).ravel()
        assert len(action) == self.action_dim, f"environment got invalid action dimension -- expected {self.action_dim}, got {len(action)}"
        obs, rew, done = self.sim.step_host(action[None])
        self.timestep += 1
        self.done = bool(done[0] == 1)
        return self._dict(obs[0]), float(rew[0]), self.done, {}

    def _get_observation(self):
        raise NotImplementedError("call reset()/step(): observations are produced by the step kernel")

    _get_observations = _get_observation

    def render(self):
        raise NotImplementedError("rendering is outside the batched hot path")


class BatchedEnv(_EnvBase):
    """N independent environments; torch tensors on the GPU in and out.  Episodes that hit the horizon are reset by
    `reset(mask)`; with `auto_reset=True`, step() does it and returns the terminal observation separately."""

    def __init__(self, *a, auto_reset=False, **k):
        super().__init__(*a, **k)
        self.auto_reset = auto_reset
        self._steps = None

    def reset(self, mask=None, obs=None):
        import torch
        out = self.sim.reset(mask=mask, obs=obs)
        if self._steps is None or mask is None:
            self._steps = torch.zeros(self.num_envs, dtype=torch.int32, device=self.sim.device)
        else:
            self._steps = torch.where(mask.to(self._steps.device).bool(), torch.zeros_like(self._steps), self._steps)
        return out

    def step(self, actions, obs=None, reward=None, done=None):
        """-> (obs [N, obs_dim], reward [N], done [N] uint8, info).  With ignore_done the `done` output stays 0 and
        `info['timeout']` marks envs whose step counter reached the horizon (what rlkit's collector truncates on)."""
        obs, reward, done = self.sim.step(actions, obs, reward, done)
        self._steps += 1
        timeout = self._steps >= self.horizon
        info = {"timeout": timeout}
        return obs, reward, done, info

    def random_actions(self, step, out=None):
        return self.sim.random_actions(step, out)


def make(env_name, robots, num_envs=1, **kwargs):
    """robosuite.make: single env by default; `num_envs=N` (N > 1) returns the batched GPU environment."""
    batched = kwargs.pop("batched", False)
    if num_envs == 1 and not batched:
        return RobosuiteEnv(env_name, robots, **kwargs)
    return BatchedEnv(env_name, robots, num_envs=num_envs, **kwargs)
