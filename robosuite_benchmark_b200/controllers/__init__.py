"""Controller configuration surface of robosuite that the reference imports:
`from robosuite.controllers import load_controller_config, ALL_CONTROLLERS`
(reference util/rlkit_utils.py:22,42-47,181-186; scripts/rollout.py:5,107-110).

The default dictionaries restate robosuite v1.0.x `controllers/config/*.json`.
Every gain stays data: the dict is compiled into `rsb_robot` (include/rsb_model.h) and the CUDA
kernels read it from there, so a maintainer can pass the JSON of another robosuite version.
"""
from __future__ import annotations

import copy
import json

ALL_CONTROLLERS = {"JOINT_VELOCITY", "JOINT_TORQUE", "JOINT_POSITION", "OSC_POSITION", "OSC_POSE", "IK_POSE"}

_OSC_POSE = {
    "type": "OSC_POSE", "input_max": 1, "input_min": -1,
    "output_max": [0.05, 0.05, 0.05, 0.5, 0.5, 0.5], "output_min": [-0.05, -0.05, -0.05, -0.5, -0.5, -0.5],
    "kp": 150, "damping_ratio": 1, "impedance_mode": "fixed", "kp_limits": [0, 300], "damping_ratio_limits": [0, 10],
    "position_limits": None, "orientation_limits": None, "uncouple_pos_ori": True, "control_delta": True,
    "interpolation": None, "ramp_ratio": 0.2,
    # how a rotation action sets the goal orientation (include/rsb_model.h RSB_ORI_DELTA_*): "euler_transpose" = euler2mat(d)^T R_ee, the convention the
    # reference's committed 2020 policies were trained under (evidence: DESIGN.md 2); "axis_angle" = robosuite >= 1.1
    "orientation_delta": "euler_transpose",
}
_OSC_POSITION = dict(_OSC_POSE, type="OSC_POSITION", output_max=[0.05, 0.05, 0.05], output_min=[-0.05, -0.05, -0.05])
_JOINT_VELOCITY = {
    "type": "JOINT_VELOCITY", "input_max": 1, "input_min": -1, "output_max": 0.5, "output_min": -0.5,
    # robosuite v1.0's law: torque = kv (goal_vel - qvel) + gravity/Coriolis compensation, kv = 4.  This is the law the reference's committed 2020 JOINT_VELOCITY
    # policies were trained under: rolled out here they score 225 vs 226 logged (Lift-Panda, 5 seeds) and 21.4 vs 21.4 (Stack-Panda) with it, 23 and 7.8 with the
    # PID law of robosuite >= 1.1 (profiles/r2_policy_transfer_jv.txt, COMPAT.md).  The later law stays available: give "kp" INSTEAD of "kv" (kp scaled by the
    # actuator range unless "kp_scale_by_actuator_range" is false, ki = "ki_ratio" kp = 0.005 kp, kd = "kd_ratio" kp = 0.001 kp).
    "kv": 4.0, "velocity_limits": [-1, 1], "interpolation": None, "ramp_ratio": 0.2,
}
_JOINT_TORQUE = {
    "type": "JOINT_TORQUE", "input_max": 1, "input_min": -1, "output_max": 0.1, "output_min": -0.1,
    "torque_limits": None, "interpolation": None, "ramp_ratio": 0.2,
}
_JOINT_POSITION = {
    # robosuite v1.0 controllers/config/joint_position.json: goal_qpos = q + action * 0.05 at every policy step, torque = M (kp (goal - q) - kd qd) + compensation
    "type": "JOINT_POSITION", "input_max": 1, "input_min": -1, "output_max": 0.05, "output_min": -0.05, "kp": 50, "damping_ratio": 1,
    "impedance_mode": "fixed", "kp_limits": [0, 300], "damping_ratio_limits": [0, 10], "qpos_limits": None, "interpolation": None, "ramp_ratio": 0.2,
}
_DEFAULTS = {"OSC_POSE": _OSC_POSE, "OSC_POSITION": _OSC_POSITION, "JOINT_VELOCITY": _JOINT_VELOCITY,
             "JOINT_TORQUE": _JOINT_TORQUE, "JOINT_POSITION": _JOINT_POSITION}

#: controllers with a batched CUDA implementation (the others are named by robosuite but not on this path)
SUPPORTED_CONTROLLERS = tuple(_DEFAULTS)


def load_controller_config(custom_fpath=None, default_controller=None):
    """robosuite.controllers.load_controller_config: a default by name, or a JSON file."""
    if default_controller is not None:
        if default_controller not in ALL_CONTROLLERS:
            raise AssertionError(f"Error: Unknown default controller specified. Requested {default_controller}, "
                                 f"available controllers: {sorted(ALL_CONTROLLERS)}")
        if default_controller not in _DEFAULTS:
            raise NotImplementedError(f"controller {default_controller} is outside the batched hot path "
                                      f"(supported: {SUPPORTED_CONTROLLERS})")
        return copy.deepcopy(_DEFAULTS[default_controller])
    if custom_fpath is None:
        raise AssertionError("Error: Either custom_fpath or default_controller must be specified!")
    try:
        with open(custom_fpath) as f:
            cfg = json.load(f)
    except FileNotFoundError:
        raise FileNotFoundError(f"Error opening controller filepath at: {custom_fpath}. Please check filepath and try again.")
    validate(cfg)
    return cfg


def validate(cfg: dict) -> None:
    """Reject options whose semantics the kernels do not implement (loudly, never silently ignored)."""
    if cfg.get("type") not in _DEFAULTS:
        raise NotImplementedError(f"controller type {cfg.get('type')!r} not supported (have {SUPPORTED_CONTROLLERS})")
    if cfg.get("interpolation") not in (None, "null"):
        raise NotImplementedError("controller interpolation is not supported on the batched path")
    if cfg.get("impedance_mode", "fixed") != "fixed":
        raise NotImplementedError("only impedance_mode='fixed' is supported")
    for k in ("position_limits", "orientation_limits", "qpos_limits"):
        if cfg.get(k) is not None:
            raise NotImplementedError(f"{k} is not supported on the batched path")
    if cfg.get("control_delta", True) is not True:
        raise NotImplementedError("absolute OSC goals (control_delta=false) are not supported")
