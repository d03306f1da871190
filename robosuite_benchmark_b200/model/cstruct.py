"""ctypes mirrors of include/rsb_model.h (`rsb_model`, `rsb_robot`, `rsb_task`).

Field order and types here MUST match the header; tests/test_abi.py checks the struct sizes
against the sizes the compiled libraries report (`rsb_sizeof_model` / `rsb_sizeof_task`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .mjcf import Model

PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int)

_MODEL_FIELDS = [
    ("sizes", ["nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "npair", "nM"], C.c_int),
    ("opt_d", ["timestep"], C.c_double),
    ("gravity", None, C.c_double * 3),
    ("opt_d2", ["impratio", "tolerance", "ls_tolerance", "meaninertia"], C.c_double),
    ("opt_i", ["cone", "iterations", "ls_iterations"], C.c_int),
    ("pi", ["body_parentid", "body_rootid", "body_jntadr", "body_jntnum", "body_dofadr", "body_dofnum"], PI),
    ("pd", ["body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_invweight0"], PD),
    ("pi", ["jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited"], PI),
    ("pd", ["jnt_pos", "jnt_axis", "jnt_range", "jnt_stiffness", "jnt_margin", "jnt_solref", "jnt_solimp"], PD),
    ("pi", ["dof_bodyid", "dof_jntid", "dof_parentid", "dof_Madr"], PI),
    ("pd", ["dof_armature", "dof_damping", "dof_frictionloss", "dof_invweight0", "dof_solref", "dof_solimp",
            "qpos0", "qpos_spring"], PD),
    ("pi", ["geom_type", "geom_bodyid"], PI),
    ("pd", ["geom_size", "geom_pos", "geom_quat", "geom_rbound"], PD),
    ("pi", ["site_bodyid"], PI),
    ("pd", ["site_pos", "site_quat"], PD),
    ("pi", ["pair_geom1", "pair_geom2", "pair_condim"], PI),
    ("pd", ["pair_friction", "pair_solref", "pair_solimp", "pair_margin", "pair_gap"], PD),
    ("pi", ["act_dofid", "act_ctrllimited", "act_forcelimited"], PI),
    ("pd", ["act_gain", "act_bias", "act_ctrlrange", "act_forcerange", "act_gear"], PD),
]


def _model_struct_fields():
    out = []
    for kind, names, typ in _MODEL_FIELDS:
        if names is None:
            out.append((kind, typ))
        else:
            out.extend((n, typ) for n in names)
    return out


class RsbModel(C.Structure):
    _fields_ = _model_struct_fields()


A7 = C.c_double * 7
I7 = C.c_int * 7
I2 = C.c_int * 2
D2 = C.c_double * 2
I4 = C.c_int * 4


class RsbRobot(C.Structure):
    _fields_ = [
        ("arm_qposadr", I7), ("arm_dofadr", I7), ("arm_act", I7),
        ("grip_ndof", C.c_int), ("grip_qposadr", I2), ("grip_dofadr", I2), ("grip_act", I2),
        ("grip_action_dim", C.c_int), ("grip_sign", D2), ("grip_speed", C.c_double), ("grip_init_qpos", D2),
        ("eef_site", C.c_int), ("eef_body", C.c_int), ("init_qpos", A7),
        ("left_finger_geoms", I4), ("n_left_finger_geoms", C.c_int),
        ("right_finger_geoms", I4), ("n_right_finger_geoms", C.c_int),
        ("ctrl_type", C.c_int), ("control_dim", C.c_int),
        ("input_max", A7), ("input_min", A7), ("output_max", A7), ("output_min", A7),
        ("kp", A7), ("kd", A7), ("ki", A7), ("nullspace_kp", C.c_double), ("uncouple_pos_ori", C.c_int), ("ori_delta_mode", C.c_int),
        ("torque_limit_lo", A7), ("torque_limit_hi", A7), ("velocity_limit_lo", A7), ("velocity_limit_hi", A7),
        ("has_velocity_limits", C.c_int),
    ]


class RsbTask(C.Structure):
    _fields_ = [
        ("task_id", C.c_int), ("nrobot", C.c_int), ("robot", RsbRobot * 2),
        ("horizon", C.c_int), ("substeps", C.c_int), ("ignore_done", C.c_int), ("reward_shaping", C.c_int),
        ("reward_scale", C.c_double), ("init_noise", C.c_double), ("table_height", C.c_double),
        ("obs_dim", C.c_int), ("act_dim", C.c_int),
        ("obj_body", I4), ("obj_geom", I4), ("obj_site", I4), ("obj_qposadr", I4), ("obj_dofadr", I4),
        ("obj_half", (C.c_double * 3) * 4),
        ("place_x", D2 * 4), ("place_y", D2 * 4), ("place_yaw", D2 * 4), ("place_z", C.c_double * 4),
        ("place_ref", C.c_double * 3), ("place_body", I4),
        ("task_par", C.c_double * 8),
    ]


def model_to_c(m: Model):
    """-> (RsbModel, keepalive list).  The arrays must outlive every C call that reads the struct."""
    s = RsbModel()
    keep = []
    for kind, names, typ in _MODEL_FIELDS:
        if names is None:
            s.gravity = (C.c_double * 3)(*[float(x) for x in m.gravity])
            continue
        for n in names:
            if typ is C.c_int:
                setattr(s, n, int(getattr(m, n)))
            elif typ is C.c_double:
                setattr(s, n, float(getattr(m, n)))
            elif typ is PI:
                a = np.ascontiguousarray(getattr(m, n), dtype=np.int32)
                keep.append(a)
                setattr(s, n, a.ctypes.data_as(PI))
            else:
                a = np.ascontiguousarray(getattr(m, n), dtype=np.float64)
                keep.append(a)
                setattr(s, n, a.ctypes.data_as(PD))
    return s, keep


def _fill(dst, src):
    src = np.asarray(src)
    if src.ndim == 2:
        for i in range(src.shape[0]):
            _fill(dst[i], src[i])
        return
    for i, v in enumerate(src.tolist()):
        dst[i] = v


def task_to_c(task: dict) -> RsbTask:
    t = RsbTask()
    for k in ("task_id", "nrobot", "horizon", "substeps", "ignore_done", "reward_shaping", "obs_dim", "act_dim"):
        setattr(t, k, int(task[k]))
    for k in ("reward_scale", "init_noise", "table_height"):
        setattr(t, k, float(task[k]))
    for k in ("obj_body", "obj_geom", "obj_site", "obj_qposadr", "obj_dofadr", "obj_half", "place_x", "place_y",
              "place_yaw", "place_z", "place_ref", "place_body"):
        _fill(getattr(t, k), task[k])
    _fill(t.task_par, np.concatenate([np.asarray(task.get("task_par", []), float), np.zeros(8)])[:8])
    for ri, rd in enumerate(task["robot"]):
        r = t.robot[ri]
        for name, typ in RsbRobot._fields_:
            v = rd[name] if name in rd else None
            if name.startswith("n_") and v is None:
                v = len(rd[name[2:]])
            if typ in (C.c_int,):
                setattr(r, name, int(v))
            elif typ is C.c_double:
                setattr(r, name, float(v))
            else:
                arr = getattr(r, name)
                vals = list(np.asarray(v).ravel().tolist())
                fill = -1 if typ in (I4, I7, I2) else 0.0
                vals = vals + [fill] * (len(arr) - len(vals))
                for i in range(len(arr)):
                    arr[i] = vals[i]
    return t
