"""Task scenes: MJCF composition + the task descriptor (include/rsb_model.h `rsb_task`).

Mirrors what robosuite's env constructors do at `suite.make` time (reference call site
util/rlkit_utils.py:49-56; behaviour per SURVEY.md A.1, A.5, A.6): arena + robot(s) + gripper(s)
+ task objects are merged into one MJCF, compiled, and the ids the env logic needs (arm dofs,
actuators, eef site, finger geoms, object bodies) are resolved by name.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from . import assets as A
from .mjcf import Model, compile_mjcf

TASK_IDS = {"Lift": 0, "Door": 1, "Stack": 2, "TwoArmLift": 3, "PickPlaceMilk": 4, "PickPlaceBread": 4, "PickPlaceCereal": 4, "PickPlaceCan": 4,
            "TwoArmPegInHole": 5, "NutAssemblyRound": 6, "TwoArmHandoff": 7}
CTRL_IDS = {"OSC_POSE": 0, "OSC_POSITION": 1, "JOINT_VELOCITY": 2, "JOINT_TORQUE": 3, "JOINT_POSITION": 4}

OBS_DIMS = {"Lift": 42, "Door": 46, "Stack": 55, "TwoArmLift": 89, "PickPlaceMilk": 46, "PickPlaceBread": 46, "PickPlaceCereal": 46, "PickPlaceCan": 46,
            "TwoArmPegInHole": 73, "NutAssemblyRound": 46, "TwoArmHandoff": 86}
#: default bounds of the per-env contact list / constraint-row list (shared-memory sizing; rsb_create ncon_max / nefc_max).
#: Maxima seen over 300 random-action control steps x 2048 envs (tools/limits_stats.py): Stack 16 contacts / 54 rows, TwoArmLift 9 / 31,
#: Door 4 / 18, Lift 9 / 29; TwoArmLift (24, 80) keeps 2.5x headroom and lets 14 envs share an SM (4096 envs = 2 waves instead of 3).
#: Round 2 (481 steps, limits 48 / 160, same tool): Lift-Sawyer reaches 20 contacts / 71 rows (the Rethink gripper's box fingers lie flat on
#: the table) -- it gets its own entry; Lift-Panda 8 / 26, Stack-Sawyer 16 / 54, TwoArmLift 11 / 37, Door 4 / 17.  Truncation beyond these limits
#: is COUNTED (RSB_INFO_NCON_OVERFLOW / NEFC_OVERFLOW): tests/test_gpu_collector.py asserts 0 events over a full random-action episode of 4096
#: envs per family.  Lift-Panda: (18, 62) costs the same shared memory as (16, 64) (28 envs per SM, 4096 envs = one wave) and has no event over a
#: random-action episode where (16, 64) had one (17 contacts; tools/limits_sweep.py, profiles/r2_limits_sweep_lift.txt).  A TRAINED policy grasps: the
#: committed Lift-Panda policy exceeds these limits in 1.4e-4 of its env-steps (pads + table + hand on the cube) --
#: `suite.make(..., ncon_max=24, nefc_max=80)` removes that at the price of a second wave at 4096 envs per GPU.
LIMITS = {"Lift": (18, 62), ("Lift", "Sawyer"): (24, 80), "Door": (16, 64), "Stack": (24, 96), "TwoArmLift": (24, 80),
          "NutAssemblyRound": (56, 184),       # nine boxes resting on the table: 36 contacts / 110 rows before the gripper touches anything; random actions, 256 envs x 500 steps: 45 / 141
          "TwoArmHandoff": (40, 128),          # random actions on the CPU oracle, 256 envs x 500 steps: 27 contacts / 91 rows (Panda fingers and the hammer's five boxes on the table)
          "TwoArmPegInHole": (8, 40),          # committed policies and random actions on the CPU oracle: at most 3 contacts (peg on the rim of the hole) / 11 rows
          # PickPlace on the CPU oracle.  Committed policies (16 episodes each): Panda reaches 13 contacts / 49 rows while carrying the object, Sawyer 18 / 64.
          # Random actions (768 envs x 500 steps): Panda 14 / 46, Sawyer 21 / 74 (the Rethink gripper's box fingers flat on the bin floor against a wall).
          **{"PickPlace" + k: (24, 80) for k in ("Milk", "Bread", "Cereal", "Can")}, **{("PickPlace" + k, "Sawyer"): (32, 104) for k in ("Milk", "Bread", "Cereal", "Can")}}


def limits_for(env_name, robots):
    return LIMITS.get((env_name, robots[0]), LIMITS[env_name])


def _robot_desc(m: Model, pf: str, robot: str, cc: dict, gripper="default") -> dict:
    R = A.ROBOTS[robot]
    jn = [m.id("joint", f"{pf}joint{i + 1}") for i in range(7)]
    ng = 2 if gripper == "default" else 0                  # gripper_types=None (robosuite's NullGripper): no finger joints, no gripper action
    fj = [m.id("joint", f"{pf}finger_joint{i + 1}") for i in range(ng)]
    act = [m.id("actuator", f"{pf}torq_j{i + 1}") for i in range(7)]
    gact = [m.id("actuator", f"{pf}gripper_finger_joint{i + 1}") for i in range(ng)]
    ctype = cc["type"]
    ndim = {"OSC_POSE": 6, "OSC_POSITION": 3, "JOINT_VELOCITY": 7, "JOINT_TORQUE": 7, "JOINT_POSITION": 7}[ctype]

    def vec(x, n, fill=0.0):
        v = np.full(7, fill, float)
        x = np.atleast_1d(np.asarray(x, float))
        v[:n] = x if x.size == n else np.full(n, x[0])
        return v

    tl_lo = np.array([float(m.act_ctrlrange[a, 0]) for a in act])
    tl_hi = np.array([float(m.act_ctrlrange[a, 1]) for a in act])
    ki = np.zeros(7)
    if ctype.startswith("OSC"):
        kp = vec(cc.get("kp", 150.0), 6)
        damping = vec(cc.get("damping_ratio", cc.get("damping", 1.0)), 6)
        kd = 2.0 * np.sqrt(kp) * damping
    elif ctype == "JOINT_POSITION":                            # per-joint kp, kd = 2 sqrt(kp) damping_ratio (robosuite JointPositionController)
        kp = vec(cc.get("kp", 50.0), 7)
        kd = 2.0 * np.sqrt(kp) * vec(cc.get("damping_ratio", 1.0), 7)
    elif ctype == "JOINT_VELOCITY" and "kv" in cc:             # robosuite v1.0: pure proportional law with gain kv (controllers/__init__.py)
        if "kp" in cc:
            raise ValueError('JOINT_VELOCITY config holds both "kv" (v1.0 P law) and "kp" (>= 1.1 PID law): give one')
        kp = vec(cc["kv"], 7)
        ki, kd = np.zeros(7), np.zeros(7)
    elif ctype == "JOINT_VELOCITY":
        kp_in = cc.get("kp", 3.0)
        kp = vec(kp_in, 7)
        if np.isscalar(kp_in) and cc.get("kp_scale_by_actuator_range", True):
            kp = float(kp_in) * (tl_hi - tl_lo)        # robosuite: kp * (high - low) of the actuator range
        ki = kp * float(cc.get("ki_ratio", 0.005))
        kd = kp * float(cc.get("kd_ratio", 0.001))
    else:
        kp, kd = np.zeros(7), np.zeros(7)
    vl = cc.get("velocity_limits")
    d = dict(
        arm_qposadr=[int(m.jnt_qposadr[j]) for j in jn], arm_dofadr=[int(m.jnt_dofadr[j]) for j in jn],
        arm_act=act, grip_ndof=ng, grip_qposadr=[int(m.jnt_qposadr[j]) for j in fj],
        grip_dofadr=[int(m.jnt_dofadr[j]) for j in fj], grip_act=gact, grip_action_dim=1 if ng else 0,
        grip_sign=list(R["grip_sign"])[:ng], grip_speed=0.01, grip_init_qpos=list(R["grip_init"])[:ng],
        eef_site=m.id("site", f"{pf}grip_site"), eef_body=m.id("body", f"{pf}right_hand"),
        init_qpos=list(R["init_qpos"]),
        left_finger_geoms=[m.id("geom", f"{pf}finger1_col"), m.id("geom", f"{pf}finger1_pad")] if ng else [],
        right_finger_geoms=[m.id("geom", f"{pf}finger2_col"), m.id("geom", f"{pf}finger2_pad")] if ng else [],
        ctrl_type=CTRL_IDS[ctype], control_dim=ndim,
        input_max=vec(cc.get("input_max", 1.0), ndim), input_min=vec(cc.get("input_min", -1.0), ndim),
        output_max=vec(cc.get("output_max", 1.0), ndim), output_min=vec(cc.get("output_min", -1.0), ndim),
        kp=kp, kd=kd, ki=ki, nullspace_kp=10.0, uncouple_pos_ori=int(bool(cc.get("uncouple_pos_ori", True))),
        ori_delta_mode={"euler_transpose": 0, "axis_angle": 1}[cc.get("orientation_delta", "euler_transpose")],
        torque_limit_lo=tl_lo, torque_limit_hi=tl_hi,
        velocity_limit_lo=vec(vl[0], 7) if vl is not None else np.zeros(7),
        velocity_limit_hi=vec(vl[1], 7) if vl is not None else np.zeros(7),
        has_velocity_limits=int(vl is not None),
    )
    return d


def build_task(env_name: str, robots: Sequence[str], controller_config: dict, horizon: int = 500,
               control_freq: float = 20, reward_scale: float = 1.0, reward_shaping: bool = True,
               ignore_done: bool = False, env_configuration: str = "single-arm-opposed", solver="fp32"):
    """-> (Model, task dict).  `controller_config` is the dict from load_controller_config.

    `solver`: the constraint solver runs with the compiled model's `<option iterations tolerance ls_iterations ls_tolerance>` -- the
    kernels read those four fields of the model, nothing is hard-coded on the device.  "model" keeps what the MJCF says (MuJoCo's
    defaults: 100 Newton iterations, tolerance 1e-8, 50 line-search steps); "fp32" (default) is the documented override for fp32
    arithmetic, where improvements below ~1e-6 of the scaled cost are round-off (SOLVER_FP32 below); a dict overrides field by field."""
    if isinstance(robots, str):
        robots = [robots]
    if env_name not in TASK_IDS:
        raise NotImplementedError(f"environment {env_name!r} is not on the batched hot path "
                                  f"(supported: {sorted(TASK_IDS)})")
    for r in robots:
        if r not in A.ROBOTS:
            raise NotImplementedError(f"robot {r!r} not supported (have {sorted(A.ROBOTS)})")
    builder = _BUILDERS[env_name]
    xml, objs = builder(robots, env_configuration)
    m = compile_mjcf(xml)
    apply_solver_option(m, solver)
    substeps = int((1.0 / control_freq) / m.timestep)
    n_rob = len(robots)
    rdesc = [_robot_desc(m, f"robot{i}_", r, controller_config, gripper=GRIPPERS.get(env_name, "default")) for i, r in enumerate(robots)]
    act_dim = sum(r["control_dim"] + r["grip_action_dim"] for r in rdesc)
    task = dict(task_id=TASK_IDS[env_name], nrobot=n_rob, robot=rdesc, horizon=int(horizon), substeps=substeps,
                ignore_done=int(bool(ignore_done)), reward_shaping=int(bool(reward_shaping)),
                reward_scale=float(reward_scale), init_noise=0.02, table_height=A.TABLE_HEIGHT,
                obs_dim=OBS_DIMS[env_name] if n_rob == (2 if env_name.startswith("TwoArm") else 1) else None,
                act_dim=act_dim, env_name=env_name, robots=list(robots), xml=xml, ncon_max=limits_for(env_name, robots)[0], nefc_max=limits_for(env_name, robots)[1])
    task.update(objs(m))
    return m, task


#: tasks whose robots carry no gripper (robosuite: gripper_types=None)
GRIPPERS = {"TwoArmPegInHole": None}

SOLVER_FP32 = dict(iterations=12, tolerance=1e-6, ls_iterations=24, ls_tolerance=0.01)


def apply_solver_option(m: Model, solver):
    """Write the solver override into the model's option block (see build_task)."""
    if solver in (None, "model"):
        return m
    over = SOLVER_FP32 if solver == "fp32" else dict(solver)
    for k, v in over.items():
        if k not in ("iterations", "tolerance", "ls_iterations", "ls_tolerance"):
            raise ValueError(f"unknown solver option {k!r}")
        m.opt[k] = type(SOLVER_FP32[k])(v)
    return m


def empty_task():
    """Task descriptor with no robots/objects: lets the physics stages run on an arbitrary compiled model."""
    t = dict(task_id=-1, nrobot=0, robot=[], horizon=1 << 30, substeps=1, ignore_done=1, reward_shaping=0,
             reward_scale=1.0, init_noise=0.0, table_height=0.0, obs_dim=0, act_dim=0)
    t.update(_empty_objs())
    return t


def _empty_objs():
    return dict(obj_body=[-1] * 4, obj_geom=[-1] * 4, obj_site=[-1] * 4, obj_qposadr=[-1] * 4, obj_dofadr=[-1] * 4,
                obj_half=np.zeros((4, 3)), place_x=np.zeros((4, 2)), place_y=np.zeros((4, 2)),
                place_yaw=np.zeros((4, 2)), place_z=np.zeros(4), place_ref=np.zeros(3), place_body=[-1] * 4, task_par=[0.0] * 8)


def _single_arm_world(robot: str):
    R = A.ROBOTS[robot]
    base = (-0.16 - A.TABLE_FULL[0] / 2, 0.0, A.ROBOT_BASE_Z)
    return R["body"]("robot0_", base), R["act"]("robot0_")


def _lift(robots, env_configuration):
    """Lift: one cube on the table."""
    assert len(robots) == 1, "Lift takes one robot"
    body, act = _single_arm_world(robots[0])
    half = 0.021
    world = A.table_arena() + body + A.box_object("cube", [half] * 3, [0, 0, A.TABLE_HEIGHT + half])
    xml = A.scene(world, act)

    def objs(m: Model):
        o = _empty_objs()
        j = m.id("joint", "cube_joint")
        o["obj_body"][0] = m.id("body", "cube")
        o["obj_geom"][0] = m.id("geom", "cube_g0")
        o["obj_qposadr"][0], o["obj_dofadr"][0] = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
        o["obj_half"][0] = half
        o["place_x"][0], o["place_y"][0] = [-0.03, 0.03], [-0.03, 0.03]
        o["place_yaw"][0] = [0.0, 2 * np.pi]
        o["place_z"][0] = A.TABLE_HEIGHT + half + 0.01  # robosuite drops objects from 1 cm
        o["place_ref"] = np.array([0.0, 0.0, A.TABLE_HEIGHT])
        return o

    return xml, objs


def _stack(robots, env_configuration):
    """Stack: cubeA (half 0.02) to be put on cubeB (half 0.025); both placed uniformly in +-0.08 with overlap rejection."""
    assert len(robots) == 1, "Stack takes one robot"
    body, act = _single_arm_world(robots[0])
    ha, hb = 0.02, 0.025
    world = (A.table_arena() + body + A.box_object("cubeA", [ha] * 3, [-0.04, -0.04, A.TABLE_HEIGHT + ha])
             + A.box_object("cubeB", [hb] * 3, [0.04, 0.04, A.TABLE_HEIGHT + hb]))
    xml = A.scene(world, act)

    def objs(m: Model):
        o = _empty_objs()
        for k, (nm, hh) in enumerate((("cubeA", ha), ("cubeB", hb))):
            j = m.id("joint", nm + "_joint")
            o["obj_body"][k], o["obj_geom"][k] = m.id("body", nm), m.id("geom", nm + "_g0")
            o["obj_qposadr"][k], o["obj_dofadr"][k] = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
            o["obj_half"][k] = hh
            o["place_x"][k], o["place_y"][k], o["place_yaw"][k] = [-0.08, 0.08], [-0.08, 0.08], [0.0, 2 * np.pi]
            o["place_z"][k] = A.TABLE_HEIGHT + hh + 0.01
        o["place_ref"] = np.array([0.0, 0.0, A.TABLE_HEIGHT])
        return o

    return xml, objs


def _door(robots, env_configuration):
    """Door: hinged door with a latch handle standing on the table; the door root is re-placed at every reset."""
    assert len(robots) == 1, "Door takes one robot"
    body, act = _single_arm_world(robots[0])
    recalled = A.DOOR_STYLE == "robosuite_recalled"
    if recalled:
        world = A.table_arena(full=A.DOOR_TABLE_FULL, offset=A.DOOR_TABLE_OFFSET) + body + A.door_object_recalled()
    else:
        world = A.table_arena() + body + A.door_object()
    xml = A.scene(world, act)

    def objs(m: Model):
        o = _empty_objs()
        jh, jl = m.id("joint", "door_hinge"), m.id("joint", "latch_joint")
        o["obj_body"][0], o["obj_body"][1], o["obj_body"][2] = m.id("body", "door"), m.id("body", "latch"), m.id("body", "door_root")
        o["obj_geom"][0] = m.id("geom", "latch_handle")
        o["obj_site"][0] = m.id("site", "door_handle")
        o["obj_qposadr"][0], o["obj_qposadr"][1] = int(m.jnt_qposadr[jh]), int(m.jnt_qposadr[jl])
        o["obj_dofadr"][0], o["obj_dofadr"][1] = int(m.jnt_dofadr[jh]), int(m.jnt_dofadr[jl])
        # placement slot 2 = the fixed root body: x in [0.07, 0.09], y in [-0.01, 0.01], yaw in [-pi/2 - 0.25, -pi/2] about place_ref
        o["place_body"][2] = m.id("body", "door_root")
        o["place_x"][2], o["place_y"][2], o["place_yaw"][2] = [0.07, 0.09], [-0.01, 0.01], [-np.pi / 2 - 0.25, -np.pi / 2]
        o["place_z"][2] = A.TABLE_HEIGHT
        o["place_ref"] = np.array([0.04, -0.2, A.TABLE_HEIGHT])
        if recalled:                                       # UniformRandomSampler about the table offset; the object's bottom (0.3 below its origin) on the table top
            o["place_z"][2] = A.TABLE_HEIGHT + 0.3
            o["place_ref"] = np.array([A.DOOR_TABLE_OFFSET[0], A.DOOR_TABLE_OFFSET[1], A.TABLE_HEIGHT])
        return o

    return xml, objs


def _two_arm_lift(robots, env_configuration):
    """TwoArmLift, `single-arm-opposed`: two arms facing each other across the table, a pot with two handles between them."""
    assert len(robots) == 2, "TwoArmLift takes two robots"
    if env_configuration not in ("single-arm-opposed", "default", None):
        raise NotImplementedError(f"env_configuration {env_configuration!r} is not implemented (single-arm-opposed only)")
    bodies, acts = "", ""
    for i, (r, yaw, y) in enumerate(zip(robots, (np.pi / 2, -np.pi / 2), (-A.TWO_ARM_BASE_Y, A.TWO_ARM_BASE_Y))):
        R = A.ROBOTS[r]
        quat = (np.cos(yaw / 2), 0, 0, np.sin(yaw / 2))
        bodies += R["body"](f"robot{i}_", (0.0, y, A.ROBOT_BASE_Z), quat)
        acts += R["act"](f"robot{i}_")
    world = A.table_arena() + bodies + A.pot_with_handles()
    xml = A.scene(world, acts)

    def objs(m: Model):
        o = _empty_objs()
        j = m.id("joint", "pot_joint")
        o["obj_body"][0] = m.id("body", "pot")
        h0, h1 = ("pot_handle1", "pot_handle0") if A.TWO_ARM_POT_YAW_PI else ("pot_handle0", "pot_handle1")
        o["obj_geom"][0], o["obj_geom"][1] = m.id("geom", h0), m.id("geom", h1)
        o["obj_site"][0], o["obj_site"][1] = m.id("site", h0), m.id("site", h1)
        o["obj_qposadr"][0], o["obj_dofadr"][0] = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
        o["obj_half"][0] = [0.07, 0.07, 0.07]
        yaw0 = np.pi if A.TWO_ARM_POT_YAW_PI else 0.0
        o["place_x"][0], o["place_y"][0], o["place_yaw"][0] = [-0.03, 0.03], [-0.03, 0.03], [yaw0 - np.pi / 3, yaw0 + np.pi / 3]
        o["place_z"][0] = A.TABLE_HEIGHT + 0.07 + 0.01
        o["place_ref"] = np.array([0.0, 0.0, A.TABLE_HEIGHT])
        return o

    return xml, objs


def _pick_place(kind):
    """PickPlace in single-object mode (robosuite's PickPlaceMilk / PickPlaceCan / ... = PickPlace(single_object_mode=2, object_type=kind)): the object starts at a
    uniform pose in bin 1 and goes into its quadrant of bin 2.  The other three objects, which robosuite parks at x = 10, are not modelled."""

    def build(robots, env_configuration):
        assert len(robots) == 1, "PickPlace takes one robot"
        R = A.ROBOTS[robots[0]]
        body, act = R["body"]("robot0_", (A.BINS_ROBOT_BASE[0], A.BINS_ROBOT_BASE[1], A.ROBOT_BASE_Z)), R["act"]("robot0_")
        P = A.PICK_OBJECTS[kind]
        half, top = np.array(P["half"]), A.BIN1_POS[2] + A.BIN_FLOOR_HALF[2]
        world = A.bins_arena() + body + A.pick_object(kind, [A.BIN1_POS[0], A.BIN1_POS[1], top + half[2]])
        xml = A.scene(world, act)

        def objs(m: Model):
            o = _empty_objs()
            j = m.id("joint", kind + "_joint")
            o["obj_body"][0], o["obj_geom"][0] = m.id("body", kind), m.id("geom", kind + "_g0")
            o["obj_qposadr"][0], o["obj_dofadr"][0] = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
            o["obj_half"][0] = half
            # robosuite's bin sampler: uniform over the bin minus the object's horizontal radius and a 5 cm border, any yaw, resting on the bin floor
            rad = float(np.hypot(half[0], half[1]))
            bx, by = A.BIN_SIZE[0] / 2 - rad - 0.05, A.BIN_SIZE[1] / 2 - rad - 0.05
            o["place_x"][0], o["place_y"][0], o["place_yaw"][0] = [-bx, bx], [-by, by], [0.0, 2 * np.pi]
            o["place_z"][0] = top + half[2] + 0.01            # dropped from 1 cm like the table tasks (an exact resting pose would leave the contact set to round-off)
            o["place_ref"] = np.array([A.BIN1_POS[0], A.BIN1_POS[1], top])
            # target placement: centre of quadrant bin_id of bin 2 (PickPlace._reset_internal: ids 0 / 2 on the low-x side, ids 0 / 1 on the low-y side)
            b = P["bin_id"]
            tx = A.BIN2_POS[0] - (A.BIN_SIZE[0] / 2 if b in (0, 2) else 0.0) + A.BIN_SIZE[0] / 4
            ty = A.BIN2_POS[1] - (A.BIN_SIZE[1] / 2 if b < 2 else 0.0) + A.BIN_SIZE[1] / 4
            o["task_par"] = [tx, ty, A.BIN2_POS[2], A.BIN_SIZE[0], A.BIN_SIZE[1], 0.25]
            return o

        return xml, objs

    return build


def _nut_assembly_round(robots, env_configuration):
    """NutAssemblyRound = NutAssembly(single_object_mode=2, nut_type="round"): the round nut starts on the table beside the robot and goes over the round peg.
    The square nut, which robosuite parks away from the scene in this mode, is not modelled (its peg is: the arm can run into it)."""
    assert len(robots) == 1, "NutAssembly takes one robot"
    R = A.ROBOTS[robots[0]]
    body, act = R["body"]("robot0_", (-0.16 - A.NUT_TABLE_FULL[0] / 2, 0.0, A.ROBOT_BASE_Z)), R["act"]("robot0_")
    hh = A.ROUND_NUT["half_h"]
    world = A.pegs_arena() + body + A.round_nut("RoundNut", [-0.1125, -0.17, A.NUT_TABLE_Z + hh])
    xml = A.scene(world, act)

    def objs(m: Model):
        o = _empty_objs()
        j = m.id("joint", "RoundNut_joint")
        o["obj_body"][0] = m.id("body", "RoundNut")
        o["obj_geom"][0], o["obj_geom"][1] = m.id("geom", "RoundNut_ring0"), m.id("geom", "RoundNut_handle")
        assert o["obj_geom"][1] - o["obj_geom"][0] == 8                     # contiguous geom ids: the grasp check tests a range
        o["obj_qposadr"][0], o["obj_dofadr"][0] = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
        o["obj_half"][0] = [A.ROUND_NUT["handle_out"], A.ROUND_NUT["ring_in"] + 2 * A.ROUND_NUT["ring_t"], hh]
        P = A.NUT_PLACE["Round"]
        o["place_x"][0], o["place_y"][0], o["place_yaw"][0] = list(P["x"]), list(P["y"]), [0.0, 2 * np.pi]
        o["place_z"][0] = A.NUT_TABLE_Z + hh + 0.02
        o["place_ref"] = np.array([0.0, 0.0, A.NUT_TABLE_Z])
        px, py = A.NUT_PEGS["Round"]
        o["task_par"] = [px, py, A.NUT_TABLE_Z, A.NUT_TABLE_Z - A.NUT_TABLE_FULL[2] / 2 + 0.2]       # lift target: the table BODY's z (its centre) + 0.2
        return o

    return xml, objs


def _two_arm_handoff(robots, env_configuration):
    """TwoArmHandoff, `single-arm-opposed`, prehensile: the hammer lies on a narrow table beside robot 0, which picks it up and hands it to robot 1."""
    assert len(robots) == 2, "TwoArmHandoff takes two robots"
    if env_configuration not in ("single-arm-opposed", "default", None):
        raise NotImplementedError(f"env_configuration {env_configuration!r} is not implemented (single-arm-opposed only)")
    H = A.HANDOFF
    bodies, acts = "", ""
    for i, (r, yaw, y) in enumerate(zip(robots, (np.pi / 2, -np.pi / 2), (-H["base_y"], H["base_y"]))):
        R = A.ROBOTS[r]
        bodies += R["body"](f"robot{i}_", (0.0, y, A.ROBOT_BASE_Z), (np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)))
        acts += R["act"](f"robot{i}_")
    rad = H["handle_radius"]
    world = A.table_arena(full=H["table_full"], offset=H["table_offset"]) + bodies + A.hammer("hammer", [H["table_offset"][0], H["table_offset"][1], A.TABLE_HEIGHT + 2 * rad])
    xml = A.scene(world, acts)

    def objs(m: Model):
        o = _empty_objs()
        j = m.id("joint", "hammer_joint")
        o["obj_body"][0] = m.id("body", "hammer")
        o["obj_geom"][0], o["obj_geom"][1] = m.id("geom", "hammer_handle"), m.id("geom", "hammer_claw")
        assert o["obj_geom"][1] - o["obj_geom"][0] == 4                     # handle first, contiguous ids: the grasp checks test a range / the first geom
        o["obj_qposadr"][0], o["obj_dofadr"][0] = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
        o["obj_half"][0] = [H["handle_length"] / 2, rad, rad]
        o["place_x"][0], o["place_y"][0] = list(H["place_x"]), list(H["place_y"])
        # a quarter turn about the world x axis lays the handle (body z) along world y, i.e. pointing from one robot to the other; the reset code draws the sign
        # (head towards robot 0 or towards robot 1)
        o["place_yaw"][0] = [np.pi / 2 - H["tilt"], np.pi / 2 + H["tilt"]]
        o["place_z"][0] = A.TABLE_HEIGHT + 1.2 * rad * H["head_half_ratio"] + 0.01
        o["place_ref"] = np.array([H["table_offset"][0], H["table_offset"][1], A.TABLE_HEIGHT])
        o["task_par"] = [H["lift_height"], rad]
        return o

    return xml, objs


def _two_arm_peg_in_hole(robots, env_configuration):
    """TwoArmPegInHole, `single-arm-opposed`: two gripper-less arms facing each other over an empty floor; robot 0 holds the peg, robot 1 the plate with the hole
    (both rigidly attached to the hands: no free bodies, nothing to place at reset)."""
    assert len(robots) == 2, "TwoArmPegInHole takes two robots"
    if env_configuration not in ("single-arm-opposed", "default", None):
        raise NotImplementedError(f"env_configuration {env_configuration!r} is not implemented (single-arm-opposed only)")
    bodies, acts = "", ""
    by = A.PEG_IN_HOLE["base_y"]
    for i, (r, yaw, y, payload) in enumerate(zip(robots, (np.pi / 2, -np.pi / 2), (-by, by), (A.peg_payload(), A.hole_payload()))):
        R = A.ROBOTS[r]
        quat = (np.cos(yaw / 2), 0, 0, np.sin(yaw / 2))
        bodies += R["body"](f"robot{i}_", (0.0, y, A.ROBOT_BASE_Z), quat, gripper=None, payload=payload)
        acts += R["act"](f"robot{i}_", gripper=None)
    xml = A.scene(A.empty_arena() + bodies, acts)

    def objs(m: Model):
        o = _empty_objs()
        o["obj_body"][0], o["obj_body"][1] = m.id("body", "hole"), m.id("body", "peg")
        o["obj_geom"][0] = m.id("geom", "peg_g0")
        o["obj_site"][0] = m.id("site", "hole_center")
        o["task_par"] = [A.PEG_IN_HOLE["hole_center"]] + [0.0] * 7
        return o

    return xml, objs


_BUILDERS: Dict[str, callable] = {"Lift": _lift, "Stack": _stack, "Door": _door, "TwoArmLift": _two_arm_lift, "TwoArmPegInHole": _two_arm_peg_in_hole, "NutAssemblyRound": _nut_assembly_round, "TwoArmHandoff": _two_arm_handoff,
                                  **{"PickPlace" + k: _pick_place(k) for k in A.PICK_OBJECTS}}
