"""Authored MJCF for the benchmark scenes (robots, grippers, arena, task objects).

robosuite's own XML/mesh assets are not available in the build environment (SURVEY.md §0.3,
A.5), so these are written from public kinematic data (Franka Panda / Rethink Sawyer link
frames, robosuite's table arena and object conventions) with PRIMITIVE collision geoms only.
Everything robosuite would take from its asset files is data here: a maintainer with the real
assets can feed their compiled mjModel through include/rsb_model.h instead.

Deviations from robosuite v1.0 (documented in DESIGN.md): arm links carry no collision geoms
(meshes upstream); gripper hand/fingers/pads are boxes; the Lift cube half-size is the mean
(0.021) of robosuite's U(0.020, 0.022) build-time draw.
"""
from __future__ import annotations

import numpy as np

PI = np.pi

BASE_OPTION = ('<compiler angle="radian"/>'
               '<option timestep="0.002" cone="elliptic" impratio="20" integrator="Euler" solver="Newton" '
               'iterations="100" tolerance="1e-8" gravity="0 0 -9.81"/>')

# collision classes: robot parts hit objects/arena (contype 1) but not each other
ROBOT_COL = 'contype="0" conaffinity="1"'
WORLD_COL = 'contype="1" conaffinity="1"'
NO_COL = 'contype="0" conaffinity="0"'


def _rotz(a, v):
    return (np.cos(a) * v[0] - np.sin(a) * v[1], np.sin(a) * v[0] + np.cos(a) * v[1], v[2])


def _f(v):
    return " ".join(f"{float(x):.10g}" for x in np.atleast_1d(v))


# ----------------------------------------------------------------------------- grippers
PANDA_GRIPPER = dict(finger_inertia=(0.0001, 0.0001, 0.0001), hand_inertia=(0.002, 0.002, 0.002), hand_mass=0.5)


def panda_gripper(pf: str) -> str:
    """Franka hand: two slide fingers (±y), box pads with high friction."""
    pad = f'{ROBOT_COL} condim="4" friction="2 0.05 0.0001" solref="0.01 0.5"'
    return f'''
<body name="{pf}right_hand" pos="0 0 0.107" quat="0.9238795 0 0 -0.3826834">
  <inertial pos="0 0 0.03" mass="{PANDA_GRIPPER["hand_mass"]}" diaginertia="{_f(PANDA_GRIPPER["hand_inertia"])}"/>
  <geom name="{pf}hand_col" type="box" pos="0 0 0.0333" size="0.0315 0.102 0.0333" {ROBOT_COL}/>
  <site name="{pf}grip_site" pos="0 0 0.1025"/>
  <body name="{pf}leftfinger" pos="0 0 0.0584">
    <inertial pos="0 0.01 0.02" mass="0.1" diaginertia="{_f(PANDA_GRIPPER["finger_inertia"])}"/>
    <joint name="{pf}finger_joint1" type="slide" axis="0 1 0" range="0 0.04" damping="100" armature="1.0" frictionloss="1.0"/>
    <geom name="{pf}finger1_col" type="box" pos="0 0.0115 0.027" size="0.0105 0.0075 0.027" {ROBOT_COL}/>
    <geom name="{pf}finger1_pad" type="box" pos="0 0.0025 0.044" size="0.008 0.0015 0.008" {pad}/>
  </body>
  <body name="{pf}rightfinger" pos="0 0 0.0584">
    <inertial pos="0 -0.01 0.02" mass="0.1" diaginertia="{_f(PANDA_GRIPPER["finger_inertia"])}"/>
    <joint name="{pf}finger_joint2" type="slide" axis="0 1 0" range="-0.04 0" damping="100" armature="1.0" frictionloss="1.0"/>
    <geom name="{pf}finger2_col" type="box" pos="0 -0.0115 0.027" size="0.0105 0.0075 0.027" {ROBOT_COL}/>
    <geom name="{pf}finger2_pad" type="box" pos="0 -0.0025 0.044" size="0.008 0.0015 0.008" {pad}/>
  </body>
</body>'''


def panda_gripper_actuators(pf: str) -> str:
    return (f'<position name="{pf}gripper_finger_joint1" joint="{pf}finger_joint1" kp="1000" ctrlrange="0 0.04" forcerange="-20 20"/>'
            f'<position name="{pf}gripper_finger_joint2" joint="{pf}finger_joint2" kp="1000" ctrlrange="-0.04 0" forcerange="-20 20"/>')


#: Rethink finger geometry, in the finger body's frame, for the finger on the +y side (mirrored for the other): (pos, half size) of the finger bar and of the pad.
#: "narrow_tall" restates robosuite's rethink_gripper.xml collision boxes as recalled (finger bar 10 x 13.5 x 75 mm without friction, pad 7 x 8 x 33 mm under the
#: fingertip; the fully open gripper clears 79 mm); "round1" is the stand-in of round 1 (16 x 3 x 16 mm pads, 63 mm clearance).  The committed Sawyer policies do
#: not tell them apart (Lift 139 vs 144, Stack 23.5 vs 25.2 of 348 / 28 logged: profiles/r2_policy_transfer_sawyer.txt), so the shipped default stays "round1".
RETHINK_FINGERS = {
    "narrow_tall": dict(bar=((0, 0.01725, 0.04), (0.005, 0.00675, 0.0375)), pad=((0, 0.01255, 0.058), (0.0035, 0.004, 0.0165)), bar_friction="0 0 0"),
    "round1": dict(bar=((0, 0.009, 0.035), (0.008, 0.006, 0.035)), pad=((0, 0.002, 0.062), (0.008, 0.0015, 0.008)), bar_friction=None),
}
RETHINK_FINGER_STYLE = "round1"
#: finger_inertia: rotational inertia of each finger body, 0.01 kg m^2 as recalled from rethink_gripper.xml (an odd value for a 20 g part, but it is what gives the
#: Sawyer's wrist joints enough inertia for the JOINT_VELOCITY law: with the 1e-5 of round 1 the last joint carried ~1.5e-3 kg m^2, below the dt kv / 2 = 4e-3 that
#: explicit Euler needs for a velocity gain kv = 4 at dt = 2 ms, and the wrist chattered).  Committed Sawyer policies, per run here vs their last logged returns
#: (profiles/r2_sawyer_fit_cpu.txt): Lift-Sawyer-JV 5 / 12 / 21 / 3 / 16 -> 88 / 58 / 23 / 22 / 50 against 84 / 53 / 20 / 12 / 49 logged; Stack-Sawyer (JV + OSC) log-return
#: loss 0.148 -> 0.018; Lift-Sawyer-OSC mean 146 -> 185 of 348; Door-Sawyer loss 2.36 -> 1.02.
RETHINK = dict(hand_yaw=0.0, finger_inertia=0.01, kp=1000.0, force=20.0, pad_friction="2 0.05 0.0001", pad_solref="0.01 0.5", finger_damping=100.0, finger_armature=1.0, finger_frictionloss=1.0)


def rethink_gripper(pf: str) -> str:
    """Rethink two-finger gripper for Sawyer: slide fingers along ±y."""
    pad = f'{ROBOT_COL} condim="4" friction="{RETHINK["pad_friction"]}" solref="{RETHINK["pad_solref"]}"'
    G = RETHINK_FINGERS[RETHINK_FINGER_STYLE]
    bar = ROBOT_COL + (f' friction="{G["bar_friction"]}"' if G["bar_friction"] else "")
    (bp, bs), (pp, ps) = G["bar"], G["pad"]
    m = lambda v: (v[0], -v[1], v[2])
    # hand_yaw (degrees): extra rotation of the `right_hand` BODY FRAME about its z axis with the gripper counter-rotated inside it, i.e. the same physical hand
    # with a differently oriented eef frame -- it changes only the eef_quat observation (the OSC law rotates in the world frame).
    yaw = np.deg2rad(90.0 + RETHINK["hand_yaw"]); cy = np.deg2rad(-RETHINK["hand_yaw"])
    qh, qc = f"{np.cos(yaw / 2):.10g} 0 0 {np.sin(yaw / 2):.10g}", f"{np.cos(cy / 2):.10g} 0 0 {np.sin(cy / 2):.10g}"
    return f'''
<body name="{pf}right_hand" pos="0 0 0.0245" quat="{qh}">
  <inertial pos="0 0 0.03" mass="0.3" diaginertia="0.001 0.001 0.001"/>
  <geom name="{pf}hand_col" type="box" pos="0 0 0.03" quat="{qc}" size="0.03 0.06 0.03" {ROBOT_COL}/>
  <site name="{pf}grip_site" pos="0 0 0.109"/>
  <body name="{pf}leftfinger" pos="{_f(_rotz(cy, (0, 0.01, 0.0444)))}" quat="{qc}">
    <inertial pos="0 0 0.03" mass="0.02" diaginertia="{_f([RETHINK["finger_inertia"]] * 3)}"/>
    <joint name="{pf}finger_joint1" type="slide" axis="0 1 0" range="-0.0115 0.020833" damping="{RETHINK["finger_damping"]}" armature="{RETHINK["finger_armature"]}" frictionloss="{RETHINK["finger_frictionloss"]}"/>
    <geom name="{pf}finger1_col" type="box" pos="{_f(bp)}" size="{_f(bs)}" {bar}/>
    <geom name="{pf}finger1_pad" type="box" pos="{_f(pp)}" size="{_f(ps)}" {pad}/>
  </body>
  <body name="{pf}rightfinger" pos="{_f(_rotz(cy, (0, -0.01, 0.0444)))}" quat="{qc}">
    <inertial pos="0 0 0.03" mass="0.02" diaginertia="{_f([RETHINK["finger_inertia"]] * 3)}"/>
    <joint name="{pf}finger_joint2" type="slide" axis="0 1 0" range="-0.020833 0.0115" damping="{RETHINK["finger_damping"]}" armature="{RETHINK["finger_armature"]}" frictionloss="{RETHINK["finger_frictionloss"]}"/>
    <geom name="{pf}finger2_col" type="box" pos="{_f(m(bp))}" size="{_f(bs)}" {bar}/>
    <geom name="{pf}finger2_pad" type="box" pos="{_f(m(pp))}" size="{_f(ps)}" {pad}/>
  </body>
</body>'''


def rethink_gripper_actuators(pf: str) -> str:
    kp, fr = RETHINK["kp"], RETHINK["force"]
    return (f'<position name="{pf}gripper_finger_joint1" joint="{pf}finger_joint1" kp="{kp}" ctrlrange="-0.0115 0.020833" forcerange="{-fr} {fr}"/>'
            f'<position name="{pf}gripper_finger_joint2" joint="{pf}finger_joint2" kp="{kp}" ctrlrange="-0.020833 0.0115" forcerange="{-fr} {fr}"/>')


# ----------------------------------------------------------------------------- robots
#: collide the Panda's arm links (capsule envelopes) with the scene, not only its hand and fingers
ARM_LINK_CAPSULES = False


def null_gripper(pf: str, hand_pos: str, hand_quat: str, payload: str = "") -> str:
    """robosuite's NullGripper (gripper_types=None, e.g. TwoArmPegInHole): the bare `right_hand` body with the `grip_site` at its origin; `payload` = bodies rigidly
    attached to the hand (the peg / the plate with the hole)."""
    return (f'<body name="{pf}right_hand" pos="{hand_pos}" quat="{hand_quat}">'
            f'<inertial pos="0 0 0" mass="0.5" diaginertia="0.002 0.002 0.002"/>'
            f'<site name="{pf}grip_site" pos="0 0 0"/>{payload}</body>')


def panda(pf: str, base_pos, base_quat=(1, 0, 0, 0), gripper="default", payload: str = "") -> str:
    """Franka Emika Panda, 7 hinge joints about local z (link frames: SURVEY.md A.5)."""
    lim = [(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
           (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)]
    frames = [("0 0 0.333", "1 0 0 0"), ("0 0 0", "0.7071068 -0.7071068 0 0"),
              ("0 -0.316 0", "0.7071068 0.7071068 0 0"), ("0.0825 0 0", "0.7071068 0.7071068 0 0"),
              ("-0.0825 0.384 0", "0.7071068 -0.7071068 0 0"), ("0 0 0", "0.7071068 0.7071068 0 0"),
              ("0.088 0 0", "0.7071068 0.7071068 0 0")]
    inert = [("0 0 -0.07", 3, 0.3), ("0 -0.1 0", 3, 0.3), ("0.04 0 -0.05", 2, 0.2), ("-0.04 0.05 0", 2, 0.2),
             ("0 0 -0.15", 2, 0.2), ("0.06 0 0", 1.5, 0.1), ("0 0 0.08", 0.5, 0.05)]
    damp = [0.1, 0.1, 0.1, 0.1, 0.1, 0.01, 0.01]
    # arm-link collision capsules (robosuite collides the links' meshes): rough envelopes of the public link geometry, in the link frames above.  Off unless
    # ARM_LINK_CAPSULES says so for the task being built (model/tasks.py): they cost collision pairs on every env, and only tasks where the arm itself touches
    # the scene need them.
    caps = {1: ("0 0 -0.15 0 0 0", 0.06), 2: ("0 0 0 0 -0.316 0", 0.06), 4: ("0 0 0 -0.0825 0.384 0", 0.06), 6: ("0 0 0 0.088 0 0", 0.05), 7: ("0 0 0 0 0 0.107", 0.04)}
    s = f'<body name="{pf}base" pos="{_f(base_pos)}" quat="{_f(base_quat)}">'
    s += f'<inertial pos="0 0 0.05" mass="4" diaginertia="0.4 0.4 0.4"/>'
    for i in range(7):
        s += (f'<body name="{pf}link{i + 1}" pos="{frames[i][0]}" quat="{frames[i][1]}">'
              f'<inertial pos="{inert[i][0]}" mass="{inert[i][1]}" diaginertia="{inert[i][2]} {inert[i][2]} {inert[i][2]}"/>'
              f'<joint name="{pf}joint{i + 1}" type="hinge" axis="0 0 1" range="{lim[i][0]} {lim[i][1]}" damping="{damp[i]}"/>')
        if ARM_LINK_CAPSULES and (i + 1) in caps:
            s += f'<geom name="{pf}link{i + 1}_col" type="capsule" fromto="{caps[i + 1][0]}" size="{caps[i + 1][1]}" {ROBOT_COL}/>'
    s += panda_gripper(pf) if gripper == "default" else null_gripper(pf, "0 0 0.107", "0.9238795 0 0 -0.3826834", payload)
    s += "</body>" * 8
    return s


def panda_actuators(pf: str, gripper="default") -> str:
    tl = [80, 80, 80, 80, 12, 12, 12]
    s = "".join(f'<motor name="{pf}torq_j{i + 1}" joint="{pf}joint{i + 1}" ctrlrange="{-tl[i]} {tl[i]}" ctrllimited="true"/>'
                for i in range(7))
    return s + (panda_gripper_actuators(pf) if gripper == "default" else "")


PANDA_INIT_QPOS = [0, PI / 16.0, 0.00, -PI / 2.0 - PI / 3.0, 0.00, PI - 0.2, PI / 4]
PANDA_GRIP_INIT = [0.020833, -0.020833]
PANDA_GRIP_SIGN = [-1.0, 1.0]


SAWYER = dict(damping=0.1, inertia_scale=1.0, armature=0.0, frictionloss=0.0)          # joint-space dynamics switches (what-if studies; defaults = round 1)


def sawyer(pf: str, base_pos, base_quat=(1, 0, 0, 0), gripper="default", payload: str = "") -> str:
    """Rethink Sawyer, 7 hinge joints about local z (public URDF link frames)."""
    lim = [(-3.0503, 3.0503), (-3.8095, 2.2736), (-3.0426, 3.0426), (-3.0439, 3.0439),
           (-2.9761, 2.9761), (-2.9761, 2.9761), (-4.7124, 4.7124)]
    frames = [("0 0 0.08", "1 0 0 0"), ("0.081 0.05 0.237", "0.5 -0.5 0.5 0.5"),
              ("0 -0.14 0.1425", "0.7071068 0.7071068 0 0"), ("0 -0.042 0.26", "0.7071068 -0.7071068 0 0"),
              ("0 -0.125 -0.1265", "0.7071068 0.7071068 0 0"), ("0 0.031 0.275", "0.7071068 -0.7071068 0 0"),
              ("0 -0.11 0.1053", "0.0616248 0.06163 -0.704416 0.704416")]
    inert = [("0.024 0.014 0.136", 5.32, 0.05), ("-0.003 -0.002 0.035", 4.51, 0.03), ("-0.002 -0.03 0.087", 1.75, 0.02),
             ("-0.002 0.0013 0.0018", 2.51, 0.01), ("0.0026 -0.02 0.09", 1.12, 0.01), ("-0.00005 0.003 0.02", 1.56, 0.004),
             ("0.008 0.005 -0.003", 0.33, 0.0003)]
    s = f'<body name="{pf}base" pos="{_f(base_pos)}" quat="{_f(base_quat)}">'
    s += f'<inertial pos="0 0 0.04" mass="2" diaginertia="0.02 0.02 0.02"/>'
    for i in range(7):
        s += (f'<body name="{pf}link{i + 1}" pos="{frames[i][0]}" quat="{frames[i][1]}">'
              f'<inertial pos="{inert[i][0]}" mass="{inert[i][1]}" diaginertia="{_f([inert[i][2] * SAWYER["inertia_scale"]] * 3)}"/>'
              f'<joint name="{pf}joint{i + 1}" type="hinge" axis="0 0 1" range="{lim[i][0]} {lim[i][1]}" damping="{SAWYER["damping"]}" '
              f'armature="{SAWYER["armature"]}" frictionloss="{SAWYER["frictionloss"]}"/>')
    s += rethink_gripper(pf) if gripper == "default" else null_gripper(pf, "0 0 0.0245", "0.7071068 0 0 0.7071068", payload)
    s += "</body>" * 8
    return s


def sawyer_actuators(pf: str, gripper="default") -> str:
    tl = [80, 80, 40, 40, 9, 9, 9]
    s = "".join(f'<motor name="{pf}torq_j{i + 1}" joint="{pf}joint{i + 1}" ctrlrange="{-tl[i]} {tl[i]}" ctrllimited="true"/>'
                for i in range(7))
    return s + (rethink_gripper_actuators(pf) if gripper == "default" else "")


SAWYER_INIT_QPOS = [0, -1.18, 0.00, 2.18, 0.00, 0.57, 3.3161]
SAWYER_GRIP_INIT = [0.020833, -0.020833]
#: a positive gripper action CLOSES (robosuite: -1 open, +1 closed): finger 1 (on +y, open at +0.020833) moves to its lower limit.  Round 1 had the signs the other
#: way round; the committed Lift-Sawyer policies never lifted the cube then (best episode 228 of 486) and do with these (best 485.7).
SAWYER_GRIP_SIGN = [-1.0, 1.0]

ROBOTS = {
    "Panda": dict(body=panda, act=panda_actuators, init_qpos=PANDA_INIT_QPOS, grip_init=PANDA_GRIP_INIT,
                  grip_sign=PANDA_GRIP_SIGN),
    "Sawyer": dict(body=sawyer, act=sawyer_actuators, init_qpos=SAWYER_INIT_QPOS, grip_init=SAWYER_GRIP_INIT,
                   grip_sign=SAWYER_GRIP_SIGN),
}


# ----------------------------------------------------------------------------- arena / objects
TABLE_HEIGHT = 0.8
TABLE_FULL = (0.8, 0.8, 0.05)
ROBOT_BASE_Z = 0.912


def table_arena(full=TABLE_FULL, friction=(1, 0.005, 0.0001), offset=(0.0, 0.0)) -> str:
    hx, hy, hz = full[0] / 2, full[1] / 2, full[2] / 2
    return (f'<geom name="floor" type="plane" pos="0 0 0" size="3 3 0.125" {WORLD_COL}/>'
            f'<body name="table" pos="{offset[0]} {offset[1]} {TABLE_HEIGHT - hz}">'
            f'<geom name="table_collision" type="box" size="{hx} {hy} {hz}" friction="{_f(friction)}" {WORLD_COL}/>'
            f'</body>')


def box_object(name, half, pos, density=1000, friction=(1, 0.005, 0.0001), solref=(0.02, 1.0),
               solimp=(0.9, 0.95, 0.001)) -> str:
    return (f'<body name="{name}" pos="{_f(pos)}"><freejoint name="{name}_joint"/>'
            f'<geom name="{name}_g0" type="box" size="{_f(half)}" density="{density}" friction="{_f(friction)}" '
            f'solref="{_f(solref)}" solimp="{_f(solimp)}" {WORLD_COL}/></body>')


def door_object(pos=(0.12, -0.2, TABLE_HEIGHT), yaw=-PI / 2, hinge_damping=0.1, hinge_frictionloss=0.0) -> str:
    """Door with a latch handle (robosuite DoorObject with lock=True, authored from primitives).  The root body is
    FIXED to the world; its pose is re-sampled at every reset (rsb_task.place_body).  Local frame: x = door width, y = door normal
    (the handle sits on the -y face), z = up; hinge on the +x post, door opens towards +y (away from the robot after the -90 deg yaw)."""
    q = f"{np.cos(yaw / 2):.10g} 0 0 {np.sin(yaw / 2):.10g}"
    col = WORLD_COL + ' friction="1 0.005 0.0001"'
    return f'''
<body name="door_root" pos="{_f(pos)}" quat="{q}">
  <geom name="door_post_l" type="box" pos="-0.165 0 0.16" size="0.015 0.015 0.16" {col}/>
  <geom name="door_post_r" type="box" pos="0.165 0 0.16" size="0.015 0.015 0.16" {col}/>
  <body name="door" pos="0.15 0 0.16">
    <inertial pos="-0.15 0 0" mass="1.0" diaginertia="0.0075 0.0141 0.0066"/>
    <joint name="door_hinge" type="hinge" axis="0 0 1" range="0 0.4" damping="{hinge_damping}" frictionloss="{hinge_frictionloss}"/>
    <geom name="door_panel" type="box" pos="-0.15 0 0" size="0.14 0.01 0.15" {col}/>
    <body name="latch" pos="-0.25 0 0">
      <inertial pos="0.03 -0.04 0" mass="0.1" diaginertia="0.0001 0.0001 0.0001"/>
      <joint name="latch_joint" type="hinge" axis="0 1 0" range="-1.57 0" damping="0.1" frictionloss="0.1" stiffness="0.5" springref="0"/>
      <geom name="latch_stem" type="box" pos="0 -0.03 0" size="0.008 0.02 0.008" {col}/>
      <geom name="latch_handle" type="box" pos="0.04 -0.05 0" size="0.05 0.01 0.01" {col}/>
      <site name="door_handle" pos="0.04 -0.05 0"/>
    </body>
  </body>
</body>'''


#: which door the Door task stands on its table: "robosuite_recalled" (door_object_recalled below, the default since round 2) or "round1" (door_object above, a
#: half-size door calibrated by the reset distance only).  Chosen by the transfer of the committed Door policies, mean return here / logged over 5 seeds:
#: round1 1.4 / 140 (Panda JOINT_VELOCITY), 52 / 368 (Panda OSC_POSE), 0.5 / 96 and 47 / 267 (Sawyer); recalled door with the hinge / latch choices below 247 / 140,
#: 216 / 368, 59 / 96 and 261 / 267 (profiles/r2_policy_transfer_all.txt, r2_door_fit_cpu.txt).
DOOR_STYLE = "robosuite_recalled"
#: the Door task's table: robosuite's Door env uses a narrow table beside the robot (full size 0.8 x 0.3 x 0.05 at offset (-0.2, -0.35, 0.8)), as recalled
DOOR_TABLE_FULL, DOOR_TABLE_OFFSET = (0.8, 0.3, 0.05), (-0.2, -0.35)
#: latch handle: spring and friction as recalled; inertia = the physical estimate for a 0.1 kg handle (the recalled XML value 0.048 / 0.041 / 0.011 transfers no
#: better); axis -y: the far end of the handle bar is pushed DOWN to turn it (with +y, as first restated, the policies hardly turn it).  `bolt`: robosuite's
#: use_latch=True locks the door until the handle is turned, but that geometry is not recalled and every lock tried here -- a bolt behind the post (True) or a
#: bolt / strike pair that collide only with each other and release at a chosen handle angle ("strike") -- makes the committed policies transfer WORSE (one OSC
#: policy that logs 444 never touches the handle), so the door is unlocked (False).
DOOR_LATCH = dict(stiffness=1.0, damping=0.0, frictionloss=0.1, inertia=(0.001, 0.001, 0.001), bolt=False, axis=(0, -1, 0),
                  bolt_pos=(-0.065, 0.044, 0.0), bolt_size=(0.045, 0.01, 0.015), bolt_r=(0.03, 0.048), bolt_h=0.004, strike_h=0.006)


#: hinge: axis -z = the panel swings AWAY from the side the handle is on (the robot pushes); damping / frictionloss 1 / 1 are the values of door_lock.xml as recalled.
#: Both chosen by the committed Door-Panda policies (tools/fit_assets_cpu.py on the CPU oracle, profiles/r2_door_fit_cpu.txt): with +z (pull) the OSC policies that
#: log 408-466 score 87-202, with -z 322-470; with the DoorObject(friction=0, damping=0.1) values recalled from door.py the JOINT_VELOCITY policies that FAIL in
#: their own logs (14, 105) open the door here (347, 437) -- a velocity-controlled arm pushes with a few newtons, which a frictionloss of 1 N m holds back.
DOOR_HINGE = dict(axis=(0, 0, -1), damping=1.0, frictionloss=1.0)
DOOR_PANEL = dict(mass=2.43455)


def door_object_recalled(pos=(-0.12, -0.35, TABLE_HEIGHT + 0.3), yaw=-PI / 2, hinge_damping=0.1, hinge_frictionloss=0.0) -> str:
    """robosuite's DoorObject(lock=True) (objects/door_lock.xml) restated from memory with box primitives -- every number here is UPSTREAM RECALL, kept because the
    committed Door policies transfer better with it than with the round-1 stand-in (COMPAT.md): a 0.44 x 0.04 x 0.58 m panel between two 0.6 m posts, hinged at one
    post (range 0..0.4 rad; direction, damping and frictionloss: DOOR_HINGE); a spring-loaded latch handle
    (stiffness 1, frictionloss 0.1, range -pi/2..0) sticking 0.10 m out of the panel with the `handle` site at the far end of its 0.15 m grip bar; optionally a
    lock (DOOR_LATCH["bolt"]).  The object's frame body (pos 0 0.22 0, yaw -90 deg in the object root)
    is folded into the coordinates: frame-local (x, y) -> root (y, 0.22 - x).  Root origin = centre height of the door (bottom_offset -0.3)."""
    q = f"{np.cos(yaw / 2):.10g} 0 0 {np.sin(yaw / 2):.10g}"
    col = WORLD_COL + ' friction="1 1 1"'
    L = DOOR_LATCH
    bolt = f'<geom name="latch_bolt" type="box" pos="{_f(L["bolt_pos"])}" size="{_f(L["bolt_size"])}" {col}/>' if L["bolt"] is True else ""
    strike = ""
    if L["bolt"] == "strike":
        # A lock of our own making (robosuite's is not recalled): a short bolt on the latch, r0..r1 from the latch axis, and a strike plate on the frame that the bolt
        # runs into when the door moves in its opening direction.  Both collide ONLY with each other (contype / conaffinity bit 2).  Turning the handle by more than
        # atan((strike_h + bolt_h) / r0) lifts the bolt over the plate.
        r0, r1 = L["bolt_r"]; bh, sh, yb = L["bolt_h"], L["strike_h"], 0.035
        sgn = 1.0 if DOOR_HINGE["axis"][2] < 0 else -1.0                # the panel swings towards +y (door frame) for axis -z: the plate sits on that side of the bolt
        bolt = f'<geom name="latch_bolt" type="box" pos="{-(r0 + r1) / 2} {yb} 0" size="{(r1 - r0) / 2} 0.006 {bh}" contype="0" conaffinity="2"/>'
        xs0, xs1 = r0 + 0.005, r1 + 0.004                                 # plate extent along the bolt, in the latch frame (-x)
        xf, yf = 0.125 - (xs0 + xs1) / 2, yb + sgn * 0.0125
        strike = f'<geom name="door_strike" type="box" pos="{yf} {0.22 - xf} -0.025" size="0.006 {(xs1 - xs0) / 2} {sh}" contype="2" conaffinity="0"/>'

    return f'''
<body name="door_root" pos="{_f(pos)}" quat="{q}">
  <geom name="door_post_l" type="box" pos="0 0.175 0" size="0.03 0.021 0.3" {col}/>
  <geom name="door_post_r" type="box" pos="0 -0.335 0" size="0.03 0.021 0.3" {col}/>
  {strike}
  <body name="door" pos="0 -0.08 0" quat="0.7071068 0 0 -0.7071068">
    <inertial pos="0.0296816 -0.00152345 0" mass="{DOOR_PANEL["mass"]}" diaginertia="{_f(np.array([0.0521615, 0.0913751, 0.043714]) * DOOR_PANEL["mass"] / 2.43455)}"/>
    <joint name="door_hinge" type="hinge" pos="0.255 0 0" axis="{_f(DOOR_HINGE["axis"])}" range="0 0.4" damping="{DOOR_HINGE["damping"]}" frictionloss="{DOOR_HINGE["frictionloss"]}"/>
    <geom name="door_panel" type="box" size="0.22 0.02 0.29" {col}/>
    <body name="latch" pos="-0.175 0 -0.025">
      <inertial pos="-0.017762 0.0138544 0" mass="0.1" diaginertia="{_f(L["inertia"])}"/>
      <joint name="latch_joint" type="hinge" axis="{_f(L["axis"])}" range="-1.57 0" damping="{L["damping"]}" frictionloss="{L["frictionloss"]}" stiffness="{L["stiffness"]}" springref="0"/>
      <geom name="latch_stem" type="box" pos="0 -0.0625 0" size="0.02 0.0625 0.02" {col}/>
      <geom name="latch_handle" type="box" pos="0.075 -0.10 0" size="0.075 0.015 0.02" {col}/>
      {bolt}
      <site name="door_handle" pos="0.125 -0.10 0"/>
    </body>
  </body>
</body>'''


#: TwoArmLift pot, robosuite's PotWithHandlesObject as recalled: body half size 0.07, wall thickness 0.025 (2.8 kg at density 1000), handle loops 9 cm out at
#: z = 0.07 - 0.01 with a main bar of half length 0.045 + 0.01 and two side bars (half 0.01 x 0.045 x 0.01 at x = +-0.045) that collide.  Round 1 had 5 mm walls
#: (1.1 kg), bars at z = 0.05 of half length 0.045 and visual-only side bars: POT = dict(thickness=0.005, handle_z=0.05, bar_half=0.045, side_bars=False).
#: Transfer of the 5 committed TwoArmLift-PandaPanda policies (mean return here / logged 110; profiles/r2_policy_transfer_pot.txt): round-1 pot 52, recalled pot 61.
POT = dict(thickness=0.025, handle_z=0.06, bar_half=0.055, side_bars=True)
#: |y| of the two robot bases in TwoArmLift's single-arm-opposed layout: robosuite rotates base_xpos_offset["table"](0.8) = (-0.56, 0, 0) by +-90 degrees.
#: Round 1 used 0.69 (fitted to the logged epoch-0 return); the committed policies say 0.56: 80 vs 52 with the round-1 pot, 86 vs 61 with the recalled one.
TWO_ARM_BASE_Y = 0.56
#: TwoArmLift pot yaw at reset: True = pi +- pi/3 with robot0's handle on the pot frame's +y side (robosuite's sampler as recalled: rotation=(pi - pi/3, pi + pi/3));
#: False = +- pi/3 with robot0's handle on -y (round 1).  Same scene either way (the pot is symmetric); only the pot_quat observation differs by half a turn.
TWO_ARM_POT_YAW_PI = False


def pot_with_handles(name="pot", pos=(0, 0, TABLE_HEIGHT + 0.07), density=1000) -> str:
    """robosuite PotWithHandlesObject from boxes: hollow body (base + 4 walls, half size 0.07) and two handle grip bars 9 cm out along
    -y / +y (geometry switches: POT above)."""
    col = WORLD_COL + f' density="{density}" friction="1 0.005 0.0001"'
    h, t = 0.07, POT["thickness"]
    bt = max(0.01, t / 2)                                    # half thickness of the base slab
    g = [f'<geom name="{name}_base" type="box" pos="0 0 {-h + bt}" size="{h} {h} {bt}" {col}/>',
         f'<geom name="{name}_wall_xp" type="box" pos="{h - t} 0 {bt}" size="{t} {h} {h - bt}" {col}/>',
         f'<geom name="{name}_wall_xn" type="box" pos="{-h + t} 0 {bt}" size="{t} {h} {h - bt}" {col}/>',
         f'<geom name="{name}_wall_yp" type="box" pos="0 {h - t} {bt}" size="{h - 2 * t} {t} {h - bt}" {col}/>',
         f'<geom name="{name}_wall_yn" type="box" pos="0 {-h + t} {bt}" size="{h - 2 * t} {t} {h - bt}" {col}/>']
    for k, sg in enumerate((-1, 1)):
        y = sg * (h + 0.09)
        g.append(f'<geom name="{name}_handle{k}" type="box" pos="0 {y} {POT["handle_z"]}" size="{POT["bar_half"]} 0.01 0.01" {col}/>')
        if POT["side_bars"]:
            for sx in (-1, 1):
                g.append(f'<geom name="{name}_handle{k}_s{(sx + 1) // 2}" type="box" pos="{sx * 0.045} {sg * (h + 0.045)} {POT["handle_z"]}" size="0.01 0.045 0.01" {col}/>')
        g.append(f'<site name="{name}_handle{k}" pos="0 {y} {POT["handle_z"]}"/>')
    return f'<body name="{name}" pos="{_f(pos)}"><freejoint name="{name}_joint"/>' + "".join(g) + "</body>"


# ----------------------------------------------------------------------------- PickPlace: bins arena and the pick objects
#: robosuite's BinsArena (arenas/bins_arena.xml) as recalled: two open bins side by side, each a 0.4 x 0.5 m floor slab of half thickness 0.02 centred at the
#: bin position (top surface at z = 0.82 = table_full_size[2] of the PickPlace env) with four 0.1 m walls.  The bin-2 partition lines and the legs are visual.
BIN1_POS, BIN2_POS = (0.1, -0.25, 0.8), (0.1, 0.28, 0.8)
BIN_SIZE = (0.39, 0.49, 0.82)                  # PickPlace's table_full_size: the sampler's and the reward's bin extent
BIN_FLOOR_HALF = (0.2, 0.25, 0.02)
#: base_xpos_offset["bins"] of robosuite's Panda and Sawyer models
BINS_ROBOT_BASE = (-0.5, -0.1)
#: the pick objects are meshes upstream (collision = the mesh's convex hull); here each is ONE box of the hull's extents -- half sizes from the objects' bottom / top /
#: horizontal-radius sites as recalled; density 100, friction (0.95, 0.3, 0.1), solref (0.001, 1), solimp (0.998, 0.998, 0.001) as in the object XMLs.
#: `bin_id` is the object's index in robosuite's item list [Milk, Bread, Cereal, Can] (its target quadrant of bin 2).  The can is a cylinder upstream: a square
#: prism of the same width stands and is grasped like it but does not roll (no cylinder narrow phase in the kernels).
PICK_OBJECTS = {
    "Milk": dict(half=(0.0225, 0.0225, 0.08), bin_id=0),
    "Bread": dict(half=(0.02, 0.04, 0.0225), bin_id=1),
    "Cereal": dict(half=(0.02, 0.05, 0.1), bin_id=2),
    "Can": dict(half=(0.03, 0.03, 0.06), bin_id=3),
}


def bins_arena(friction=(1, 0.005, 0.0001)) -> str:
    s = f'<geom name="floor" type="plane" pos="0 0 0" size="3 3 0.125" {WORLD_COL}/>'
    hx, hy, hz = BIN_FLOOR_HALF
    for name, pos in (("bin1", BIN1_POS), ("bin2", BIN2_POS)):
        col = f'friction="{_f(friction)}" {WORLD_COL}'
        s += (f'<body name="{name}" pos="{_f(pos)}">'
              f'<geom name="{name}_floor" type="box" size="{hx} {hy} {hz}" {col}/>'
              f'<geom name="{name}_wall_yp" type="box" pos="0 {hy} 0.05" size="{hx + 0.01} 0.01 0.05" {col}/>'
              f'<geom name="{name}_wall_yn" type="box" pos="0 {-hy} 0.05" size="{hx + 0.01} 0.01 0.05" {col}/>'
              f'<geom name="{name}_wall_xp" type="box" pos="{hx} 0 0.05" size="0.01 {hy} 0.05" {col}/>'
              f'<geom name="{name}_wall_xn" type="box" pos="{-hx} 0 0.05" size="0.01 {hy} 0.05" {col}/>'
              f'</body>')
    return s


def pick_object(kind: str, pos) -> str:
    return box_object(kind, PICK_OBJECTS[kind]["half"], pos, density=100, friction=(0.95, 0.3, 0.1), solref=(0.001, 1.0), solimp=(0.998, 0.998, 0.001))


# ----------------------------------------------------------------------------- TwoArmPegInHole: peg and plate, rigidly attached to the two hands
#: robosuite's TwoArmPegInHole as recalled: no grippers, no table (EmptyArena); robot 0 carries a cylinder peg at (0, 0, 0.15) of its hand (axis = the hand's z,
#: half length 0.13, radius drawn once per model from U(0.015, 0.03): the mean here), robot 1 a plate with a hole at (0.11, 0, 0.17) / quat (0, 0, 0.707, 0.707) of
#: its hand; the hole's centre is 0.1 along the plate body's x axis, its normal is the plate's z axis (TwoArmPegInHole._compute_orientation).  The plate's own
#: geometry (plate-with-hole.xml) is NOT recalled: a 0.2 m square frame of four boxes around a square hole of half width `hole_half`, wide enough for the largest peg
#: plus the success tolerance d < 0.06.  The peg is a capsule of the same radius and overall length (the kernels have no cylinder narrow phase; the side surface,
#: which is what meets the hole's rim, is the same).  |y| of the bases: base_xpos_offset["empty"] = (-0.6, 0, 0) turned by +-90 degrees.
PEG_IN_HOLE = dict(peg_radius=0.0225, peg_half_length=0.13, peg_pos=(0, 0, 0.15), hole_pos=(0.11, 0, 0.17), hole_quat=(0, 0, 0.7071068, 0.7071068),
                   hole_center=0.1, hole_half=0.06, plate_half=0.1, plate_thickness=0.01, base_y=0.6, plate_density=500, collide=True)


def peg_payload(name="peg") -> str:
    P = PEG_IN_HOLE
    r, hl = P["peg_radius"], P["peg_half_length"]
    return (f'<body name="{name}" pos="{_f(P["peg_pos"])}">'
            f'<geom name="{name}_g0" type="capsule" fromto="0 0 {-(hl - r)} 0 0 {hl - r}" size="{r}" density="1000" friction="1 0.005 0.0001" {WORLD_COL}/></body>')


def hole_payload(name="hole") -> str:
    P = PEG_IN_HOLE
    c, h, w, t = P["hole_center"], P["hole_half"], P["plate_half"], P["plate_thickness"]
    col = f'density="{P["plate_density"]}" friction="1 0.005 0.0001" {WORLD_COL if P["collide"] else NO_COL}'
    bar = (w - h) / 2                                   # half width of the frame's bars
    g = (f'<geom name="{name}_xp" type="box" pos="{c + h + bar} 0 0" size="{bar} {w} {t}" {col}/>'
         f'<geom name="{name}_xn" type="box" pos="{c - h - bar} 0 0" size="{bar} {w} {t}" {col}/>'
         f'<geom name="{name}_yp" type="box" pos="{c} {h + bar} 0" size="{h} {bar} {t}" {col}/>'
         f'<geom name="{name}_yn" type="box" pos="{c} {-h - bar} 0" size="{h} {bar} {t}" {col}/>')
    return f'<body name="{name}" pos="{_f(P["hole_pos"])}" quat="{_f(P["hole_quat"])}">{g}<site name="{name}_center" pos="{c} 0 0"/></body>'


def empty_arena() -> str:
    return f'<geom name="floor" type="plane" pos="0 0 0" size="3 3 0.125" {WORLD_COL}/>'


# ----------------------------------------------------------------------------- NutAssembly: pegs arena and the nuts
#: robosuite's PegsArena / NutAssembly as recalled: table 0.45 x 0.69 m with its top at z = 0.82 (table_offset), the square peg (box, half 0.016 x 0.016 x 0.1) at
#: (0.23, 0.1, 0.85) and the round peg (cylinder, radius 0.02, half height 0.1) at (0.23, -0.1, 0.85); nuts start at x in [-0.115, -0.11], y in [0.11, 0.225] (square) /
#: [-0.225, -0.11] (round), any yaw, dropped from 2 cm.  The round peg is a CAPSULE of the same radius and height here (no cylinder narrow phase; same side
#: surface).  The nut meshes' collision boxes (round-nut.xml / square-nut.xml) are NOT recalled: the round nut is an octagonal ring of eight boxes (inner apothem
#: `ring_in`, radial half thickness `ring_t`) 4 cm high with a handle bar reaching out to `handle_out` along +x, the LAST geom (robosuite's reach reward aims at the
#: nut's last geom).  Density 100, friction (0.95, 0.3, 0.1), solref (0.001, 1), solimp (0.998, 0.998, 0.001) as in robosuite's object XMLs.
NUT_TABLE_FULL, NUT_TABLE_Z = (0.45, 0.69, 0.05), 0.82
NUT_PEGS = {"Square": (0.23, 0.1), "Round": (0.23, -0.1)}
NUT_PLACE = {"Square": dict(x=(-0.115, -0.11), y=(0.11, 0.225)), "Round": dict(x=(-0.115, -0.11), y=(-0.225, -0.11))}
ROUND_NUT = dict(ring_in=0.028, ring_t=0.0075, half_h=0.02, handle_out=0.09, handle_half_w=0.01, handle_dir=0.0)      # handle_dir: direction of the handle bar in the nut frame (degrees from +x)


def pegs_arena(friction=(1, 0.005, 0.0001)) -> str:
    hx, hy, hz = NUT_TABLE_FULL[0] / 2, NUT_TABLE_FULL[1] / 2, NUT_TABLE_FULL[2] / 2
    col = f'friction="{_f(friction)}" {WORLD_COL}'
    (sx, sy), (rx, ry) = NUT_PEGS["Square"], NUT_PEGS["Round"]
    return (f'<geom name="floor" type="plane" pos="0 0 0" size="3 3 0.125" {WORLD_COL}/>'
            f'<body name="table" pos="0 0 {NUT_TABLE_Z - hz}"><geom name="table_collision" type="box" size="{hx} {hy} {hz}" {col}/></body>'
            f'<body name="peg1" pos="{sx} {sy} 0.85"><geom name="peg1_col" type="box" size="0.016 0.016 0.1" {col}/></body>'
            f'<body name="peg2" pos="{rx} {ry} 0.85"><geom name="peg2_col" type="capsule" fromto="0 0 -0.1 0 0 0.08" size="0.02" {col}/></body>')


def round_nut(name, pos) -> str:
    N = ROUND_NUT
    a_in, t, hh = N["ring_in"], N["ring_t"], N["half_h"]
    a_out = a_in + 2 * t
    col = f'density="100" friction="0.95 0.3 0.1" solref="0.001 1" solimp="0.998 0.998 0.001" {WORLD_COL}'
    g = ""
    for k in range(8):
        a = 2 * PI * k / 8
        g += (f'<geom name="{name}_ring{k}" type="box" pos="{(a_in + t) * np.cos(a):.8g} {(a_in + t) * np.sin(a):.8g} 0" '
              f'quat="{np.cos(a / 2):.8g} 0 0 {np.sin(a / 2):.8g}" size="{t} {a_out * np.tan(PI / 8):.8g} {hh}" {col}/>')
    hl = (N["handle_out"] - a_out) / 2
    hd = np.deg2rad(N["handle_dir"])
    g += (f'<geom name="{name}_handle" type="box" pos="{(a_out + hl) * np.cos(hd):.8g} {(a_out + hl) * np.sin(hd):.8g} 0" quat="{np.cos(hd / 2):.8g} 0 0 {np.sin(hd / 2):.8g}" '
          f'size="{hl} {N["handle_half_w"]} {hh}" {col}/>')
    return f'<body name="{name}" pos="{_f(pos)}"><freejoint name="{name}_joint"/>{g}</body>'


# ----------------------------------------------------------------------------- TwoArmHandoff: narrow table and the hammer
#: robosuite's TwoArmHandoff as recalled: table_full_size (0.8, 1.2, 0.05) of which only a quarter of the width is real -- a 0.8 x 0.3 m table at
#: (0, -0.45) with its top at 0.8, beside robot 0; the arms face each other along y.  `base_y`: where the bases stand; 0.76 = 0.16 + 1.2 / 2 is what the committed
#: runs' epoch-0 reward level says (reach term 0.125 (1 - tanh d0): logged 0.094-0.100 for Panda, 0.079-0.094 for Sawyer; 0.096 / 0.090 here; with 0.56 it
#: would be 0.079).  HammerObject (generated per model upstream with random sizes: handle radius U(0.015, 0.02), length U(0.1, 0.25), density U(100, 250), friction
#: U(3, 5), head density x 2): the MEANS here; the neck / face cylinders are boxes of the same cross-section area.  The hammer starts lying along y on the table, its
#: head towards robot 0 or robot 1 at random (a quarter turn about x, +- `tilt`).  Chosen by the committed policies: lying along x (a turn about y, as robosuite's
#: sampler argument rotation_axis='y' was first read) none of them holds the hammer; lying along y the best ones grasp, lift and hold it for whole episodes
#: (profiles/r2_policy_transfer_handoff_cpu.txt).
HANDOFF = dict(table_full=(0.8, 0.3, 0.05), table_offset=(0.0, -0.45), base_y=0.76, handle_radius=0.0175, handle_length=0.175, handle_density=175.0, handle_friction=4.0,
               head_density_ratio=2.0, head_half_ratio=1.1, place_x=(-0.1, 0.1), place_y=(-0.05, 0.05), tilt=0.1, lift_height=0.1)


def hammer(name, pos) -> str:
    H = HANDOFF
    r, hl = H["handle_radius"], H["handle_length"] / 2
    h = r * H["head_half_ratio"]
    zc = hl + h
    sq = 0.8862269                                   # side / diameter of the square with a circle's area
    col = lambda dens, fr: f'density="{dens}" friction="{fr} 0.005 0.0001" {WORLD_COL}'
    hd = H["handle_density"] * H["head_density_ratio"]
    g = (f'<geom name="{name}_handle" type="box" size="{r} {r} {hl}" {col(H["handle_density"], H["handle_friction"])}/>'
         f'<geom name="{name}_head" type="box" pos="0 0 {zc}" size="{2 * h} {h} {h}" {col(hd, 1)}/>'
         f'<geom name="{name}_neck" type="box" pos="{2.2 * h} 0 {zc}" size="{0.2 * h} {0.8 * h * sq} {0.8 * h * sq}" {col(hd, 1)}/>'
         f'<geom name="{name}_face" type="box" pos="{2.8 * h} 0 {zc}" size="{0.4 * h} {h * sq} {h * sq}" {col(hd, 1)}/>'
         f'<geom name="{name}_claw" type="box" pos="{-2 * h} 0 {zc}" quat="0.9238795 0 0.3826834 0" size="{0.7072 * h} {0.95 * h} {0.7072 * h}" {col(hd, 1)}/>')
    return f'<body name="{name}" pos="{_f(pos)}"><freejoint name="{name}_joint"/>{g}</body>'


def scene(world: str, actuators: str, extra: str = "") -> str:
    return f'<mujoco model="rsb">{BASE_OPTION}<worldbody>{world}</worldbody><actuator>{actuators}</actuator>{extra}</mujoco>'
