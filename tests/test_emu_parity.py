"""CPU parity of the device code (compiled for the host by tests/emu, lanes = fibers) against the fp64 oracle, for every BASELINE
config family.  This is the no-GPU guard for the logic of csrc/rsb_dev.h; the `-m gpu` tests run the same checks on the real kernels."""
import numpy as np
import pytest

from oracle.oracle import OracleEnv
from robosuite_benchmark_b200.controllers import load_controller_config
from robosuite_benchmark_b200.model.tasks import build_task
from tests.emu.emu import EmuEnv

CONFIGS = [("Lift", ["Panda"], "OSC_POSE", 42, 7), ("Lift", ["Panda"], "JOINT_VELOCITY", 42, 8), ("Lift", ["Sawyer"], "OSC_POSITION", 42, 4),
           ("Lift", ["Panda"], "JOINT_POSITION", 42, 8), ("Lift", ["Sawyer"], "JOINT_TORQUE", 42, 8),
           ("Door", ["Panda"], "JOINT_VELOCITY", 46, 8), ("Stack", ["Sawyer"], "OSC_POSE", 55, 7), ("TwoArmLift", ["Panda", "Panda"], "OSC_POSE", 89, 14),
           ("PickPlaceCan", ["Panda"], "OSC_POSE", 46, 7), ("PickPlaceMilk", ["Sawyer"], "OSC_POSE", 46, 7),
           ("TwoArmPegInHole", ["Panda", "Sawyer"], "OSC_POSE", 73, 12), ("NutAssemblyRound", ["Panda"], "OSC_POSE", 46, 7),
           ("TwoArmHandoff", ["Panda", "Sawyer"], "OSC_POSE", 86, 14)]


@pytest.mark.parametrize("env_name,robots,ctrl,obs_dim,act_dim", CONFIGS)
def test_control_step_matches_oracle(env_name, robots, ctrl, obs_dim, act_dim, lanes=32):
    m, t = build_task(env_name, robots, load_controller_config(default_controller=ctrl), ignore_done=True)
    assert (t["obs_dim"], t["act_dim"]) == (obs_dim, act_dim)          # dims pinned by the committed networks
    nc, ne = t["ncon_max"], t["nefc_max"]
    orc, emu = OracleEnv(m, t, ncon_max=nc, nefc_max=ne), EmuEnv(m, t, nc, ne, lanes=lanes)
    o1, o2 = orc.reset(seed=17, env_id=3), emu.reset(seed=17, env_id=3)
    assert np.abs(o1 - o2).max() < 2e-6
    for k in range(2):
        a = orc.random_action(17, 3, k)
        assert np.abs(a - emu.random_action(17, 3, k)).max() < 1e-6
        qp, qv, w, cs = orc.get_state()
        emu.set_state(qp, qv, w, cs, timestep=k, bpose=orc.get_bpose())
        o1, r1, _ = orc.step(a)
        o2, r2, _ = emu.step(a)
        qp, qv, _, _ = orc.get_state()
        qp2, qv2, _, _ = emu.get_state()
        assert np.abs(qp - qp2).max() <= 1e-4 and np.abs(qv - qv2).max() <= 1e-4      # north_star tolerance for one control step (fp32)
        assert np.abs(o1 - o2).max() <= 1e-4 and abs(r1 - r2) <= 1e-5


@pytest.mark.parametrize("env_name,robots,ctrl,obs_dim,act_dim", [c for c in CONFIGS if c[0] in ("Lift", "Door", "PickPlaceCan", "PickPlaceMilk", "TwoArmPegInHole", "NutAssemblyRound")])
def test_control_step_matches_oracle_16_lane_groups(env_name, robots, ctrl, obs_dim, act_dim):
    """Models with nv <= 16 run as 16-lane groups (two envs per warp) on the GPU (csrc/rsb_cuda.cu): the same check on the device code compiled for 16 lanes."""
    test_control_step_matches_oracle(env_name, robots, ctrl, obs_dim, act_dim, lanes=16)


def test_lane_order_independence():
    """A missing gsync() would make the result depend on the order in which the emulator runs the lanes."""
    m, t = build_task("Lift", "Panda", load_controller_config(default_controller="OSC_POSE"), ignore_done=True)
    outs = []
    for order in (0, 1, 2):
        emu = EmuEnv(m, t, 16, 64)
        emu.set_order(order)
        emu.reset(seed=5, env_id=1)
        outs.append(emu.step(emu.random_action(5, 1, 0))[0])
    emu.set_order(0)
    assert (outs[0] == outs[1]).all() and (outs[0] == outs[2]).all()


def test_done_protocol_and_terminated_episode():
    m, t = build_task("Lift", "Panda", load_controller_config(default_controller="OSC_POSE"), horizon=2, ignore_done=False)
    emu = EmuEnv(m, t, 16, 64)
    emu.reset(seed=1, env_id=0)
    assert emu.step(np.zeros(7))[2] is False
    assert emu.step(np.zeros(7))[2] is True
    with pytest.raises(ValueError):
        emu.step(np.zeros(7))


def _quat_xyzw_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@pytest.mark.parametrize("mode,sign", [("euler_transpose", -1.0), ("axis_angle", +1.0)])
def test_orientation_delta_convention(mode, sign):
    """OSC_POSE rotation actions (include/rsb_model.h RSB_ORI_DELTA_*): the shipped default turns the end effector by -d to first order (euler2mat(d)^T R_ee, the
    convention the committed 2020 policies transfer under), "axis_angle" by +d; device code == oracle under both."""
    cfg = load_controller_config(default_controller="OSC_POSE")
    assert cfg["orientation_delta"] == "euler_transpose"
    cfg["orientation_delta"] = mode
    m, t = build_task("Lift", "Panda", cfg, ignore_done=True)
    assert t["robot"][0]["ori_delta_mode"] == (0 if mode == "euler_transpose" else 1)
    orc, emu = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"]), EmuEnv(m, t, t["ncon_max"], t["nefc_max"])
    o0 = orc.reset(seed=3, env_id=0)
    emu.reset(seed=3, env_id=0)
    a = np.array([0.1, -0.2, 0.3, 0.7, -0.9, 1.0, 0.0])                # all three rotation components large: the two conventions differ at second order too
    qp, qv, w, cs = orc.get_state()
    emu.set_state(qp, qv, w, cs, timestep=0, bpose=orc.get_bpose())
    o1, r1, _ = orc.step(a)
    o2, r2, _ = emu.step(a)
    assert np.abs(o1 - o2).max() <= 1e-4 and abs(r1 - r2) <= 1e-5
    # direction: a pure +z rotation action for a few control steps
    orc.reset(seed=3, env_id=0)
    for _ in range(6):
        o = orc.step(np.array([0, 0, 0, 0, 0, 1.0, 0]))[0]
    R0, R1 = _quat_xyzw_to_mat(o0[24:28]), _quat_xyzw_to_mat(o[24:28])
    D = R1 @ R0.T                                                       # world-frame rotation the end effector made
    wz = 0.5 * (D[1, 0] - D[0, 1])                                      # sin(angle) * axis_z
    assert sign * wz > 0.05, (mode, wz)


@pytest.mark.parametrize("law", ["kv", "kp"])
def test_joint_velocity_laws(law):
    """JOINT_VELOCITY: the default is robosuite v1.0's proportional law (kv = 4; the law the committed JV policies transfer under, COMPAT.md); a config with "kp"
    instead of "kv" selects the PID law of robosuite >= 1.1 (integral with anti-windup, 5-sample derivative average).  Device code == oracle over several control
    steps, so that the integrator / derivative state carried between steps is covered."""
    cfg = load_controller_config(default_controller="JOINT_VELOCITY")
    assert cfg["kv"] == 4.0 and "kp" not in cfg
    if law == "kp":
        del cfg["kv"]
        cfg["kp"] = 3.0
    m, t = build_task("Lift", "Panda", cfg, ignore_done=True)
    rb = t["robot"][0]
    if law == "kv":
        assert np.allclose(rb["kp"], 4.0) and not np.any(rb["ki"]) and not np.any(rb["kd"])
    else:
        assert np.allclose(rb["kp"], 3.0 * (np.asarray(rb["torque_limit_hi"]) - np.asarray(rb["torque_limit_lo"]))) and np.allclose(rb["ki"], 0.005 * np.asarray(rb["kp"]))
    with pytest.raises(ValueError):
        build_task("Lift", "Panda", dict(cfg, kv=4.0, kp=3.0))
    orc, emu = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"]), EmuEnv(m, t, t["ncon_max"], t["nefc_max"])
    orc.reset(seed=9, env_id=2)
    emu.reset(seed=9, env_id=2)
    for k in range(4):
        a = orc.random_action(9, 2, k)
        o1, r1, _ = orc.step(a)
        o2, r2, _ = emu.step(a)
        # both sides run free (no re-synchronisation): the PID law's gains (3 x the actuator range = 522 N m s/rad on joints 1-4) amplify fp32 round-off of the velocities
        assert np.abs(o1 - o2).max() <= (2e-4 if law == "kv" else 5e-3) and abs(r1 - r2) <= 1e-5, (k, np.abs(o1 - o2).max())


NEW_FAMILIES = [("PickPlaceCan", ["Panda"]), ("PickPlaceMilk", ["Sawyer"]), ("PickPlaceCan", ["Sawyer"]), ("PickPlaceMilk", ["Panda"]), ("TwoArmPegInHole", ["Panda", "Panda"]),
                ("TwoArmPegInHole", ["Panda", "Sawyer"]), ("TwoArmPegInHole", ["Sawyer", "Sawyer"]), ("NutAssemblyRound", ["Panda"]), ("NutAssemblyRound", ["Sawyer"]),
                ("TwoArmHandoff", ["Panda", "Panda"]), ("TwoArmHandoff", ["Sawyer", "Sawyer"])]


@pytest.mark.parametrize("env_name,robots,ctrl", [(e, r, "OSC_POSE") for e, r in NEW_FAMILIES] + [("Lift", ["Panda"], "JOINT_POSITION"), ("Stack", ["Sawyer"], "JOINT_TORQUE")])
def test_emulator_twin_of_the_gpu_one_control_step_test(env_name, robots, ctrl):
    """tests/test_gpu_parity.py::test_other_config_families_one_control_step (which tests/test_gpu_zz_pickplace.py runs for the families added last) with the
    emulator standing where the CUDA library stands: same seed, same six environments 0 .. 5 control steps into their episodes, same assertions -- contact-pair
    lists bit-exact and torques 1e-5 relative after the first substep, 1e-4 on qpos / qvel / observation and 1e-5 on the reward after the control step."""
    m, t = build_task(env_name, robots, load_controller_config(default_controller=ctrl), ignore_done=True)
    nc, ne = t["ncon_max"], t["nefc_max"]
    for i in range(6):
        orc, emu = OracleEnv(m, t, ncon_max=nc, nefc_max=ne), EmuEnv(m, t, nc, ne, lanes=16 if m.nv <= 16 else 32)
        o = orc.reset(seed=83, env_id=i, episode=0)
        assert np.abs(o - emu.reset(seed=83, env_id=i, episode=0)).max() < 2e-6
        for k in range(i):
            orc.step(orc.random_action(83, i, k))
        qpos, qvel, warm, cs = orc.get_state()
        a = orc.random_action(83, i, i)
        emu.set_state(qpos, qvel, warm, cs, timestep=i, episode=1, bpose=orc.get_bpose())
        d = emu.debug_substep(a, True)
        orc.substep(a, True)
        assert orc.get("contact_geoms").reshape(-1, 2).astype(int).tolist() == d["contact_geoms"].tolist()
        tau = orc.get("torques")[:7 * len(robots)]
        assert np.abs(tau - d["torques"][:7 * len(robots)]).max() <= 1e-5 * max(1.0, np.abs(tau).max())
        orc.set_state(qpos, qvel, warm, cs); orc.set_timestep(i)
        emu.set_state(qpos, qvel, warm, cs, timestep=i, episode=1, bpose=orc.get_bpose())
        o1, r1, _ = orc.step(a)
        o2, r2, _ = emu.step(a)
        qp1, qv1, _, _ = orc.get_state()
        qp2, qv2, _, _ = emu.get_state()
        assert np.abs(qp1 - qp2).max() <= 1e-4 and np.abs(qv1 - qv2).max() <= 1e-4, (i, np.abs(qv1 - qv2).max())
        assert np.abs(o1 - o2).max() <= 1e-4 and abs(r1 - r2) <= 1e-5
    assert emu.counters() == (0, 0, 0)


def test_joint_position_law():
    """JOINT_POSITION (robosuite v1.0 JointPositionController, the last joint-space controller of SURVEY 8f-3): the goal is re-based on the current joint angles at
    every policy step (goal = q + 0.05 a), the torque is M_arm (kp (goal - q) - kd qd) + compensation with kp = 50, kd = 2 sqrt(kp).  Zero actions hold the pose;
    a constant action moves exactly the commanded joints, in the commanded direction, by less than 0.05 rad per control step; device code == oracle while running free."""
    cfg = load_controller_config(default_controller="JOINT_POSITION")
    assert (cfg["kp"], cfg["damping_ratio"], cfg["output_max"]) == (50, 1, 0.05)
    with pytest.raises(NotImplementedError):
        from robosuite_benchmark_b200.controllers import validate
        validate(dict(cfg, qpos_limits=[[-1] * 7, [1] * 7]))
    m, t = build_task("Lift", "Panda", cfg, ignore_done=True)
    rb = t["robot"][0]
    assert rb["ctrl_type"] == 4 and np.allclose(rb["kp"], 50.0) and np.allclose(rb["kd"], 2 * np.sqrt(50.0))
    orc, emu = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"]), EmuEnv(m, t, t["ncon_max"], t["nefc_max"], lanes=16)
    orc.reset(seed=3, env_id=0); emu.reset(seed=3, env_id=0)
    q0 = orc.get_state()[0][:7].copy()
    for _ in range(10):
        orc.step(np.zeros(8)); emu.step(np.zeros(8))
    assert np.abs(orc.get_state()[0][:7] - q0).max() < 1e-6
    a = np.zeros(8); a[1], a[3] = 1.0, -0.5
    for k in range(10):
        o1, r1, _ = orc.step(a)
        o2, r2, _ = emu.step(a)
        assert np.abs(o1 - o2).max() <= 2e-4 and abs(r1 - r2) <= 1e-5, (k, np.abs(o1 - o2).max())
    dq = orc.get_state()[0][:7] - q0
    assert 0.02 < dq[1] < 0.5 and -0.25 < dq[3] < -0.01 and abs(dq[3] / dq[1] + 0.5) < 0.05 and np.abs(dq[[0, 2, 4, 5, 6]]).max() < 2e-3, dq
