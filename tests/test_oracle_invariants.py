"""The oracle has no MuJoCo to be pinned against (PARITY UNPINNED, see oracle/rsb_oracle.c), so it is
validated by physics invariants."""
import numpy as np
import pytest

from oracle.oracle import OracleEnv, box_box
from robosuite_benchmark_b200.model.mjcf import compile_mjcf, dense_mass_matrix, kinematics

CHAIN = """<mujoco><compiler angle="radian"/><option timestep="0.0005" gravity="0 0 -9.81"/>
<worldbody>
 <body name="l1" pos="0 0 1"><inertial pos="0.1 0.02 -0.2" mass="1.3" diaginertia="0.02 0.03 0.01"/>
  <joint name="j1" type="hinge" axis="0 1 0"/>
  <body name="l2" pos="0.05 0 -0.4" quat="0.9238795 0.3826834 0 0"><inertial pos="0 0.1 -0.15" quat="0.9 0.1 0.3 0.2" mass="0.7" diaginertia="0.01 0.004 0.008"/>
   <joint name="j2" type="hinge" axis="1 0 0" pos="0 0.02 0"/>
   <body name="l3" pos="0 0 -0.3"><inertial pos="0 0 -0.1" mass="0.4" diaginertia="0.002 0.002 0.001"/>
    <joint name="j3" type="slide" axis="0 0.6 0.8"/>
    <joint name="j4" type="hinge" axis="0 0 1"/>
   </body></body></body>
 <body name="ball" pos="1 0 1"><freejoint name="fj"/><inertial pos="0.01 0.02 0.03" quat="0.8 0.2 0.5 0.1" mass="0.9" diaginertia="0.01 0.02 0.03"/></body>
</worldbody></mujoco>"""


def _task_stub():
    from robosuite_benchmark_b200.model.tasks import empty_task
    return empty_task()


def _energy(m, env):
    qpos, qvel, _, _ = env.get_state()
    M = env.get("M", (m.nv, m.nv))
    xipos = env.get("xipos", (m.nbody, 3))
    return 0.5 * qvel @ M @ qvel + 9.81 * float(np.sum(m.body_mass * xipos[:, 2]))


def test_mass_matrix_matches_jacobian_sum():
    m = compile_mjcf(CHAIN)
    env = OracleEnv(m, _task_stub())
    rng = np.random.default_rng(0)
    for _ in range(5):
        q = m.qpos0 + rng.normal(size=m.nq) * 0.7
        q[-4:] /= np.linalg.norm(q[-4:])
        env.set_state(q, np.zeros(m.nv))
        env.forward()
        M = env.get("M", (m.nv, m.nv))
        assert np.allclose(M, dense_mass_matrix(m, q), atol=1e-12)
        assert np.linalg.eigvalsh(M).min() > 0


def test_energy_conserved_without_damping_or_contacts():
    """RNE bias + CRB + Euler: total energy drift of an unactuated, undamped system is O(h)."""
    m = compile_mjcf(CHAIN)
    env = OracleEnv(m, _task_stub())
    rng = np.random.default_rng(1)
    q = m.qpos0.copy()
    q[:4] = [0.5, -0.8, 0.05, 1.0]
    v = rng.normal(size=m.nv) * 1.5
    v[4:7] = [0.2, -0.1, 0.5]
    env.set_state(q, v)
    env.forward()
    e0 = _energy(m, env)
    for _ in range(2000):          # 1 s
        env.forward()
        env.set_ctrl(np.zeros(0))
        env.fwd_actuation()
        env.fwd_constraint()
        env.euler()
    env.forward()
    e1 = _energy(m, env)
    assert abs(e1 - e0) < 2e-2 * max(1.0, abs(e0)), (e0, e1)


def test_free_body_angular_momentum():
    """Torque-free tumbling: world-frame angular momentum about the COM is conserved (gyroscopic bias terms)."""
    xml = CHAIN.replace('gravity="0 0 -9.81"', 'gravity="0 0 0"')
    m = compile_mjcf(xml)
    env = OracleEnv(m, _task_stub())
    q = m.qpos0.copy()
    v = np.zeros(m.nv)
    v[7:10] = [3.0, -2.0, 1.0]

    def angmom():
        qpos, qvel, _, _ = env.get_state()
        xq = env.get("xquat", (m.nbody, 4))[m.id("body", "ball")]
        from robosuite_benchmark_b200.model.mjcf import quat2mat
        R = quat2mat(xq)
        Ri = R @ quat2mat(m.body_iquat[m.id("body", "ball")])
        Iw = Ri @ np.diag(m.body_inertia[m.id("body", "ball")]) @ Ri.T
        # note: body frame origin != COM; angular velocity is what matters for spin momentum
        return Iw @ (R @ qvel[7:10])

    env.set_state(q, v)
    env.forward()
    L0 = angmom()
    for _ in range(2000):
        env.forward(); env.fwd_actuation(); env.fwd_constraint(); env.euler()
    env.forward()
    L1 = angmom()
    assert np.linalg.norm(L1 - L0) < 2e-2 * np.linalg.norm(L0)


def test_newton_residual_and_bias_consistency(lift_panda_osc):
    """qacc_smooth solves M a = passive - bias + actuator; solver result minimises the cost (gradient ~ 0,
    no lower cost along random directions) and contact forces satisfy the elliptic cone."""
    m, task = lift_panda_osc
    env = OracleEnv(m, task)
    env.reset(seed=3, env_id=5)
    rng = np.random.default_rng(2)
    for _ in range(4):
        env.step(rng.uniform(-1, 1, size=7))
    env.forward()
    env.fwd_actuation()
    M = env.get("M", (m.nv, m.nv))
    rhs = env.get("qfrc_passive") - env.get("qfrc_bias") + env.get("qfrc_actuator")
    assert np.allclose(M @ env.get("qacc_smooth"), rhs, atol=1e-9)
    env.fwd_constraint()
    qacc = env.get("qacc")
    ncon, nefc = int(env.get("counts")[0]), int(env.get("counts")[1])
    assert ncon >= 4 and nefc >= 12            # cube resting on the table
    c0 = env.cost(qacc)
    for _ in range(50):
        d = rng.normal(size=m.nv) * 1e-3
        assert env.cost(qacc + d) >= c0 - 1e-12
    # stationarity: M (a - a_s) = J^T f
    J = env.get("efc_J", (nefc, m.nv))
    f = env.get("efc_force")
    assert np.allclose(M @ (qacc - env.get("qacc_smooth")), J.T @ f, atol=1e-7)
    assert np.allclose(env.get("qfrc_constraint"), J.T @ f, atol=1e-10)
    # friction cone (regularised): |f_t| <= mu * f_n with mu from contact friction
    typ = env.get("efc_type")
    i = 0
    while i < nefc:
        if typ[i] == 2:
            assert f[i] >= -1e-12
        i += 1


def test_cube_rests_on_table(lift_panda_osc):
    m, task = lift_panda_osc
    env = OracleEnv(m, task)
    env.reset(seed=0, env_id=0)
    for _ in range(40):
        obs, r, d = env.step(np.zeros(7))
    cube_z = obs[34]
    assert abs(cube_z - (0.8 + 0.021)) < 1e-3
    qpos, qvel, _, _ = env.get_state()
    assert np.abs(qvel).max() < 1e-3           # arm holds pose (gravity compensation + OSC), cube at rest
    # zero-action reward per step: the reference logged 6.42/500 = 0.01285 for this family
    assert 0.008 < r < 0.02


def test_box_box_face_and_edge():
    I = np.eye(3)
    # small cube resting 1 mm inside a big slab: 4 corner contacts, normal +z (A -> B), dist -1 mm
    c = box_box([0, 0, 0], I, [0.4, 0.4, 0.025], [0.1, 0.05, 0.025 + 0.02 - 0.001], I, [0.02, 0.02, 0.02])
    assert c.shape[0] == 4
    assert np.allclose(c[:, 3:6], [0, 0, 1]) and np.allclose(c[:, 6], -0.001)
    assert np.allclose(sorted(c[:, 0]), sorted([0.08, 0.08, 0.12, 0.12]))
    # separated
    assert box_box([0, 0, 0], I, [0.1] * 3, [0.3, 0, 0], I, [0.1] * 3).shape[0] == 0
    # edge-edge: two boxes rotated 45 deg about different axes
    from robosuite_benchmark_b200.model.mjcf import quat2mat, axisangle2quat
    Ra = quat2mat(axisangle2quat([0, 0, 1], np.pi / 4))
    Rb = quat2mat(axisangle2quat([0, 1, 0], np.pi / 4))
    d = 0.1 * np.sqrt(2)
    c = box_box([0, 0, 0], Ra, [0.1] * 3, [2 * d - 0.005, 0, 0], Rb, [0.1] * 3)
    assert c.shape[0] == 1 and abs(c[0, 6] + 0.005) < 1e-9 and np.allclose(c[0, 3:6], [1, 0, 0], atol=1e-9)


@pytest.mark.parametrize("switch,value,env,robots", [("DOOR_STYLE", "round1", "Door", ["Panda"]), ("DOOR_STYLE", "robosuite_recalled", "Door", ["Sawyer"]),
                                                     ("RETHINK_FINGER_STYLE", "narrow_tall", "Lift", ["Sawyer"]),
                                                     ("POT", dict(thickness=0.005, handle_z=0.05, bar_half=0.045, side_bars=False), "TwoArmLift", ["Panda", "Panda"])])
def test_asset_switches_build_and_step(switch, value, env, robots):
    """model/assets.py keeps the round-1 stand-ins selectable next to the variants the policy-transfer study chose (COMPAT.md): every switch must still compile
    and step without NaN, the settled scene must stay settled (no initial penetration kicking things around)."""
    from robosuite_benchmark_b200.model import assets as A
    from robosuite_benchmark_b200.controllers import load_controller_config
    from robosuite_benchmark_b200.model.tasks import build_task
    from oracle.oracle import OracleEnv
    old = getattr(A, switch)
    try:
        setattr(A, switch, dict(old, **value) if isinstance(value, dict) else value)
        m, t = build_task(env, robots, load_controller_config(default_controller="OSC_POSE"), ignore_done=True)
        orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
        o = orc.reset(seed=3, env_id=0)
        for _ in range(5):
            o, r, _ = orc.step(np.zeros(t["act_dim"]))
        qv = orc.get_state()[1]
        assert np.isfinite(o).all() and np.isfinite(r) and np.abs(qv).max() < 0.5, np.abs(qv).max()
    finally:
        setattr(A, switch, old)
