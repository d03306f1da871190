"""TwoArmPegInHole (3 of the reference's committed run families: runs/TwoArmPegInHole-{PandaPanda,PandaSawyer,SawyerSawyer}-OSC-POSE-*): two gripper-less arms,
the peg rigid on robot 0's hand, the plate with the hole on robot 1's.  CPU checks: the fp64 oracle against what the reference's logs and committed policies pin,
and the device code (tests/emu) against the oracle.  On the CUDA kernels: tests/test_gpu_zz_pickplace.py.

What the reference pins (it ships no tests):
  * network sizes: observation 73 = 2 x 28 (robot-state without gripper) + 17 (object-state), action 12;
  * reward levels in the committed progress.csv files: maximum 0.98 = (1 + (1 - tanh 0.1) + 3) / 5, i.e. the hole's centre 0.1 from the plate's origin, and the
    epoch-0 evaluation level of a freshly initialised (near-zero-action) policy, per robot pair: 0.515 / 0.466 / 0.426 (mean of 5 seeds) -- a function of where the
    two hands, the peg and the plate ARE after reset, i.e. of the whole recalled layout;
  * the 15 committed policies: every one transfers (0.60 - 0.92 of its logged return; profiles/r2_policy_transfer_peginhole_cpu.txt)."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import OracleEnv
from robosuite_benchmark_b200.controllers import load_controller_config
from robosuite_benchmark_b200.model.tasks import build_task
from robosuite_benchmark_b200.policy_io import DeterministicPolicy
from tests.emu.emu import EmuEnv

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
#: evaluation/Rewards Mean at epoch 0, mean over the 5 committed seeds of each family (range over seeds +-0.02)
LOGGED_EPOCH0 = {("Panda", "Panda"): 0.5144, ("Panda", "Sawyer"): 0.4668, ("Sawyer", "Sawyer"): 0.4261}


@pytest.mark.parametrize("robots", list(LOGGED_EPOCH0))
def test_dims_and_reset_reward_level_match_the_logs(robots):
    m, t = build_task("TwoArmPegInHole", list(robots), load_controller_config(default_controller="OSC_POSE"), ignore_done=True, reward_shaping=True)
    assert (t["obs_dim"], t["act_dim"], t["task_id"], m.nv) == (73, 12, 5, 14)
    assert all(r["grip_ndof"] == 0 and r["grip_action_dim"] == 0 and r["left_finger_geoms"] == [] for r in t["robot"])
    orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
    level = []
    for ep in range(4):
        o = orc.reset(seed=3, env_id=0, episode=ep)
        tot = 0.0
        for _ in range(40):
            o, r, _ = orc.step(np.zeros(12))
            tot += r
        level.append(tot / 40)
        hole_pos, peg_to_hole, cos, tt, d = o[56:59], o[63:66], o[70], o[71], o[72]
        assert 0 <= cos <= 1 and d >= 0 and abs(np.linalg.norm(o[59:63]) - 1) < 1e-9 and abs(np.linalg.norm(o[66:70]) - 1) < 1e-9
        # the reward recomputed from the observation's own (cos, t, d) and peg-to-hole vector
        assert r == pytest.approx(((1 - np.tanh(np.linalg.norm(peg_to_hole))) + (1 - np.tanh(d)) + (1 - np.tanh(abs(tt))) + cos) / 5, abs=1e-12)
    assert np.mean(level) == pytest.approx(LOGGED_EPOCH0[robots], abs=0.025), (robots, level)


def test_success_reward_is_the_logged_maximum():
    """Peg through the hole's centre, aligned with its normal: (1 + (1 - tanh 0.1) + 1 + 1 + 1) / 5 = 0.98, the maximum every committed run logs."""
    m, t = build_task("TwoArmPegInHole", ["Panda", "Panda"], load_controller_config(default_controller="OSC_POSE"), ignore_done=True, reward_shaping=True)
    orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
    orc.reset(seed=0, env_id=0)
    from robosuite_benchmark_b200.model import mjcf
    # search the two arms' joint angles for the aligned pose by a few Gauss-Newton steps on the oracle's own observation (cos, t, d, peg-to-hole)
    qp, qv, w, cs = orc.get_state()
    rng = np.random.default_rng(0)

    def f(q):
        orc.set_state(q, np.zeros_like(qv), w, cs)
        o, r = orc.observe()
        return np.array([1 - o[70], o[71], o[72]]), r

    q = qp.copy()
    best = f(q)
    for it in range(400):
        J = np.zeros((3, 14))
        for j in range(14):
            dq = q.copy(); dq[j] += 1e-5
            J[:, j] = (f(dq)[0] - best[0]) / 1e-5
        step = -np.linalg.lstsq(J, best[0], rcond=1e-3)[0]
        step *= min(1.0, 0.2 / max(1e-9, np.abs(step).max()))
        q2 = q + step
        cand = f(q2)
        if np.linalg.norm(cand[0]) < np.linalg.norm(best[0]):
            q, best = q2, cand
        if np.linalg.norm(best[0]) < 1e-6:
            break
    assert best[1] == pytest.approx(0.98, abs=2e-3), best
    # sparse reward: 5 x success / 5
    m2, t2 = build_task("TwoArmPegInHole", ["Panda", "Panda"], load_controller_config(default_controller="OSC_POSE"), ignore_done=True, reward_shaping=False)
    o2 = OracleEnv(m2, t2, ncon_max=t2["ncon_max"], nefc_max=t2["nefc_max"])
    o2.reset(seed=0, env_id=0)
    assert o2.observe()[1] == 0.0
    o2.set_state(q, np.zeros_like(qv), w, cs)
    assert o2.observe()[1] == pytest.approx(1.0)


def test_committed_policy_inserts_the_peg_and_device_code_follows_it():
    """The committed TwoArmPegInHole-PandaSawyer-SEED251 policy (logged 478 over its last 50 epochs) aligns and inserts in the fp64 oracle; along its rollout the
    device code (emulator) makes the same control step from the same state every 25 steps: 1e-4 on qpos, 5e-4 on qvel / observation, 1e-5 on the reward."""
    d = dict(np.load(os.path.join(GOLDEN, "policy_TwoArmPegInHole-PandaSawyer-OSC-POSE-SEED251.npz")))
    logged, cfg = d.pop("logged"), json.loads(str(d.pop("env_kwargs")))
    pol = DeterministicPolicy({k: v.astype(np.float64) for k, v in d.items()})
    m, t = build_task(cfg["env_name"], cfg["robots"], load_controller_config(default_controller=cfg["controller"]), horizon=cfg["horizon"], ignore_done=True)
    nc, ne = t["ncon_max"], t["nefc_max"]
    orc, emu = OracleEnv(m, t, ncon_max=nc, nefc_max=ne), EmuEnv(m, t, nc, ne, lanes=16)          # nv = 14: the GPU runs this model as 16-lane groups
    o = orc.reset(seed=17, env_id=0)
    emu.reset(seed=17, env_id=0)
    ret, best, worst = 0.0, 0.0, 0.0
    for k in range(500):
        a = pol(np.asarray(o))
        if k % 25 == 12:
            qp, qv, w, cs = orc.get_state()
            emu.set_state(qp, qv, w, cs, timestep=k, bpose=orc.get_bpose())
            o2, r2, _ = emu.step(a)
        o, r, _ = orc.step(a)
        if k % 25 == 12:
            qp1, qv1, _, _ = orc.get_state()
            qp2, qv2, _, _ = emu.get_state()
            dq, dv = np.abs(qp1 - qp2).max(), max(np.abs(qv1 - qv2).max(), np.abs(o - o2).max())
            worst = max(worst, dv)
            assert dq <= 1e-4 and dv <= 5e-4 and abs(r - r2) <= 1e-5, (k, dq, dv, r, r2)
        ret += r; best = max(best, r)
    assert ret > 330.0 and best > 0.9 and ret < 1.02 * logged.max(), (ret, best, logged[-50:].mean())
    assert emu.counters() == (0, 0, 0)


def test_observation_slices_follow_the_gripper_dofs():
    """Host side of the obs-dict contract (environments._obs_slices, what GymWrapper(keys=...) and the single-env OrderedDict use): a gripper-less robot's
    robot-state block is 28 wide, not 32, and the blocks tile the row exactly for every implemented family."""
    from robosuite_benchmark_b200.environments import _obs_slices
    from robosuite_benchmark_b200.model.tasks import _BUILDERS
    two_arm = ("TwoArmLift", "TwoArmPegInHole", "TwoArmHandoff")
    for env_name in _BUILDERS:
        robots = ["Panda", "Sawyer"] if env_name in two_arm else ["Sawyer"]
        m, t = build_task(env_name, robots, load_controller_config(default_controller="OSC_POSE"))
        sl = _obs_slices(t)
        edges = [(v.start, v.stop) for v in sl.values()]
        assert edges[0][0] == 0 and edges[-1][1] == t["obs_dim"] and all(a[1] == b[0] for a, b in zip(edges, edges[1:])), (env_name, edges)
        assert list(sl)[-1] == "object-state" and len(sl) == len(robots) + 1
        width = 28 if env_name == "TwoArmPegInHole" else 32
        assert all(v.stop - v.start == width for k, v in sl.items() if k != "object-state")
