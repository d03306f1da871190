"""TwoArmHandoff (2 of the reference's committed run families: runs/TwoArmHandoff-{PandaPanda,SawyerSawyer}-OSC-POSE-*), CPU checks.

What the reference pins (it ships no tests):
  * network sizes: observation 86 = 2 x 32 + 22 (hammer pos, quat, handle pos, both eef positions, handle - eef for both arms), action 14;
  * the stage values of the reward, read off the committed progress.csv files (per-epoch maxima over 20 000 rows): reach term never above 0.125, a plateau at exactly
    0.25 (arm 0 holds the hammer), 0.5 .. 0.625 (lifted, arm 1 approaching), exactly 1.0 (handed over);
  * the epoch-0 evaluation reward level, i.e. 0.125 (1 - tanh |gripper 0 - handle|) after reset: 0.0974 (PandaPanda), 0.0917 (SawyerSawyer; mean of the seeds whose
    epoch-0 episodes stayed in the reach stage) -- it pins where the arms, the table and the hammer stand.
Transfer of the committed policies is PARTIAL (profiles/r2_policy_transfer_handoff_cpu.txt): with the hammer lying along y the better policies grasp, lift and hold it
for whole episodes, the others stay at the reach level; the hammer is generated with random sizes upstream and is authored from the recalled means here
(model/assets.py HANDOFF)."""
import numpy as np
import pytest

from oracle.oracle import OracleEnv
from robosuite_benchmark_b200.controllers import load_controller_config
from robosuite_benchmark_b200.model.tasks import build_task

LOGGED_EPOCH0 = {("Panda", "Panda"): 0.0974, ("Sawyer", "Sawyer"): 0.0917}


@pytest.mark.parametrize("robots", list(LOGGED_EPOCH0))
def test_dims_and_reset_reward_level_match_the_logs(robots):
    m, t = build_task("TwoArmHandoff", list(robots), load_controller_config(default_controller="OSC_POSE"), ignore_done=True, reward_shaping=True)
    assert (t["obs_dim"], t["act_dim"], t["task_id"], m.nv) == (86, 14, 7, 24)
    orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
    level = []
    for ep in range(10):
        o = orc.reset(seed=4, env_id=0, episode=ep)
        for _ in range(30):
            o, r, _ = orc.step(np.zeros(14))
        hammer, handle, e0, e1, g0, g1 = o[64:67], o[71:74], o[74:77], o[77:80], o[80:83], o[83:86]
        assert np.abs(g0 - (handle - e0)).max() < 1e-12 and np.abs(g1 - (handle - e1)).max() < 1e-12 and np.abs(e0 - o[21:24]).max() < 1e-12 and np.abs(e1 - o[53:56]).max() < 1e-12
        assert r == pytest.approx(0.125 * (1 - np.tanh(np.linalg.norm(g0))), abs=1e-12)         # reach stage: nobody touches the hammer
        assert abs(hammer[2] - (0.8 + 0.0175)) < 0.03 and -0.13 <= hammer[0] <= 0.13 and -0.52 <= hammer[1] <= -0.38       # resting on the narrow table beside robot 0 (propped up a little by its head)
        level.append(r)
    assert np.mean(level) == pytest.approx(LOGGED_EPOCH0[robots], abs=0.003), (robots, level)


def test_stage_values_are_the_logged_ones():
    m, t = build_task("TwoArmHandoff", ["Panda", "Panda"], load_controller_config(default_controller="OSC_POSE"), ignore_done=True, reward_shaping=True)
    orc = OracleEnv(m, t, ncon_max=48, nefc_max=160)
    o = orc.reset(seed=1, env_id=0)
    qp, qv, w, cs = orc.get_state()
    qa = t["obj_qposadr"][0]
    c = np.sqrt(0.5)

    def rew(pos, fingers0=None):
        q = qp.copy(); q[qa:qa + 3] = pos; q[qa + 3:qa + 7] = [c, 0, c, 0]                 # handle along x: across robot 0's closed fingers
        if fingers0 is not None:
            q[t["robot"][0]["grip_qposadr"][0]], q[t["robot"][0]["grip_qposadr"][1]] = fingers0, -fingers0
        orc.set_state(q, np.zeros_like(qv), w, cs)
        return orc.observe()

    e0, e1 = o[74:77], o[77:80]
    # free in the air, high above the table, far from both grippers: lifted stage, 0.5 + 0.125 (1 - tanh |gripper 1 - handle|)
    ob, r = rew([0.0, 0.0, 1.4])
    assert r == pytest.approx(0.5 + 0.125 * (1 - np.tanh(np.linalg.norm(ob[83:86]))), abs=1e-12) and 0.5 < r < 0.625
    # on the table: reach stage, at most 0.125
    ob, r = rew([0.0, -0.45, 0.8 + 0.0175])
    assert r == pytest.approx(0.125 * (1 - np.tanh(np.linalg.norm(ob[80:83]))), abs=1e-12) and r < 0.125
    # handle between robot 0's fingers, closed onto it, at the gripper's height (not lifted: the table is 0.2 below, the threshold is table + 0.1 ... so lower the
    # whole comparison: put it just under the lift threshold by using robot 0's own height only if that is below it; otherwise check the lifted-and-held value)
    ob, r = rew([e0[0], e0[1], e0[2]], fingers0=0.0176)
    lifted = e0[2] - 0.0175 - 0.8 > 0.1
    assert r == pytest.approx((1.0 + 0.25 * (1 - np.tanh(np.linalg.norm(ob[83:86])))) / 2 if lifted else 0.25, abs=1e-9), r
    # sparse reward: nothing until the hammer is handed over
    m2, t2 = build_task("TwoArmHandoff", ["Panda", "Panda"], load_controller_config(default_controller="OSC_POSE"), ignore_done=True, reward_shaping=False)
    o2 = OracleEnv(m2, t2, ncon_max=48, nefc_max=160)
    o2.reset(seed=1, env_id=0)
    assert o2.observe()[1] == 0.0
