"""PickPlace in single-object mode (PickPlaceMilk / PickPlaceCan, 4 of the reference's committed run families: runs/PickPlace{Can,Milk}-{Panda,Sawyer}-OSC-POSE-*),
checked on the CPU: the fp64 oracle against what the reference's own logs and committed policies pin, and the device code (tests/emu: csrc/rsb_dev.h compiled
for the host) against the oracle on contact-rich states taken from a committed policy's rollout.  The same checks on the CUDA kernels: tests/test_gpu_zz_pickplace.py.

What the reference pins for this task (it ships no tests; robosuite itself is not installable here):
  * network sizes of the committed params.pkl: observation 46 = 32 (robot) + 14 (object-state), action 7;
  * the reward levels in the committed progress.csv files: plateaus at 0.35 (grasp) and 0.5 (lift), values up to 0.7 (hover), exactly 1.0 on success;
  * the committed policies themselves: policies whose runs log a working pick transfer to this restatement, policies whose runs log a failure fail here too
    (COMPAT.md; tools/eval_committed_runs_cpu.py PickPlace).
Fixtures: tests/golden/policy_PickPlace*.npz, exported by `tools/eval_committed_runs.py export` from runs/<run>/.../params.pkl."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import OracleEnv
from robosuite_benchmark_b200.controllers import load_controller_config
from robosuite_benchmark_b200.model import assets as A
from robosuite_benchmark_b200.model.tasks import build_task
from robosuite_benchmark_b200.policy_io import DeterministicPolicy
from tests.emu.emu import EmuEnv

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _build(env_name="PickPlaceCan", robot="Panda", **kw):
    kw.setdefault("reward_shaping", True)
    return build_task(env_name, [robot], load_controller_config(default_controller="OSC_POSE"), ignore_done=True, **kw)


def _policy(run):
    d = dict(np.load(os.path.join(GOLDEN, f"policy_{run}.npz")))
    logged, cfg = d.pop("logged"), json.loads(str(d.pop("env_kwargs")))
    return DeterministicPolicy({k: v.astype(np.float64) for k, v in d.items()}), logged, cfg


@pytest.mark.parametrize("env_name,robot", [("PickPlaceCan", "Panda"), ("PickPlaceCan", "Sawyer"), ("PickPlaceMilk", "Panda"), ("PickPlaceMilk", "Sawyer")])
def test_dims_reset_distribution_and_target_quadrant(env_name, robot):
    m, t = _build(env_name, robot)
    assert (t["obs_dim"], t["act_dim"], t["task_id"]) == (46, 7, 4)                         # the committed networks' sizes
    kind = env_name[len("PickPlace"):]
    tx, ty = t["task_par"][:2]
    b = A.PICK_OBJECTS[kind]["bin_id"]                                                      # Milk 0: low x / low y quadrant of bin 2, Can 3: high x / high y
    assert (tx < A.BIN2_POS[0]) == (b in (0, 2)) and (ty < A.BIN2_POS[1]) == (b < 2)
    orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
    half = np.array(A.PICK_OBJECTS[kind]["half"])
    xy = []
    for ep in range(40):
        o = orc.reset(seed=5, env_id=1, episode=ep)
        pos, quat, rel, relq = o[32:35], o[35:39], o[39:42], o[42:46]
        xy.append(pos[:2])
        assert abs(pos[2] - (0.82 + half[2] + 0.01)) < 1e-9 and abs(np.linalg.norm(quat) - 1) < 1e-9 and abs(quat[0]) + abs(quat[1]) < 1e-12    # upright, yaw only
        # object pose in the gripper frame: |rel| is the world distance, and rotating back by the eef orientation gives the world offset
        eef, q = o[21:24], o[24:28]
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(R @ rel - (pos - eef)).max() < 1e-9 and relq[3] >= 0 and abs(np.linalg.norm(relq) - 1) < 1e-9
    xy = np.array(xy) - np.array(A.BIN1_POS[:2])
    rad = np.hypot(half[0], half[1])
    assert np.abs(xy[:, 0]).max() <= A.BIN_SIZE[0] / 2 - rad - 0.05 + 1e-9 and np.abs(xy[:, 1]).max() <= A.BIN_SIZE[1] / 2 - rad - 0.05 + 1e-9
    assert np.abs(xy[:, 0]).max() > 0.5 * (A.BIN_SIZE[0] / 2 - rad - 0.05) and xy[:, 1].std() > 0.04                                             # and it does spread over the bin


def _reward_at(env, m, t, obj_pos, arm_q=None):
    """reward of the state (reset state with the object moved to obj_pos), evaluated by a zero-length observe: kinematics + contacts + task_reward."""
    qp, qv, w, cs = env["orc"].get_state()
    qa = t["obj_qposadr"][0]
    qp = qp.copy(); qp[qa:qa + 3] = obj_pos; qp[qa + 3:qa + 7] = [1, 0, 0, 0]
    qv = np.zeros_like(qv)
    env["orc"].set_state(qp, qv, w, cs)
    _, r = env["orc"].observe()
    return r


def test_reward_levels_match_the_logged_plateaus():
    """Success is exactly 1; without a grasp the reward is the reach term (<= 0.1) until the object hovers over its quadrant (0.5 .. 0.7); scaled by reward_scale."""
    m, t = _build("PickPlaceCan", "Panda")
    orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
    o = orc.reset(seed=1, env_id=0)
    env = dict(orc=orc)
    tx, ty, bz = t["task_par"][:3]
    half_z = A.PICK_OBJECTS["Can"]["half"][2]
    eef = o[21:24]
    # resting on the floor of its quadrant, gripper far away: success
    assert _reward_at(env, m, t, [tx, ty, 0.82 + half_z]) == pytest.approx(1.0, abs=1e-12)
    # same place but in the WRONG quadrant (the milk's): hover term of a not-above object, no lift -> 0.2 (1 - tanh(10 d)) with d ~ 0.3: tiny
    r = _reward_at(env, m, t, [tx - A.BIN_SIZE[0] / 2, ty - A.BIN_SIZE[1] / 2, 0.82 + half_z])
    assert 0 < r < 0.01
    # held high above the centre of the quadrant (no contact): 0.5 + 0.2 = 0.7
    assert _reward_at(env, m, t, [tx, ty, bz + 0.3]) == pytest.approx(0.7, abs=1e-9)
    # above the quadrant's edge region: between 0.5 and 0.7, decreasing with the distance
    r1, r2 = _reward_at(env, m, t, [tx + 0.03, ty, bz + 0.3]), _reward_at(env, m, t, [tx + 0.08, ty + 0.05, bz + 0.3])
    assert 0.5 < r2 < r1 < 0.7
    # in bin 1, 15 cm under the gripper (clear of the fingers): reach term 0.1 (1 - tanh(1.5)) (the hover term is ~1e-5 from there)
    assert _reward_at(env, m, t, [eef[0], eef[1], eef[2] - 0.15]) == pytest.approx(0.1 * (1 - np.tanh(1.5)), abs=1e-4)
    # between the open fingers: both pads touch -> grasp 0.35 and the lift term on top of it (0.35 .. 0.5)
    assert 0.35 <= _reward_at(env, m, t, [eef[0], eef[1], eef[2] - 0.05]) <= 0.5
    # sparse reward: nothing but success counts
    m2, t2 = _build("PickPlaceCan", "Panda", reward_shaping=False, reward_scale=3.0)
    orc2 = OracleEnv(m2, t2, ncon_max=t2["ncon_max"], nefc_max=t2["nefc_max"])
    orc2.reset(seed=1, env_id=0)
    assert _reward_at(dict(orc=orc2), m2, t2, [tx, ty, bz + 0.3]) == 0.0 and _reward_at(dict(orc=orc2), m2, t2, [tx, ty, 0.82 + half_z]) == pytest.approx(3.0)


@pytest.mark.parametrize("run,min_best,min_mean", [("PickPlaceCan-Sawyer-OSC-POSE-SEED59", 120.0, 40.0), ("PickPlaceMilk-Panda-OSC-POSE-SEED59", 120.0, 25.0)])
def test_committed_policy_picks_and_carries(run, min_best, min_mean):
    """A policy trained against real robosuite + MuJoCo grasps, lifts and carries the object towards its bin in the fp64 oracle: the grasp plateau (0.35) of its own training log is passed
    into the lift term (0.35 .. 0.5), and the best episode's return is of the size its run logged (Can-Sawyer-59: 74 over the last 50 epochs, best
    189; Milk-Panda-59: 98, best 218)."""
    pol, logged, cfg = _policy(run)
    m, t = build_task(cfg["env_name"], cfg["robots"], load_controller_config(default_controller=cfg["controller"]), horizon=cfg["horizon"], ignore_done=True)
    rets, best_r = [], 0.0
    for ep in range(6):
        orc = OracleEnv(m, t, ncon_max=48, nefc_max=160)
        o, ret = orc.reset(seed=17, env_id=ep), 0.0
        for _ in range(500):
            o, r, _ = orc.step(pol(np.asarray(o)))
            ret += r; best_r = max(best_r, r)
        rets.append(ret)
    rets = np.array(rets)
    assert best_r >= 0.40 and rets.max() > min_best and rets.mean() > min_mean, (rets, best_r, logged[-50:].mean(), logged.max())
    assert rets.max() < 1.3 * logged.max()


def test_device_code_matches_oracle_on_policy_states():
    """north_star: one control step from identical (qpos, qvel, action) on states sampled from rollouts of the committed variants.  The committed Can-Sawyer policy
    drives the oracle through reach -> grasp -> lift -> carry; every 20 control steps the device code (emulator) makes the same control step from the same state.
    Bound: 1e-4 on qpos, 5e-4 (median 1e-4) on qvel / observation, 1e-5 on the reward, as long as both sides work on the same contact set; a state whose contact
    set differs somewhere within the 25 substeps (a grazing contact that fp32 and fp64 switch at different substeps) is counted and must stay rare."""
    pol, _, cfg = _policy("PickPlaceCan-Sawyer-OSC-POSE-SEED59")
    m, t = build_task(cfg["env_name"], cfg["robots"], load_controller_config(default_controller=cfg["controller"]), horizon=cfg["horizon"], ignore_done=True)
    nc, ne = 32, 112
    orc, emu = OracleEnv(m, t, ncon_max=nc, nefc_max=ne), EmuEnv(m, t, nc, ne, lanes=16)          # nv = 15: the GPU runs this model as 16-lane groups
    o = orc.reset(seed=17, env_id=1)
    emu.reset(seed=17, env_id=1)
    checked = grazing = in_contact = 0
    dqs, dvs = [], []
    for k in range(300):
        a = pol(np.asarray(o))
        if k % 20 == 10:
            qp, qv, w, cs = orc.get_state()
            emu.set_state(qp, qv, w, cs, timestep=k, bpose=orc.get_bpose())
            # substep by substep on both sides: do they see the same contact pairs all along?
            same = True
            for sub in range(25):
                orc.substep(a, sub == 0)
                d = emu.debug_substep(a, sub == 0)
                same &= orc.get("contact_geoms").reshape(-1, 2).astype(int).tolist() == d["contact_geoms"].tolist()
            orc.set_state(qp, qv, w, cs)
            emu.set_state(qp, qv, w, cs, timestep=k, bpose=orc.get_bpose())
            o2, r2, _ = emu.step(a)
            o, r, _ = orc.step(a)
            qp1, qv1, _, _ = orc.get_state()
            qp2, qv2, _, _ = emu.get_state()
            dq, dv = np.abs(qp1 - qp2).max(), max(np.abs(qv1 - qv2).max(), np.abs(o - o2).max())
            checked += 1
            in_contact += int(orc.get("counts")[0] > 4)
            if same:
                dqs.append(dq); dvs.append(dv)
                assert dq <= 1e-4 and dv <= 5e-4 and abs(r - r2) <= 1e-5, (k, dq, dv, r, r2)
            else:
                grazing += 1
                assert dq <= 1e-3 and dv <= 0.2, (k, dq, dv)
        else:
            o, r, _ = orc.step(a)
    # same tolerances as the Lift policy-state test on the CUDA kernels (tests/test_gpu_parity.py): qpos 1e-4 everywhere, velocities 1e-4 typically and 5e-4 where
    # the gripper squeezes the object (fp32 resolution of large opposing contact forces)
    assert checked == 15 and in_contact >= 3 and grazing <= 3 and np.median(dvs) <= 1e-4, (checked, in_contact, grazing, np.sort(dvs))


#: the reference's 29 committed run families (runs/<family>-SEED*) with the observation / action sizes of their committed networks
RUN_FAMILIES = {
    **{f"{e}-{r}-{c}": (o, 7 if c == "OSC-POSE" else 8) for e, o in (("Lift", 42), ("Door", 46), ("Stack", 55)) for r in ("Panda", "Sawyer") for c in ("OSC-POSE", "JOINT-VELOCITY")},
    **{f"{e}-{r}-OSC-POSE": (46, 7) for e in ("PickPlaceCan", "PickPlaceMilk", "NutAssemblyRound") for r in ("Panda", "Sawyer")},
    **{f"TwoArmLift-{r}-OSC-POSE": (89, 14) for r in ("PandaPanda", "SawyerSawyer")},
    **{f"TwoArmHandoff-{r}-OSC-POSE": (86, 14) for r in ("PandaPanda", "SawyerSawyer")},
    **{f"TwoArmPegInHole-{r}-OSC-POSE": (73, 12) for r in ("PandaPanda", "PandaSawyer", "SawyerSawyer")},
    **{f"Wipe-{r}-{c}": (379, 6 if c == "OSC-POSE" else 7) for r in ("Panda", "Sawyer") for c in ("OSC-POSE", "JOINT-VELOCITY")},
}


def test_every_committed_run_family_builds_with_its_networks_dims_or_refuses_loudly():
    """SURVEY 8f-3 coverage, family by family: 25 of the 29 committed run families compile into a task whose observation / action sizes are those of the family's
    committed policy and Q networks; the other four (Wipe) raise NotImplementedError -- never a silent stand-in."""
    assert len(RUN_FAMILIES) == 29
    built = 0
    for fam, (obs_dim, act_dim) in RUN_FAMILIES.items():
        env_name, robots, ctrl = fam.split("-", 2)
        robots = {"PandaPanda": ["Panda", "Panda"], "SawyerSawyer": ["Sawyer", "Sawyer"], "PandaSawyer": ["Panda", "Sawyer"]}.get(robots, [robots])
        cfg = load_controller_config(default_controller=ctrl.replace("-", "_"))
        if env_name == "Wipe":
            with pytest.raises(NotImplementedError):
                build_task(env_name, robots, cfg)
            continue
        m, t = build_task(env_name, robots, cfg, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
        assert (t["obs_dim"], t["act_dim"]) == (obs_dim, act_dim), (fam, t["obs_dim"], t["act_dim"])
        assert t["substeps"] == 25 and m.nv <= 32
        built += 1
    assert built == 25
