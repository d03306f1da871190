"""NutAssemblyRound (2 of the reference's committed run families: runs/NutAssemblyRound-{Panda,Sawyer}-OSC-POSE-*), CPU checks: the fp64 oracle against what the
reference's logs and committed policies pin, the device code (tests/emu, 16-lane groups as on the GPU) against the oracle.  CUDA kernels: tests/test_gpu_zz_pickplace.py.

What the reference pins (it ships no tests):
  * network sizes: observation 46 = 32 + 14 (nut pos, quat, pose in the gripper frame), action 7;
  * reward plateaus in the committed progress.csv files: 0.35 (grasp), 0.5 (lift), up to 0.7 (hover), exactly 1.0 on success;
  * the epoch-0 evaluation reward of the Sawyer runs, 0.00035 - 0.00039 per step with minimum 0.00027 - 0.00034: the Sawyer's hand starts too far from the nut for the
    reach term, so this is the HOVER term 0.2 (1 - tanh(10 |nut - peg|)) of an untouched nut -- it pins the distance between the nut's start range and its peg;
  * the committed policies: partial transfer (Sawyer 0.54, Panda 0.25 of the logged returns; profiles/r2_policy_transfer_nutassembly_cpu.txt) -- the nut's collision
    geometry is authored, not recalled (model/assets.py ROUND_NUT), and the sweep over it in that file moves these ratios little."""
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


def _build(robot="Sawyer", **kw):
    kw.setdefault("reward_shaping", True)
    return build_task("NutAssemblyRound", [robot], load_controller_config(default_controller="OSC_POSE"), ignore_done=True, **kw)


def test_dims_and_untouched_nut_reward_level_match_the_sawyer_logs():
    m, t = _build("Sawyer")
    assert (t["obs_dim"], t["act_dim"], t["task_id"]) == (46, 7, 6)
    orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
    levels = []
    for ep in range(12):
        o = orc.reset(seed=2, env_id=0, episode=ep)
        assert -0.115 <= o[32] <= -0.11 and -0.225 <= o[33] <= -0.11                           # NUT_PLACE["Round"]
        for _ in range(10):
            o, r, _ = orc.step(np.zeros(7))
        px, py = t["task_par"][:2]
        assert r == pytest.approx(0.2 * (1 - np.tanh(10 * np.hypot(o[32] - px, o[33] - py))), rel=2e-2)     # the hover term; the reach term is ~1e-5 from where the Sawyer starts
        levels.append(r)
    # logged, 5 Sawyer seeds x 5 evaluation episodes at epoch 0: mean 0.00035 .. 0.00039, min 0.00027 .. 0.00034, max 0.00037 .. 0.00064
    assert 0.00033 <= np.mean(levels) <= 0.00041 and min(levels) >= 0.00026 and max(levels) <= 0.00045, levels


def test_reward_plateaus():
    m, t = _build("Panda")
    orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
    o = orc.reset(seed=1, env_id=0)
    qp, qv, w, cs = orc.get_state()
    qa = t["obj_qposadr"][0]
    px, py, tz, zt = t["task_par"][:4]

    def rew(pos, yaw=0.0):
        q = qp.copy(); q[qa:qa + 3] = pos; q[qa + 3:qa + 7] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
        orc.set_state(q, np.zeros_like(qv), w, cs)
        return orc.observe()[1]

    hh = A.ROUND_NUT["half_h"]
    assert rew([px, py, tz + hh]) == pytest.approx(1.0)                                        # around the peg, on the table, gripper away: success
    assert rew([px, py, tz + 0.06]) == pytest.approx(0.2, abs=1e-3)                            # over the peg but above table + 0.05: hover term at distance 0, no lift
    assert rew([px + 0.05, py, tz + hh]) == pytest.approx(0.2 * (1 - np.tanh(0.5)), abs=1e-3)   # beside the peg
    # the reach term aims at the HANDLE geom (the nut's last geom), not at the nut's centre: handle 10 cm straight under the gripper site
    eef = o[21:24]
    hx = (A.ROUND_NUT["ring_in"] + 2 * A.ROUND_NUT["ring_t"] + A.ROUND_NUT["handle_out"]) / 2           # handle geom centre in the nut frame
    for yaw in (0.0, 1.3):
        r = rew([eef[0] - hx * np.cos(yaw), eef[1] - hx * np.sin(yaw), eef[2] - 0.10], yaw)
        assert r == pytest.approx(0.1 * (1 - np.tanh(1.0)), abs=2e-4), (yaw, r)
    m2, t2 = _build("Panda", reward_shaping=False)
    o2 = OracleEnv(m2, t2, ncon_max=t2["ncon_max"], nefc_max=t2["nefc_max"])
    o2.reset(seed=1, env_id=0)
    assert o2.observe()[1] == 0.0


def test_committed_policy_and_device_parity_along_its_rollout():
    """The committed NutAssemblyRound-Sawyer-SEED251 policy (logged 111 over its last 50 epochs) reaches and grasps the nut's handle in the fp64 oracle; along its
    rollout the device code (emulator, 16-lane groups) makes the same control step from the same state every 20 steps.  With 36+ contacts of a nine-box nut on the
    table the bound on velocities is the squeezed-contact one (5e-4, median 1e-4; tests/test_gpu_parity.py), and contact-set switches are counted."""
    d = dict(np.load(os.path.join(GOLDEN, "policy_NutAssemblyRound-Sawyer-OSC-POSE-SEED251.npz")))
    logged, cfg = d.pop("logged"), json.loads(str(d.pop("env_kwargs")))
    pol = DeterministicPolicy({k: v.astype(np.float64) for k, v in d.items()})
    m, t = build_task(cfg["env_name"], cfg["robots"], load_controller_config(default_controller=cfg["controller"]), horizon=cfg["horizon"], ignore_done=True)
    nc, ne = t["ncon_max"], t["nefc_max"]
    best_ret, best_r = 0.0, 0.0
    dvs, switches, checked = [], 0, 0
    for ep in (0, 1, 2):
        orc, emu = OracleEnv(m, t, ncon_max=nc, nefc_max=ne), EmuEnv(m, t, nc, ne, lanes=16)
        o = orc.reset(seed=17, env_id=ep)
        emu.reset(seed=17, env_id=ep)
        ret = 0.0
        for k in range(500):
            a = pol(np.asarray(o))
            if ep == 0 and k % 20 == 10 and k < 300:
                qp, qv, w, cs = orc.get_state()
                emu.set_state(qp, qv, w, cs, timestep=k, bpose=orc.get_bpose())
                same = True
                for sub in range(25):
                    orc.substep(a, sub == 0)
                    dbg = emu.debug_substep(a, sub == 0)
                    same &= orc.get("contact_geoms").reshape(-1, 2).astype(int).tolist() == dbg["contact_geoms"].tolist()
                orc.set_state(qp, qv, w, cs)
                emu.set_state(qp, qv, w, cs, timestep=k, bpose=orc.get_bpose())
                o2, r2, _ = emu.step(a)
            o, r, _ = orc.step(a)
            if ep == 0 and k % 20 == 10 and k < 300:
                qp1, qv1, _, _ = orc.get_state()
                qp2, qv2, _, _ = emu.get_state()
                dq, dv = np.abs(qp1 - qp2).max(), max(np.abs(qv1 - qv2).max(), np.abs(o - o2).max())
                checked += 1
                if same:
                    dvs.append(dv)
                    assert dq <= 1e-4 and dv <= 5e-4 and abs(r - r2) <= 1e-5, (k, dq, dv, r, r2)
                else:
                    switches += 1
                    assert dq <= 1e-3 and dv <= 0.2, (k, dq, dv)
            ret += r; best_r = max(best_r, r)
        best_ret = max(best_ret, ret)
    assert checked == 15 and switches <= 4 and np.median(dvs) <= 1e-4, (checked, switches, np.sort(dvs))
    assert best_r >= 0.35 and best_ret > 40.0 and best_ret < 1.3 * logged.max(), (best_ret, best_r)
