"""The reference's committed 2020 policies, rolled out in the fp64 CPU ORACLE (no GPU): the defaults that the transfer study established (COMPAT.md: OSC
orientation rule, JOINT_VELOCITY law) are what lets a policy trained against real robosuite + MuJoCo reach, grasp and lift here.  Guards those defaults on the
CPU; the same check on the CUDA path is tests/test_gpu_train_loop.py::test_committed_policy_transfers.

Fixtures: tests/golden/policy_*.npz = first/second-layer and head weights of `trainer/policy` from runs/<run>/.../params.pkl plus the run's logged evaluation
returns, exported by tools/eval_committed_policy.py (OSC) / tools/eval_committed_runs.py (JOINT_VELOCITY) in the build container."""
import os

import numpy as np
import pytest

from oracle.oracle import OracleEnv
from robosuite_benchmark_b200.controllers import load_controller_config
from robosuite_benchmark_b200.model.tasks import build_task
from robosuite_benchmark_b200.policy_io import DeterministicPolicy

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _returns(run, ctrl, overrides=None, episodes=4):
    d = dict(np.load(os.path.join(GOLDEN, f"policy_{run}.npz")))
    logged = d.pop("logged")
    d.pop("env_kwargs", None)
    pol = DeterministicPolicy({k: v.astype(np.float64) for k, v in d.items()})
    cfg = load_controller_config(default_controller=ctrl)
    for k, v in (overrides or {}).items():
        if k == "kp":
            cfg.pop("kv", None)
        cfg[k] = v
    m, t = build_task("Lift", "Panda", cfg, ignore_done=True)
    out = []
    for i in range(episodes):
        orc = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
        o, ret = orc.reset(seed=17, env_id=i), 0.0
        for _ in range(500):
            o, r, _ = orc.step(pol(np.asarray(o)))
            ret += r
        out.append(ret)
    return np.array(out), logged


def test_committed_osc_policy_lifts_under_the_shipped_orientation_rule():
    ret, logged = _returns("Lift-Panda-OSC-POSE-SEED17", "OSC_POSE")
    assert ret.mean() > 200.0 and ret.max() > 0.95 * logged.max(), (ret, logged.max())      # logged: 364 (last 50 epochs), best 486.4
    ret2, _ = _returns("Lift-Panda-OSC-POSE-SEED17", "OSC_POSE", {"orientation_delta": "axis_angle"})
    assert ret2.mean() < 60.0, ret2                                                        # robosuite >= 1.1's rule: the same policy never grasps


def test_committed_joint_velocity_policy_lifts_under_the_v10_law():
    ret, logged = _returns("Lift-Panda-JOINT-VELOCITY-SEED129", "JOINT_VELOCITY")
    assert ret.mean() > 120.0 and ret.max() > 350.0, (ret, logged[-50:].mean())             # logged: 341 (last 50 epochs)
    ret2, _ = _returns("Lift-Panda-JOINT-VELOCITY-SEED129", "JOINT_VELOCITY", {"kp": 3.0})
    assert ret2.mean() < 80.0, ret2                                                        # the PID law of robosuite >= 1.1
