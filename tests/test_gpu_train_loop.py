"""End-to-end: the reference's `--variant` JSON drives the batched training loop; outputs carry the reference's schema."""
import csv
import glob
import json
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _variant(tmp, n_epochs=2):
    v = json.load(open(os.path.join(GOLDEN, "variant_Lift-Panda-OSC-POSE-SEED17.json")))      # verbatim copy of the committed run's variant.json
    v["algorithm_kwargs"].update(num_epochs=n_epochs, num_eval_steps_per_epoch=40, num_expl_steps_per_train_loop=40,
                                 num_trains_per_train_loop=12, min_num_steps_before_training=80, expl_max_path_length=20, eval_max_path_length=20)
    v["replay_buffer_size"] = 4096
    p = os.path.join(tmp, "variant.json")
    json.dump(v, open(p, "w"))
    return p


@pytest.mark.parametrize("num_envs", [1, 2])
def test_variant_json_trains_and_logs_reference_schema(tmp_path, num_envs):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from robosuite_benchmark_b200.train import build_parser, run_experiment
    args = build_parser().parse_args(["--variant", _variant(str(tmp_path)), "--seed", "17", "--log_dir", str(tmp_path / "log"), "--num_envs", str(num_envs)])
    algo, log_dir = run_experiment(args)
    assert os.path.basename(os.path.dirname(log_dir)) == "Lift_Panda_OSC_POSE_SEED17"
    rows = list(csv.DictReader(open(os.path.join(log_dir, "progress.csv"))))
    cols = json.load(open(os.path.join(GOLDEN, "progress_columns.json")))
    assert list(rows[0].keys()) == cols                      # the reference's 83 columns, same order
    assert len(rows) == 2 and rows[1]["Epoch"] == "1"
    r0 = rows[0]
    assert float(r0["replay_buffer/size"]) == 120.0           # 80 warm-up + 40 exploration transitions
    assert float(r0["exploration/num paths total"]) == 6.0 and float(r0["evaluation/num paths total"]) == 2.0
    assert np.float32(float(r0["trainer/Alpha"])) == np.float32(0.9990004897117615) and float(r0["trainer/Alpha Loss"]) == 0.0
    assert abs(float(r0["trainer/Log Pis Mean"]) + 0.67 * 7) < 0.6
    from robosuite_benchmark_b200.sac import register_safe_globals
    register_safe_globals()
    snap = torch.load(os.path.join(log_dir, "params.pkl"))        # the reference's own way of reading it (util/rlkit_utils.py:173-174)
    assert sorted(snap) == sorted(["trainer/policy", "trainer/qf1", "trainer/qf2", "trainer/target_qf1", "trainer/target_qf2",
                                   "exploration/policy", "evaluation/policy"])
    assert tuple(snap["trainer/policy"].state_dict()["fc0.weight"].shape) == (256, 42) and tuple(snap["trainer/qf1"].state_dict()["fc0.weight"].shape) == (256, 49)
    a, info = snap["evaluation/policy"].get_action(np.zeros(42))          # ... and of using it: a policy object with get_action
    assert a.shape == (7,) and np.all(np.abs(a) <= 1.0) and info == {}
    live = algo.trainer.policy.get_actions(torch.zeros(1, 42, device="cuda:0"), deterministic=True)[0].cpu().numpy()
    assert np.abs(a - live).max() < 1e-6
    assert json.load(open(os.path.join(log_dir, "variant.json")))["trainer_kwargs"]["qf_lr"] == 0.0005


@pytest.mark.gpu
def test_rollout_cli_on_a_committed_variant(tmp_path):
    """scripts/rollout.py on the batched backend: a run directory with the reference's variant.json and a snapshot written by this
    package's own _get_snapshot layout (torch-pickled network objects) -> deterministic-policy returns; the untrained policy scores the logged
    epoch-0 level (Lift-Panda-OSC_POSE zero/untrained-policy return ~3-20 over 500 steps)."""
    import json, os, pickle, shutil
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from robosuite_benchmark_b200.rollout import main
    from robosuite_benchmark_b200.sac import MakeDeterministic, ParamStore, TanhGaussianPolicy
    here = os.path.dirname(__file__)
    shutil.copy(os.path.join(here, "golden", "variant_Lift-Panda-OSC-POSE-SEED17.json"), tmp_path / "variant.json")
    store = ParamStore(42, 7, torch.device("cuda", 0), seed=17)
    torch.save({"evaluation/policy": MakeDeterministic(TanhGaussianPolicy.of(store))}, tmp_path / "params.pkl")
    rets = main(["--load_dir", str(tmp_path), "--num_episodes", "4", "--horizon", "100"])
    assert rets.shape == (4,) and np.isfinite(rets).all() and 0.0 < rets.mean() < 20.0


@pytest.mark.gpu
def test_committed_policy_transfers():
    """The reference's committed Lift-Panda-OSC_POSE-SEED17 policy (trained in 2020 against real robosuite + MuJoCo; exported to
    tests/golden/policy_*.npz by tools/eval_committed_policy.py) is rolled out deterministically in the CUDA env for full 500-step episodes.
    Its logged evaluation return is 364 (last 50 epochs), maximum 486.4.  Here: 328 +- 8 over 512 episodes, maximum 487.6, the cube is lifted in 74 % of the
    episodes (profiles/r2_policy_transfer.txt) -- with the axis-angle orientation convention the same policy scores 21.  This is the one check in this
    repository that ties the physics + controller + observation layout to the real simulator the policy was trained in."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import robosuite_benchmark_b200 as suite
    from robosuite_benchmark_b200.rollout import policy_from_state_dict
    d = dict(np.load(os.path.join(GOLDEN, "policy_Lift-Panda-OSC-POSE-SEED17.npz")))
    logged = d.pop("logged")
    pol = policy_from_state_dict(d)
    E, dev = 256, torch.device("cuda", 0)

    def rollout(cfg):
        env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20,
                         reward_shaping=True, ignore_done=True)
        sim = env.sim
        obs = sim.reset()
        ret, lifted = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
        act, rew, done = torch.empty(E, 7, device=dev), torch.empty(E, device=dev), torch.empty(E, dtype=torch.uint8, device=dev)
        for _ in range(500):
            pol.get_actions(obs, deterministic=True, out=act)
            sim.step(act, obs, rew, done)
            ret += rew
            lifted = torch.maximum(lifted, (obs[:, 34] > 0.84).float())          # object-state cube z above the table top + 4 cm
        env.close()
        return ret.cpu().numpy(), lifted.mean().item()

    cfg = suite.load_controller_config(default_controller="OSC_POSE")
    ret, lifted = rollout(cfg)
    assert ret.mean() > 250.0 and lifted > 0.5, (ret.mean(), lifted)               # logged: 364; the untrained level is ~7-20
    assert abs(ret.max() - logged.max()) < 25.0, (ret.max(), logged.max())         # best episode == best logged evaluation (486.4)
    cfg["orientation_delta"] = "axis_angle"                                         # the robosuite >= 1.1 convention: the policy does not transfer
    ret2, lifted2 = rollout(cfg)
    assert ret2.mean() < 60.0 and lifted2 < 0.1, (ret2.mean(), lifted2)
