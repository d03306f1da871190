"""CPU checks of the SAC oracle against the reference's committed known answers, and of the host Philox."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sac_epoch0_known_answers.json")))
DIMS = {"Lift-Panda-OSC-POSE": (42, 7), "Door-Panda-JOINT-VELOCITY": (46, 8), "Stack-Sawyer-OSC-POSE": (55, 7),
        "TwoArmLift-PandaPanda-OSC-POSE": (89, 14)}


def _untrained_batch(rng, O, A, B, r_mean):
    """A batch shaped like the warm-up data of an untrained policy: bounded observations, tanh-Gaussian actions, small rewards."""
    obs = rng.normal(size=(B, O)).astype(np.float32) * 0.5
    nxt = obs + rng.normal(size=(B, O)).astype(np.float32) * 0.05
    act = np.tanh(rng.normal(size=(B, A))).astype(np.float32)
    rew = rng.uniform(0, 2 * r_mean, size=(B, 1)).astype(np.float32)
    return dict(observations=obs, actions=act, rewards=rew, terminals=np.zeros((B, 1), np.float32), next_observations=nxt)


@pytest.mark.parametrize("run", sorted(GOLD))
def test_first_update_known_answers(run):
    from oracle.sac_oracle import SacOracle
    from robosuite_benchmark_b200.sac import init_host_params
    fam = run.rsplit("-SEED", 1)[0]
    O, A = DIMS[fam]
    g, kw, B = GOLD[run]["epoch0"], GOLD[run]["trainer_kwargs"], GOLD[run]["batch_size"]
    seed = int(run.rsplit("SEED", 1)[1])
    params = init_host_params(O, A, seed=seed)
    targets = {k: v.copy() for k, v in params.items() if k.startswith("q_")}
    orc = SacOracle(params, targets, O, A, discount=kw["discount"], reward_scale=kw["reward_scale"], policy_lr=kw["policy_lr"],
                    qf_lr=kw["qf_lr"], soft_target_tau=kw["soft_target_tau"], target_update_period=kw["target_update_period"])
    rng = np.random.default_rng(seed)
    batch = _untrained_batch(rng, O, A, B, 0.013)
    orc.train(batch, rng.normal(size=(2 * B, A)).astype(np.float32))
    s = orc.stats
    # bit-identical across all 145 committed runs: one Adam step of -policy_lr on log_alpha = 0, and a vanishing alpha loss
    assert np.float32(s["Alpha"]) == np.float32(g["trainer/Alpha"]) == np.float32(0.9990004897117615)
    assert s["Alpha Loss"] == 0.0 and g["trainer/Alpha Loss"] == 0.0
    # statistical known answers of an untrained tanh-Gaussian policy (mu ~ 0, sigma ~ 1): logged value within Monte-Carlo error
    se = g["trainer/Log Pis Std"] / np.sqrt(B)
    assert abs(s["Log Pis Mean"] - g["trainer/Log Pis Mean"]) < 5 * se + 0.05
    # policy loss = 1.0 * mean(log pi) - mean(min Q) with |Q| ~ 1e-2 at init: alpha = 1 (pre-update) in the loss
    assert abs(s["Policy Loss"] - s["Log Pis Mean"]) < 0.05 and abs(g["trainer/Policy Loss"] - g["trainer/Log Pis Mean"]) < 0.05
    # TD target = r + 0.99 * (min Q_target - alpha * log pi'), alpha = 1
    assert abs(s["Q Targets Mean"] - (0.013 - 0.99 * g["trainer/Log Pis Mean"])) < 0.35
    # QF loss = (E[y] - E[q])^2 + Var(y - q)
    assert abs(g["trainer/QF1 Loss"] - (g["trainer/Q Targets Mean"] ** 2 + g["trainer/Q Targets Std"] ** 2)) < 0.05 * g["trainer/QF1 Loss"]
    assert abs(s["QF1 Loss"] - g["trainer/QF1 Loss"]) < 0.25 * g["trainer/QF1 Loss"]


def test_alpha_after_1001_updates_matches_logged_epoch1():
    """Alpha at epoch 1 in the logs is exp(-1.001): 1001 Adam steps whose gradient sign never changes move log_alpha by lr each."""
    import torch
    la = torch.zeros(1, requires_grad=True)
    opt = torch.optim.Adam([la], lr=1e-3)
    for _ in range(1001):
        opt.zero_grad(); (-(la * (-4.7 - 7.0))).sum().backward(); opt.step()   # log pi + target_entropy < 0 throughout epoch 0
    for run, g in GOLD.items():
        assert abs(float(la.detach().exp()) - g["epoch1"]["trainer/Alpha"]) < 2e-4


def test_host_philox_matches_c_oracle():
    from oracle.oracle import philox
    from robosuite_benchmark_b200.philox import philox4x32
    rng = np.random.default_rng(0)
    for _ in range(50):
        seed, env, stream, idx = (int(x) for x in rng.integers(0, 2 ** 31, size=4))
        ref = philox(seed, env, stream, idx)
        got = philox4x32([env & 0xFFFFFFFF, env >> 32, stream, idx], [seed & 0xFFFFFFFF, seed >> 32])
        assert list(ref) == got


def test_replay_index_rule_is_uniform_and_in_range():
    from oracle.sac_oracle import replay_indices
    idx = replay_indices(17, 3, 4096, 5800)
    assert idx.min() >= 0 and idx.max() < 5800 and len(np.unique(idx)) > 2500
    assert (replay_indices(17, 3, 64, 5800) == idx[:64]).all() and (replay_indices(17, 4, 64, 5800) != idx[:64]).any()
