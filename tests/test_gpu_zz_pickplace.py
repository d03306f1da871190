"""PickPlace (single-object mode: runs/PickPlace{Can,Milk}-{Panda,Sawyer}-OSC-POSE-*) TwoArmPegInHole (runs/TwoArmPegInHole-*-OSC-POSE-*) and NutAssemblyRound (runs/NutAssemblyRound-*-OSC-POSE-*) on the CUDA kernels, through the C-ABI: the checks of tests/test_pickplace.py
and of the other families' GPU tests, on the real kernels.  The file sorts after the other GPU tests on purpose: these families were added last."""
import json
import os

import numpy as np
import pytest

from tests.test_gpu_parity import test_other_config_families_one_control_step as _one_control_step

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FAMILIES = [("PickPlaceCan", ["Panda"], "OSC_POSE"), ("PickPlaceMilk", ["Sawyer"], "OSC_POSE"), ("PickPlaceCan", ["Sawyer"], "OSC_POSE"), ("PickPlaceMilk", ["Panda"], "OSC_POSE"),
            ("TwoArmPegInHole", ["Panda", "Panda"], "OSC_POSE"), ("TwoArmPegInHole", ["Panda", "Sawyer"], "OSC_POSE"), ("TwoArmPegInHole", ["Sawyer", "Sawyer"], "OSC_POSE"),
            ("NutAssemblyRound", ["Panda"], "OSC_POSE"), ("NutAssemblyRound", ["Sawyer"], "OSC_POSE"),
            ("Lift", ["Panda"], "JOINT_POSITION"), ("Stack", ["Sawyer"], "JOINT_TORQUE")]            # the joint-space controllers the other GPU tests do not reach


@pytest.mark.parametrize("env_name,robots,ctrl", FAMILIES)
def test_pickplace_one_control_step(env_name, robots, ctrl, torch_cuda):
    """reset, contact-pair lists (bit-exact), torques (1e-5 relative) and one control step (1e-4) from identical state, against the oracle."""
    _one_control_step(env_name, robots, ctrl, torch_cuda)


@pytest.mark.parametrize("robots", [["Panda", "Panda"], ["Sawyer", "Sawyer"]])
def test_handoff_one_control_step(robots, torch_cuda):
    """TwoArmHandoff: the same protocol as test_pickplace_one_control_step (seed 83, six environments 0 .. 5 control steps into their episodes, one control step
    from identical state), with the velocity bound of the contact-rich policy-state tests (5e-4): the hammer lands on its head and settles during the first control
    steps (12 contacts switching on and off), where the CPU emulator already differs from the fp64 oracle by 7e-5 (profiles/r2_new_families_cpu_parity.txt)."""
    from oracle.oracle import OracleEnv
    from robosuite_benchmark_b200.backend import BatchSim
    from robosuite_benchmark_b200.controllers import load_controller_config
    from robosuite_benchmark_b200.model.tasks import build_task
    torch = torch_cuda
    m, t = build_task("TwoArmHandoff", robots, load_controller_config(default_controller="OSC_POSE"), ignore_done=True)
    nc, ne, n = t["ncon_max"], t["nefc_max"], 6
    sim = BatchSim(m, t, n, device="cuda:0", seed=83)
    obs0 = sim.reset().cpu().numpy()
    orcs, rows, acts = [], [], []
    for i in range(n):
        orc = OracleEnv(m, t, ncon_max=nc, nefc_max=ne)
        o = orc.reset(seed=83, env_id=i, episode=0)
        assert np.abs(o - obs0[i]).max() < 2e-6
        for k in range(i):
            orc.step(orc.random_action(83, i, k))
        qpos, qvel, warm, cs = orc.get_state()
        rows.append(sim.pack_state(qpos, qvel, warm, cs, timestep=i, episode=1, bpose=orc.get_bpose())[0])
        acts.append(orc.random_action(83, i, i)); orcs.append(orc)
    a = torch.as_tensor(np.stack(acts), dtype=torch.float32, device=sim.device)
    sim.set_state(torch.as_tensor(np.stack(rows)))
    obs, rew, _ = sim.step(a)
    st = sim.unpack_state(sim.get_state().cpu().numpy())
    obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
    for i, orc in enumerate(orcs):
        orc.set_timestep(i)
        o, r, _ = orc.step(acts[i])
        qpos, qvel, _, _ = orc.get_state()
        assert np.abs(qpos - st["qpos"][i]).max() <= 1e-4 and np.abs(qvel - st["qvel"][i]).max() <= 5e-4, (i, np.abs(qvel - st["qvel"][i]).max())
        assert np.abs(o - obs[i]).max() <= 5e-4 and abs(r - rew[i]) <= 1e-5
    assert sim.counters() == dict(ncon_overflow=0, nefc_overflow=0, steps_after_done=0)
    sim.close()


@pytest.mark.parametrize("env_name,robots", [("PickPlaceCan", "Panda"), ("PickPlaceMilk", "Sawyer"), ("TwoArmPegInHole", ["Panda", "Sawyer"])])
def test_pickplace_full_episode_no_truncation_and_reward_mean(env_name, robots, torch_cuda):
    """A full 500-step random-action episode of 2048 envs: no contact / row truncation, finite observations; and the reward mean over the first 150 control steps
    within 1 % of the oracle's over the same (seed, env id) streams (north_star: random-action episode reward means match within 1 %)."""
    import robosuite_benchmark_b200 as suite
    from oracle.oracle import OracleEnv
    torch = torch_cuda
    n, n_ref, steps, steps_ref = 2048, 12, 500, 150
    env = suite.make(env_name, robots, controller_configs=suite.load_controller_config(default_controller="OSC_POSE"), num_envs=n, device="cuda:0", seed=59,
                     horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
    sim = env.sim
    assert (sim.obs_dim, sim.act_dim) == ((73, 12) if env_name == "TwoArmPegInHole" else (46, 7))
    obs = sim.reset()
    act = torch.empty(n, sim.act_dim, device="cuda:0"); rew = torch.empty(n, device="cuda:0"); done = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    tot = torch.zeros(n, device="cuda:0")
    for k in range(steps):
        sim.random_actions(k, out=act)
        sim.step(act, obs, rew, done)
        tot += rew
        if k == steps_ref - 1:
            tot_ref = tot[:n_ref].cpu().numpy().copy()
    assert sim.counters() == dict(ncon_overflow=0, nefc_overflow=0, steps_after_done=0)
    assert torch.isfinite(obs).all() and torch.isfinite(tot).all()
    ref = np.zeros(n_ref)
    for i in range(n_ref):
        orc = OracleEnv(env.model, env.task, ncon_max=sim.info("ncon_max"), nefc_max=sim.info("nefc_max"))
        orc.reset(seed=59, env_id=i, episode=0)
        for k in range(steps_ref):
            ref[i] += orc.step(orc.random_action(59, i, k))[1]
    assert abs(tot_ref.mean() - ref.mean()) <= 0.01 * abs(ref.mean()), (tot_ref, ref)
    sim.close()


def test_committed_pickplace_policy_transfers_on_the_cuda_path(torch_cuda):
    """The reference's committed PickPlaceCan-Sawyer-OSC-POSE-SEED59 policy (trained against real robosuite + MuJoCo; its run logs 74 over the last 50 epochs, best
    189) rolled out deterministically in the batched CUDA env through the package's own evaluation path: it picks the can up and carries it."""
    import robosuite_benchmark_b200 as suite
    from robosuite_benchmark_b200.policy_io import DeterministicPolicy
    torch = torch_cuda
    d = dict(np.load(os.path.join(GOLDEN, "policy_PickPlaceCan-Sawyer-OSC-POSE-SEED59.npz")))
    logged, cfg = d.pop("logged"), json.loads(str(d.pop("env_kwargs")))
    W = {k: torch.as_tensor(v, device="cuda:0") for k, v in d.items()}
    pol = DeterministicPolicy({k: v.astype(np.float64) for k, v in d.items()})
    n = 256
    env = suite.make(cfg["env_name"], cfg["robots"], controller_configs=suite.load_controller_config(default_controller=cfg["controller"]), num_envs=n, device="cuda:0",
                     seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
    obs = env.reset()
    ret = torch.zeros(n, device="cuda:0"); best = torch.zeros(n, device="cuda:0")
    a0 = pol(obs[0].cpu().numpy().astype(np.float64))
    for k in range(500):
        h = obs
        i = 0
        while f"fc{i}.weight" in W:
            h = torch.relu(h @ W[f"fc{i}.weight"].T + W[f"fc{i}.bias"]); i += 1
        act = torch.tanh(h @ W["last_fc.weight"].T + W["last_fc.bias"])
        if k == 0:
            assert np.abs(act[0].cpu().numpy() - a0).max() < 5e-3                  # the torch forward here == policy_io's numpy forward
        obs, rew, done, _ = env.step(act)
        ret += rew; best = torch.maximum(best, rew)
    ret, best = ret.cpu().numpy(), best.cpu().numpy()
    # CPU oracle, 16 episodes: mean 71-102, best 166; a third to a half of the episodes grasp
    assert ret.mean() > 35.0 and ret.max() > 120.0 and (best >= 0.35).mean() > 0.2, (ret.mean(), ret.max(), (best >= 0.35).mean(), logged[-50:].mean())
    assert env.sim.counters()["steps_after_done"] == 0
    env.close()
