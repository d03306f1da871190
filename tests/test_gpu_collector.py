"""GPU tests of the fused collector data path, all through the C-ABI:

  * k_policy_act against the fp64 oracle restatement of TanhGaussianPolicy.get_action (oracle/sac_oracle.py policy_act), stochastic and
    deterministic, plain rows and ring rows that wrap;
  * the ring the env kernels fill in place == the ring the copy-based collector of round 1 fills, bit for bit;
  * device-side path statistics against numpy on the same transitions;
  * contact / constraint-row truncation counters are 0 for all six families over a full 500-step episode;
  * the solver runs with the model's <option> (nothing hard-coded on the device).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def _make(env="Lift", robots="Panda", ctrl="OSC_POSE", n=64, seed=17, base=0, **kw):
    import robosuite_benchmark_b200 as suite
    cfg = suite.load_controller_config(default_controller=ctrl)
    return suite.make(env, robots, controller_configs=cfg, num_envs=n, batched=True, device="cuda:0", seed=seed, env_id_base=base,
                      horizon=kw.pop("horizon", 500), control_freq=20, reward_shaping=True, ignore_done=kw.pop("ignore_done", True), **kw)


@pytest.mark.parametrize("O,A,n", [(42, 7, 100), (89, 14, 33), (46, 8, 4096)])
def test_policy_forward_kernel_matches_oracle(torch_cuda, O, A, n):
    torch = torch_cuda
    from oracle.sac_oracle import policy_act
    from robosuite_benchmark_b200.sac import MakeDeterministic, ParamStore, TanhGaussianPolicy
    store = ParamStore(O, A, "cuda:0", seed=3)
    # larger head weights than rlkit's 1e-3 init so that mean / log_std (and the clamp) matter
    g = torch.Generator(device="cpu").manual_seed(1)
    store.P["p_W2"].copy_(torch.randn(256, 2 * A, generator=g) * 0.08); store.P["p_b2"].copy_(torch.randn(2 * A, generator=g) * 0.5)
    pol = TanhGaussianPolicy.of(store, seed=99, env_id_base=5000)
    params, _ = store.to_host()
    obs = torch.randn(n, O, generator=g).mul(0.7).to("cuda:0")
    m = min(n, 200)                                           # rows checked against the (pure-Python Philox) oracle
    for step in (0, 7):
        a = pol.get_actions(obs, step=step).cpu().numpy()
        ref = policy_act(params, obs[:m].cpu().numpy(), seed=99, env_ids=5000 + np.arange(m), step=step)
        assert np.abs(a[:m] - ref).max() < 2e-5, np.abs(a[:m] - ref).max()
        assert np.isfinite(a).all() and np.abs(a).max() <= 1.0
    a0, a7 = pol.get_actions(obs, step=0), pol.get_actions(obs, step=7)
    assert (a0 != a7).any() and torch.equal(a0, pol.get_actions(obs, step=0))          # noise keyed by the step, reproducible
    det = MakeDeterministic(pol).get_actions(obs).cpu().numpy()
    assert np.abs(det[:m] - policy_act(params, obs[:m].cpu().numpy(), deterministic=True)).max() < 2e-5
    one, info = pol.get_action(obs[3].cpu().numpy(), deterministic=True)              # rlkit's single-observation call
    assert one.dtype == np.float64 and info == {} and np.abs(one - det[3]).max() < 1e-6


def test_policy_forward_reads_and_writes_ring_rows_with_wrap(torch_cuda):
    torch = torch_cuda
    from robosuite_benchmark_b200.sac import EnvReplayBuffer, ParamStore, TanhGaussianPolicy, _ptr
    O, A, n, cap = 42, 7, 50, 128
    store = ParamStore(O, A, "cuda:0", seed=3)
    pol = TanhGaussianPolicy.of(store, seed=4)
    rb = EnvReplayBuffer(cap, obs_dim=O, action_dim=A, device="cuda:0")
    g = torch.Generator(device="cpu").manual_seed(2)
    rb._observations.copy_(torch.randn(cap, O, generator=g))
    rb._actions.fill_(-7.0)
    slot0 = 100                                               # rows 100..127 then 0..21
    pol.act_into(_ptr(rb._observations), O, _ptr(rb._actions), A, n, slot0=slot0, cap=cap, step=5)
    rows = (slot0 + torch.arange(n, device="cuda:0")) % cap
    ref = pol.get_actions(rb._observations[rows].contiguous(), step=5)
    assert torch.equal(rb._actions[rows], ref)
    untouched = torch.ones(cap, dtype=torch.bool, device="cuda:0"); untouched[rows] = False
    assert (rb._actions[untouched] == -7.0).all()


@pytest.mark.parametrize("env,robots,ctrl,n,T,cap_extra", [("Lift", "Panda", "OSC_POSE", 64, 6, 17), ("TwoArmLift", ["Panda", "Panda"], "OSC_POSE", 20, 4, 0)])
def test_fused_ring_equals_copy_path_bit_for_bit(torch_cuda, env, robots, ctrl, n, T, cap_extra):
    """The same rounds collected twice from identical seeds: (a) fused -- reset / policy / step kernels read and write the replay ring in
    place; (b) round-1 flow -- scratch observation row, torch copies into a [T+1, N, O] stack, add_batch into the ring.  Every ring array
    must be identical, including across the ring's wrap-around, and so must the ring pointer / size bookkeeping."""
    torch = torch_cuda
    from robosuite_benchmark_b200.algorithm import BatchedPathCollector, CopyingPathCollector, NormalizedBoxEnv, add_paths
    from robosuite_benchmark_b200.sac import EnvReplayBuffer, ParamStore, TanhGaussianPolicy
    from robosuite_benchmark_b200.wrappers import GymWrapper
    cap = 2 * n * T + n + cap_extra                         # the third round wraps
    rings, stats = [], []
    for fused in (True, False):
        e = NormalizedBoxEnv(GymWrapper(_make(env, robots, ctrl, n=n, seed=23, base=1000)))
        store = ParamStore(e.observation_space.low.size, e.action_space.low.size, "cuda:0", seed=3)
        store.P["p_W2"].mul_(50.0)                            # actions well away from 0 so the arms move
        pol = TanhGaussianPolicy.of(store, seed=23, env_id_base=1000)
        rb = EnvReplayBuffer(cap, e, device="cuda:0")
        col = BatchedPathCollector(e, pol, replay=rb) if fused else CopyingPathCollector(e, pol)
        st = []
        for rnd in range(3):
            paths = col.collect_new_paths(T, n * T, discard_incomplete_paths=False)
            assert len(paths) == 1 and paths[0].T == T and paths[0].N == n
            if fused:
                assert rb._size == min(cap, rnd * n * T), "in-place rounds become visible to sampling only through add_paths"
            add_paths(rb, paths)
            st.append(paths[0].host_stats().copy())
        torch.cuda.synchronize()
        assert e.sim.counters() == dict(ncon_overflow=0, nefc_overflow=0, steps_after_done=0)
        rings.append(rb); stats.append(st)
        e.sim.close()
    a, b = rings
    assert (a._top, a._size) == (b._top, b._size) == ((3 * n * T) % cap, cap)
    for name in ("_observations", "_actions", "_rewards", "_terminals", "_next_obs"):
        x, y = getattr(a, name), getattr(b, name)
        assert torch.equal(x, y), (name, (x != y).sum().item())
    assert a._rewards.abs().sum().item() > 0 and (a._observations[: n] != a._observations[n: 2 * n]).any()
    for sa, sb in zip(*stats):
        assert np.array_equal(sa, sb)


def test_device_path_statistics_match_numpy(torch_cuda):
    torch = torch_cuda
    from robosuite_benchmark_b200.algorithm import (BatchedPathCollector, NormalizedBoxEnv, batched_path_information, get_custom_generic_path_information,
                                                    get_generic_path_information)
    from robosuite_benchmark_b200.sac import MakeDeterministic, ParamStore, TanhGaussianPolicy
    from robosuite_benchmark_b200.wrappers import GymWrapper
    n, T = 300, 9
    e = NormalizedBoxEnv(GymWrapper(_make(n=n)))
    store = ParamStore(42, 7, "cuda:0", seed=3); store.P["p_W2"].mul_(30.0)
    col = BatchedPathCollector(e, MakeDeterministic(TanhGaussianPolicy.of(store)), deterministic=True, expl_len=5)
    rounds = col.collect_new_paths(T, 2 * n * T, discard_incomplete_paths=True)          # two rounds in the collector's scratch ring
    assert len(rounds) == 2 and col.get_diagnostics()["path length Mean"] == T
    dev = batched_path_information(rounds, custom=True)
    paths = [p for r in rounds for p in r]                     # numpy path dicts (rewards [T,1], actions [T,A]) from the same ring rows
    assert len(paths) == 2 * n
    ref = get_custom_generic_path_information(paths, 5, 1.0)
    assert list(dev.keys()) == list(ref.keys())
    for k in ref:
        assert abs(dev[k] - ref[k]) <= 1e-9 + 1e-6 * abs(ref[k]), (k, dev[k], ref[k])
    dev2, ref2 = batched_path_information(rounds, custom=False), get_generic_path_information(paths)
    assert list(dev2.keys()) == list(ref2.keys()) and all(abs(dev2[k] - ref2[k]) <= 1e-9 + 1e-6 * abs(ref2[k]) for k in ref2)


def test_collector_truncates_at_the_horizon_and_raises_after_done(torch_cuda):
    """ignore_done=False: an episode ends at the horizon with terminal = 1; the collector never steps a terminated episode,
    and a caller that does gets robosuite's ValueError instead of garbage rows."""
    torch = torch_cuda
    from robosuite_benchmark_b200.algorithm import BatchedPathCollector, NormalizedBoxEnv, add_paths
    from robosuite_benchmark_b200.sac import EnvReplayBuffer, ParamStore, TanhGaussianPolicy
    from robosuite_benchmark_b200.wrappers import GymWrapper
    n = 8
    e = NormalizedBoxEnv(GymWrapper(_make(n=n, horizon=5, ignore_done=False)))
    store = ParamStore(42, 7, "cuda:0", seed=3)
    rb = EnvReplayBuffer(1024, e, device="cuda:0")
    col = BatchedPathCollector(e, TanhGaussianPolicy.of(store), replay=rb)
    paths = col.collect_new_paths(12, 15 * n, discard_incomplete_paths=False)             # asks for 12-step paths of a 5-step episode
    assert [p.T for p in paths] == [5, 5, 5]                 # rlkit: the rollout breaks on done, the collector starts the next one
    add_paths(rb, paths)
    term = rb._terminals[: 15 * n].view(3, 5, n).cpu().numpy()
    assert (term[:, :4] == 0).all() and (term[:, 4] == 1).all()
    e.sim.step_ring(rb.ring, rb.top, False)                  # one more step on the terminated episodes
    torch.cuda.synchronize()
    assert e.sim.counters()["steps_after_done"] == n
    with pytest.raises(ValueError):
        col.collect_new_paths(1, 0, discard_incomplete_paths=False)


FAMILIES = [("Lift", "Panda", "OSC_POSE"), ("Lift", "Panda", "JOINT_VELOCITY"), ("Lift", "Sawyer", "OSC_POSITION"), ("Door", "Panda", "JOINT_VELOCITY"),
            ("Stack", "Sawyer", "OSC_POSE"), ("TwoArmLift", ["Panda", "Panda"], "OSC_POSE")]


@pytest.mark.parametrize("env,robots,ctrl", FAMILIES)
def test_no_contact_or_row_truncation_over_a_full_episode(torch_cuda, env, robots, ctrl):
    """contacts beyond ncon_max / rows beyond nefc_max used to be dropped silently.  They are counted now; a full
    500-step random-action episode of 4096 envs must not truncate once, for every family."""
    torch = torch_cuda
    e = _make(env, robots, ctrl, n=4096)
    sim = e.sim
    obs = sim.reset()
    act = torch.empty(4096, sim.act_dim, device="cuda:0"); rew = torch.empty(4096, device="cuda:0"); done = torch.empty(4096, dtype=torch.uint8, device="cuda:0")
    for k in range(500):
        sim.random_actions(k, out=act)
        sim.step(act, obs, rew, done)
    c = sim.counters()
    it = sim.newton_iterations().cpu().numpy()
    assert c == dict(ncon_overflow=0, nefc_overflow=0, steps_after_done=0), (c, sim.info("ncon_max"), sim.info("nefc_max"))
    assert torch.isfinite(obs).all() and it.min() >= 0 and it.max() <= 25 * sim.solver_option()[0]
    sim.close()


def test_solver_option_comes_from_the_model(torch_cuda):
    """iterations / tolerance / ls_iterations / ls_tolerance are read from the model's <option>; the fp32 budget is an
    explicit override of the model on the host, "model" keeps the MJCF's own values."""
    e = _make(n=4)
    assert e.sim.solver_option() == (12, pytest.approx(1e-6), 24, pytest.approx(0.01))
    e.sim.close()
    e = _make(n=4, solver="model")
    assert e.sim.solver_option() == (100, pytest.approx(1e-8), 50, pytest.approx(0.01)) and e.model.opt["iterations"] == 100
    e.sim.close()
    e = _make(n=4, solver=dict(iterations=30, tolerance=1e-7))
    assert e.sim.solver_option() == (30, pytest.approx(1e-7), 50, pytest.approx(0.01))
    e.sim.close()
