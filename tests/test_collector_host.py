"""Host-side logic of the fused collector / data-parallel training path, on CPU (no GPU, no compute calls through the C-ABI):
network handles with the reference's constructor signatures and their snapshot format, combination of the device-reduced path
statistics (and its world-2 gloo collective), the emulator's truncation counters and model-driven solver option."""
import io
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_reference_constructor_signatures_and_rlkit_init():
    """util/rlkit_utils.py:64-92 builds the networks with these keyword arguments, before any trainer exists."""
    from robosuite_benchmark_b200.sac import FlattenMlp, MakeDeterministic, TanhGaussianPolicy
    np.random.seed(7)
    qf = FlattenMlp(input_size=49, output_size=1, hidden_sizes=[256, 256])
    pol = TanhGaussianPolicy(obs_dim=42, action_dim=7, hidden_sizes=[256, 256])
    h = qf.host_params()
    assert h["W0"].shape == (49, 256) and h["W2"].shape == (256, 1) and np.abs(h["W0"]).max() <= 1 / np.sqrt(49) and (h["b0"] == 0.1).all()
    assert np.abs(h["W2"]).max() <= 3e-3 and np.abs(pol.host_params()["W2"]).max() <= 1e-3 and pol.host_params()["W2"].shape == (256, 14)
    sd = pol.state_dict()                                     # rlkit module names, torch [out, in] layout
    assert list(sd) == ["fc0.weight", "fc0.bias", "fc1.weight", "fc1.bias", "last_fc.weight", "last_fc.bias", "last_fc_log_std.weight", "last_fc_log_std.bias"]
    assert tuple(sd["fc0.weight"].shape) == (256, 42) and tuple(sd["last_fc_log_std.weight"].shape) == (7, 256)
    assert qf.to("cpu") is qf and qf.train(False) is qf and MakeDeterministic(pol).stochastic_policy is pol      # what CustomTorchBatchRLAlgorithm.to / training_mode call
    np.random.seed(7)
    again = FlattenMlp(input_size=49, output_size=1, hidden_sizes=[256, 256]).host_params()
    assert all((again[k] == h[k]).all() for k in h)           # np.random.seed (scripts/train.py:112) fixes the init
    with pytest.raises(NotImplementedError):
        FlattenMlp(input_size=49, output_size=1, hidden_sizes=[400, 300])
    if not torch.cuda.is_available():
        from robosuite_benchmark_b200.backend import RsbError
        with pytest.raises(RsbError):
            pol.get_action(np.zeros(42))                      # no CPU fallback for the forward


def test_snapshot_is_torch_loadable_and_returns_policy_objects():
    """params.pkl written with torch.save must come back through the reference's plain `torch.load(path)` (util/rlkit_utils.py:173) -- which
    is weights_only=True on torch >= 2.6 -- as objects exposing get_action / state_dict."""
    from robosuite_benchmark_b200.sac import FlattenMlp, MakeDeterministic, TanhGaussianPolicy, register_safe_globals
    np.random.seed(3)
    pol = TanhGaussianPolicy(obs_dim=46, action_dim=8, hidden_sizes=[256, 256])
    snap = {"trainer/policy": pol, "evaluation/policy": MakeDeterministic(pol), "trainer/qf1": FlattenMlp(input_size=54, output_size=1, hidden_sizes=[256, 256])}
    buf = io.BytesIO()
    torch.save(snap, buf)
    register_safe_globals()
    buf.seek(0)
    back = torch.load(buf)
    ev = back["evaluation/policy"]
    assert isinstance(ev, MakeDeterministic) and isinstance(ev.stochastic_policy, TanhGaussianPolicy) and hasattr(ev, "get_action")
    assert ev.stochastic_policy.obs_dim == 46 and ev.stochastic_policy.action_dim == 8
    for k, v in pol.state_dict().items():
        assert torch.equal(v, ev.state_dict()[k])
    assert torch.equal(back["trainer/qf1"].state_dict()["fc0.weight"], snap["trainer/qf1"].state_dict()["fc0.weight"])


class _FakeRing:
    action_dim = 3


class _FakeRound:
    def __init__(self, rew, act):
        self.T, self.N = rew.shape
        self.ring = _FakeRing()
        self.rew, self.act = rew, act

    def host_stats(self, expl_len=2):
        r, a = self.rew.astype(np.float64), self.act.astype(np.float64)
        ret, eret = r.sum(0), r[:expl_len].sum(0)
        out = []
        for x in (r, ret, eret, a):
            out += [x.sum(), (x * x).sum(), x.max(), x.min()]
        return np.array(out)

    def __iter__(self):
        for n in range(self.N):
            yield dict(rewards=self.rew[:, n].reshape(-1, 1), actions=self.act[:, n])


def test_combined_device_statistics_equal_rlkit_statistics():
    from robosuite_benchmark_b200.algorithm import batched_path_information, get_custom_generic_path_information
    rng = np.random.default_rng(0)
    rounds = [_FakeRound(rng.uniform(0, 1, (5, 4)).astype(np.float32), rng.uniform(-1, 1, (5, 4, 3)).astype(np.float32)),
              _FakeRound(rng.uniform(0, 1, (3, 4)).astype(np.float32), rng.uniform(-1, 1, (3, 4, 3)).astype(np.float32))]
    dev = batched_path_information(rounds, stat_prefix="", custom=True)
    ref = get_custom_generic_path_information([p for r in rounds for p in r], 2, 1.0)
    assert list(dev) == list(ref)
    for k in ref:
        assert abs(dev[k] - ref[k]) < 1e-6, k            # the numpy arm reduces fp32 arrays, the device arm fp64 accumulators
    assert batched_path_information([], custom=True) == OrderedDict()


def _stats_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robosuite_benchmark_b200.algorithm import batched_path_information
    from robosuite_benchmark_b200.parallel import reduce_path_stats
    rng = np.random.default_rng(100 + rank)
    rounds = [_FakeRound(rng.uniform(0, 1, (4, 3)).astype(np.float32), rng.uniform(-1, 1, (4, 3, 3)).astype(np.float32))]
    out[rank] = dict(batched_path_information(rounds, custom=True, reduce_fn=lambda a, c: reduce_path_stats(a, c, None)))
    dist.destroy_process_group()


def test_two_rank_statistics_collective_gloo():
    """The per-epoch collective of the logged statistics: both ranks end with the statistics of the union of their paths."""
    from robosuite_benchmark_b200.algorithm import get_custom_generic_path_information
    world, port = 2, 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_stats_worker, args=(world, port, out), nprocs=world, join=True)
        r0, r1 = dict(out[0]), dict(out[1])
    paths = []
    for rank in range(2):
        rng = np.random.default_rng(100 + rank)
        paths += list(_FakeRound(rng.uniform(0, 1, (4, 3)).astype(np.float32), rng.uniform(-1, 1, (4, 3, 3)).astype(np.float32)))
    ref = get_custom_generic_path_information(paths, 2, 1.0)
    assert r0 == r1 and list(r0) == list(ref)
    for k in ref:
        assert abs(r0[k] - ref[k]) < 1e-6, k


def test_experiment_step_counts_split_over_ranks():
    """algorithm_kwargs counts are job totals: under W ranks each rank collects ceil(count / W) (algorithm.build_experiment)."""
    import inspect
    from robosuite_benchmark_b200 import algorithm
    src = inspect.getsource(algorithm.build_experiment)
    assert "-(-ak[k] // world)" in src and "replay_buffer_size\"] // world" in src
    assert algorithm.distributed_context() == (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def test_emulated_device_code_counts_truncation_and_reads_solver_option():
    """The device source itself (csrc/rsb_dev.h compiled for the host): with generous limits no truncation event is counted; with ncon_max = 2 a
    cube resting on the table (4 corner contacts) is truncated in every control step -- and the counter says so.  The Newton settings are the
    model's <option> ("fp32" override, or the MJCF's own values with solver="model")."""
    from robosuite_benchmark_b200.controllers import load_controller_config
    from robosuite_benchmark_b200.model.tasks import SOLVER_FP32, build_task
    from tests.emu.emu import EmuEnv
    cc = load_controller_config(default_controller="OSC_POSE")
    m, t = build_task("Lift", "Panda", cc, ignore_done=True)
    emu = EmuEnv(m, t, 16, 64)
    assert emu.solver_option() == (SOLVER_FP32["iterations"], pytest.approx(SOLVER_FP32["tolerance"]), SOLVER_FP32["ls_iterations"], pytest.approx(0.01))
    emu.reset(seed=17, env_id=0)
    for k in range(3):
        emu.step(emu.random_action(17, 0, k))
    assert emu.counters() == (0, 0, 0)
    tight = EmuEnv(m, t, 2, 64)
    tight.reset(seed=17, env_id=0)
    for k in range(3):
        tight.step(tight.random_action(17, 0, k))
    assert tight.counters()[0] == 3 and tight.counters()[1] == 0
    m2, t2 = build_task("Lift", "Panda", cc, ignore_done=True, solver="model")
    assert EmuEnv(m2, t2, 16, 64).solver_option() == (100, pytest.approx(1e-8), 50, pytest.approx(0.01))
    m3, _ = build_task("Lift", "Panda", cc, ignore_done=True, solver=dict(iterations=30))
    assert m3.opt["iterations"] == 30 and m3.opt["tolerance"] == 1e-8
    with pytest.raises(ValueError):
        build_task("Lift", "Panda", cc, solver=dict(iteratons=3))


def test_reference_entry_point_resolves_on_compat_and_fails_loudly_without_a_gpu(tmp_path):
    """The reference's unmodified scripts/train.py (byte-compiled by oracle/build_ref.py) imports everything it needs from
    robosuite_benchmark_b200/compat/ and reaches the first device call; without a CUDA device that call raises -- no CPU fallback.
    (With a GPU the full run is tests/test_gpu_reference_dropin.py.)"""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from oracle import build_ref
    refpy = build_ref.build()
    if not refpy or not os.path.exists(refpy):
        pytest.skip("neither /root/reference nor a built oracle/_ref/refpy.bin here")
    run_train = "import runpy; runpy.run_module('scripts.train', run_name='__main__', alter_sys=True)"
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the GPU drop-in test")
    v = json.load(open(os.path.join(root, "tests", "golden", "variant_Lift-Panda-OSC-POSE-SEED17.json")))
    json.dump(v, open(tmp_path / "v.json", "w"))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "robosuite_benchmark_b200", "compat"), root, refpy]))
    r = subprocess.run([sys.executable, "-c", run_train, "--variant", str(tmp_path / "v.json"), "--seed", "17",
                        "--log_dir", str(tmp_path / "log")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "RsbError" in r.stderr and "CUDA device only" in r.stderr, r.stderr[-2000:]
    assert "ModuleNotFoundError" not in r.stderr and "ImportError" not in r.stderr
    # the import surface of SURVEY.md 8b, name by name
    code = ("import util.rlkit_utils as u, util.rlkit_custom as c, gtimer, robosuite, rlkit.pythonplusplus\n"
            "from robosuite.wrappers import GymWrapper\n"
            "assert {'OSC_POSE', 'JOINT_VELOCITY'} <= set(robosuite.controllers.ALL_CONTROLLERS)\n"
            "assert u.AGENTS == {'SAC', 'TD3'} and hasattr(c, 'CustomTorchBatchRLAlgorithm') and callable(gtimer.stamp)\n")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
