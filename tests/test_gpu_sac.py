"""GPU parity of the replay sampling and SAC update kernels (through the C-ABI of include/rsb_sac.h) against the CPU oracle
(oracle/sac_oracle.py, PyTorch fp32 autograd).  Replay indices: bit-exact.  Floating point: the non-GEMM kernels with fp32 cuBLAS products
(gemm="cublas_fp32") within 2e-5 relative of the oracle's gradients / 1e-6 absolute on parameters after an update; the PRODUCT path
(gemm="tcgen05": hand-written TF32 tensor-core kernel) within 2e-4 relative of an oracle whose products use TF32 operands as the tensor
core does, and within the loose TF32 bound (5 % Frobenius) of the plain fp32 oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
O, A, B = 42, 7, 128


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def _ring(torch, n=5800, cap=8192, seed=17):
    from robosuite_benchmark_b200.sac import EnvReplayBuffer
    rng = np.random.default_rng(0)
    rb = EnvReplayBuffer(cap, obs_dim=O, action_dim=A, device="cuda:0", seed=seed)
    data = dict(obs=rng.normal(size=(n, O)).astype(np.float32) * 0.5, act=np.tanh(rng.normal(size=(n, A))).astype(np.float32),
                rew=rng.uniform(0, 0.03, size=n).astype(np.float32), term=(rng.uniform(size=n) < 0.01).astype(np.uint8),
                nxt=rng.normal(size=(n, O)).astype(np.float32) * 0.5)
    for lo in range(0, n, 1450):                  # appended in batches, like the collectors do
        sl = slice(lo, min(n, lo + 1450))
        rb.add_batch(*(torch.as_tensor(data[k][sl], device="cuda:0") for k in ("obs", "act", "rew", "term", "nxt")))
    return rb, data


def test_replay_indices_bit_exact_and_rows_gathered(torch_cuda):
    from oracle.sac_oracle import replay_indices
    rb, data = _ring(torch_cuda)
    assert rb.get_diagnostics()["size"] == 5800
    for draw in range(3):
        b = rb.random_batch(B)
        idx = b["indices"].cpu().numpy()
        assert (idx == replay_indices(17, draw, B, 5800)).all()                # bit-exact sample indices
        assert (b["observations"].cpu().numpy() == data["obs"][idx]).all() and (b["next_observations"].cpu().numpy() == data["nxt"][idx]).all()
        assert (b["actions"].cpu().numpy() == data["act"][idx]).all() and (b["rewards"].cpu().numpy()[:, 0] == data["rew"][idx]).all()
        assert (b["terminals"].cpu().numpy()[:, 0] == data["term"][idx]).all()


def test_ring_wraps_like_rlkit(torch_cuda):
    torch = torch_cuda
    from robosuite_benchmark_b200.sac import EnvReplayBuffer
    rb = EnvReplayBuffer(10, obs_dim=2, action_dim=1, device="cuda:0")
    for k in range(4):
        n = 3
        v = torch.full((n, 2), float(k), device="cuda:0")
        rb.add_batch(v, v[:, :1], v[:, 0], torch.zeros(n, dtype=torch.uint8, device="cuda:0"), v + 0.5)
    assert rb._size == 10 and rb._top == 2
    assert rb._observations[:, 0].cpu().tolist() == [3.0, 3.0, 0.0, 1.0, 1.0, 1.0, 2.0, 2.0, 2.0, 3.0]


def _pair(torch, tf32, graph=False, period=5, oracle_tf32=None):
    from oracle.sac_oracle import SacOracle
    from robosuite_benchmark_b200.sac import ParamStore, SACTrainer
    store = ParamStore(O, A, "cuda:0", seed=3)
    params, targets = store.to_host()
    kw = dict(discount=0.99, reward_scale=1.0, policy_lr=1e-3, qf_lr=5e-4, soft_target_tau=0.005, target_update_period=period)
    tr = SACTrainer(store=store, batch_size=B, gemm="tcgen05" if tf32 else "cublas_fp32", use_graph=graph, seed=5, **kw)
    return store, tr, SacOracle(params, targets, O, A, tf32=oracle_tf32, **kw)


def test_update_matches_oracle_fp32(torch_cuda):
    torch = torch_cuda
    store, tr, orc = _pair(torch, tf32=False)
    rng = np.random.default_rng(1)
    for step in range(6):                          # crosses a Polyak step (period 5 -> steps 0 and 5)
        batch = dict(observations=rng.normal(size=(B, O)).astype(np.float32) * 0.5, actions=np.tanh(rng.normal(size=(B, A))).astype(np.float32),
                     rewards=rng.uniform(0, 0.03, size=(B, 1)).astype(np.float32), terminals=(rng.uniform(size=(B, 1)) < 0.05).astype(np.float32),
                     next_observations=rng.normal(size=(B, O)).astype(np.float32) * 0.5)
        eps = rng.normal(size=(2 * B, A)).astype(np.float32)
        g_ref = orc.train(batch, eps)
        tr.train_step(batch=batch, eps=eps)
        torch.cuda.synchronize()
        for k, gr in g_ref.items():
            got = store.G[k].cpu().numpy()
            scale = max(np.abs(gr.numpy()).max(), 1e-8)
            assert np.abs(got - gr.numpy()).max() <= 2e-5 * scale + 1e-9, (step, k)
        p_ref, t_ref = orc.params()
        p_got, t_got = store.to_host()
        for k in p_ref:
            # Adam normalises each element by its own |g|: where |g| ~ eps (1e-8) a 1e-5 relative gradient difference moves
            # the step by a visible fraction of lr, hence 2 % of lr here while the gradients themselves agree to 2e-5
            assert np.abs(p_ref[k] - p_got[k]).max() < 2e-5, (step, k)
        for k in t_ref:
            assert np.abs(t_ref[k] - t_got[k]).max() < 2e-6, (step, k)
        st = tr.get_diagnostics() if step == 0 else None
        if st:
            for k in ("QF1 Loss", "QF2 Loss", "Policy Loss", "Alpha", "Alpha Loss"):
                assert abs(st[k] - orc.stats[k]) <= 1e-5 * max(1.0, abs(orc.stats[k])), k
            assert np.float32(st["Alpha"]) == np.float32(0.9990004897117615)      # the reference's logged first-update value


def test_tf32_graph_update_close_to_oracle_and_to_eager(torch_cuda):
    torch = torch_cuda
    rb, _ = _ring(torch)
    store, tr, orc = _pair(torch, tf32=True, graph=True)
    tr.replay = rb
    store2, tr2, _ = _pair(torch, tf32=True, graph=False)
    tr2.replay = rb
    for step in range(7):
        tr.train_step(); tr2.train_step()
    torch.cuda.synchronize()
    a, _ = store.to_host(); b, _ = store2.to_host()
    for k in a:
        assert np.abs(a[k] - b[k]).max() < 1e-6, k           # graph replay == eager launches
    # against the fp32 oracle fed the same sampled batches and the same Philox noise
    from oracle.sac_oracle import replay_indices
    store3, tr3, orc3 = _pair(torch, tf32=True, graph=False)
    tr3.replay = rb
    tr3.train_step()
    torch.cuda.synchronize()
    idx = tr3.idx.cpu().numpy()
    assert (idx == replay_indices(17, 0, B, 5800)).all()
    batch = dict(observations=tr3.Xp[:B].cpu().numpy(), next_observations=tr3.Xp[B:].cpu().numpy(), actions=tr3.act.cpu().numpy(),
                 rewards=tr3.rew.cpu().numpy(), terminals=tr3.term.cpu().numpy())
    g_ref = orc3.train(batch, tr3.eps.cpu().numpy())
    for k, gr in g_ref.items():
        # TF32 inputs carry 10 mantissa bits; 256-term sums with cancellation through three chained layers forward and back:
        # relative Frobenius error of each gradient tensor <= 5 % (the fp32 mode of the same code is held to 2e-5 above)
        d = store3.G[k].cpu().numpy() - gr.numpy()
        assert np.linalg.norm(d) <= 5e-2 * max(np.linalg.norm(gr.numpy()), 1e-12), (k, np.linalg.norm(d) / np.linalg.norm(gr.numpy()))


def _random_batch(rng):
    return dict(observations=rng.normal(size=(B, O)).astype(np.float32) * 0.5, actions=np.tanh(rng.normal(size=(B, A))).astype(np.float32),
                rewards=rng.uniform(0, 0.03, size=(B, 1)).astype(np.float32), terminals=(rng.uniform(size=(B, 1)) < 0.05).astype(np.float32),
                next_observations=rng.normal(size=(B, O)).astype(np.float32) * 0.5)


def test_tcgen05_product_path_tight_against_tf32_operand_oracle(torch_cuda):
    """The PRODUCT path (gemm="tcgen05") against an oracle that computes every product -- forward, input gradient, weight gradient -- with
    TF32 operands (oracle/sac_oracle.py tf32_operand) and exact accumulation: what is left is fp32 accumulation order, so every gradient
    tensor agrees to 2e-4 of its largest entry (a dropped term, a wrong mask or a wrong operand would be O(1)).  The tensor core's operand
    reduction is identified from the data: truncation of the low 13 mantissa bits vs round-to-nearest -- the test reports which model fits
    and requires the better one to pass."""
    torch = torch_cuda
    rng = np.random.default_rng(11)
    batches = [(_random_batch(rng), rng.normal(size=(2 * B, A)).astype(np.float32)) for _ in range(3)]
    worst = {}
    for mode in ("trunc", "rna"):
        store, tr, orc = _pair(torch, tf32=True, graph=False, oracle_tf32=mode)
        w = 0.0
        for batch, eps in batches:
            g_ref = orc.train(batch, eps)
            tr.train_step(batch=batch, eps=eps)
            torch.cuda.synchronize()
            for k, gr in g_ref.items():
                got, ref = store.G[k].cpu().numpy(), gr.numpy()
                w = max(w, float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-8)))
            # parameters follow the oracle's so that the next update starts from the same point (Adam amplifies last-bit differences)
            p_ref, t_ref = orc.params()
            store.load_host(p_ref, t_ref); tr.refresh_alpha()
        worst[mode] = w
    from robosuite_benchmark_b200 import gemm
    assert gemm.timeouts() == 0
    best = min(worst, key=worst.get)
    print(f"tcgen05 vs TF32-operand oracle: max relative gradient error trunc {worst['trunc']:.2e}, rna {worst['rna']:.2e} -> operand model: {best}")
    assert worst[best] <= 2e-4, worst


def test_parallel_graph_branches_equal_the_serial_order(torch_cuda):
    """The update graph runs the target-Q forward and the weight-gradient GEMMs on side streams.  A missing dependency would be a timing-
    dependent race.  300 updates, each from IDENTICAL parameters (copied before every update): graph replay with the branches against eager
    launches on ONE stream; the gradient buffers must agree to fp32 round-off (cuBLAS may pick another summation order when streams run
    concurrently: last-bit differences, measured 1.6e-7 relative -- which Adam amplifies along a trajectory, hence the per-update form)."""
    torch = torch_cuda
    from robosuite_benchmark_b200.sac import ParamStore, SACTrainer
    rb, _ = _ring(torch)
    kw = dict(discount=0.99, reward_scale=1.0, policy_lr=1e-3, qf_lr=5e-4, soft_target_tau=0.005, target_update_period=5)
    sa, sb = ParamStore(O, A, "cuda:0", seed=3), ParamStore(O, A, "cuda:0", seed=3)
    ta = SACTrainer(store=sa, batch_size=B, gemm="tcgen05", use_graph=True, seed=5, parallel_branches=True, **kw); ta.replay = rb
    tb = SACTrainer(store=sb, batch_size=B, gemm="tcgen05", use_graph=False, seed=5, parallel_branches=False, **kw); tb.replay = rb
    worst = 0.0
    for _ in range(300):
        for name in ("flat", "m", "v", "target"):
            getattr(sb, name).copy_(getattr(sa, name))
        tb.bc.copy_(ta.bc); tb.alpha.copy_(ta.alpha)
        ta.train_step(); tb.train_step()
        torch.cuda.synchronize()
        worst = max(worst, ((sa.grad - sb.grad).abs().max() / sb.grad.abs().max().clamp_min(1e-12)).item())
    assert worst < 1e-5, worst
