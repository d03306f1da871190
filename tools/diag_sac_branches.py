#!/usr/bin/env python3
"""Is the nondeterminism of the SAC update graph with side-stream branches a race or cuBLAS summation order? (run under gpurun)
Two trainers share the replay ring; before every update trainer B's parameters/moments are overwritten with A's, both do ONE update
(A: graph + parallel branches, B: eager, one stream) and the gradient buffers are compared."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from robosuite_benchmark_b200.sac import ParamStore, SACTrainer, EnvReplayBuffer
O, A, B = 42, 7, 128
rng = np.random.default_rng(0); n = 5800
rb = EnvReplayBuffer(8192, obs_dim=O, action_dim=A, device="cuda:0", seed=17)
rb.add_batch(*(torch.as_tensor(x, device="cuda:0") for x in (rng.normal(size=(n, O)).astype(np.float32) * 0.5, np.tanh(rng.normal(size=(n, A))).astype(np.float32),
             rng.uniform(0, 0.03, size=n).astype(np.float32), (rng.uniform(size=n) < 0.01).astype(np.uint8), rng.normal(size=(n, O)).astype(np.float32) * 0.5)))
kw = dict(discount=0.99, reward_scale=1.0, policy_lr=1e-3, qf_lr=5e-4, soft_target_tau=0.005, target_update_period=5)
sa, sb = ParamStore(O, A, "cuda:0", seed=3), ParamStore(O, A, "cuda:0", seed=3)
ta = SACTrainer(store=sa, batch_size=B, gemm="tcgen05", use_graph=True, seed=5, parallel_branches=True, **kw); ta.replay = rb
tb = SACTrainer(store=sb, batch_size=B, gemm="tcgen05", use_graph=False, seed=5, parallel_branches=False, **kw); tb.replay = rb
worst = 0.0; bad = 0
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 600):
    for name in ("flat", "m", "v", "target"):
        getattr(sb, name).copy_(getattr(sa, name))
    tb.bc.copy_(ta.bc); tb.alpha.copy_(ta.alpha)
    ta.train_step(); tb.train_step()
    torch.cuda.synchronize()
    ga, gb = sa.grad, sb.grad
    rel = ((ga - gb).abs().max() / gb.abs().max().clamp_min(1e-12)).item()
    worst = max(worst, rel)
    if rel > 1e-4:
        bad += 1
        if bad <= 5: print("step", step, "max |dG| / max |G| =", rel)
print("updates", step + 1, "worst relative gradient deviation", worst, "updates above 1e-4:", bad)
