#!/usr/bin/env python3
"""Newton-iteration load of the env batch (developer tool, run under gpurun): who makes k_step's tail?
Per control step of the steady-state window: distribution of per-env Newton iterations (sum over the 25 substeps), how persistent the hard envs
are from step to step, and the kernel time under smaller iteration caps (RSB_SOLVER_ITERS) -- the price of the tail."""
import os, sys, subprocess
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import robosuite_benchmark_b200 as suite

def run(E=4096, pre=100, steps=40):
    dev = torch.device("cuda", 0)
    cfg = suite.load_controller_config(default_controller="OSC_POSE")
    env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
    sim = env.sim
    obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev); done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
    sim.reset(obs=obs)
    for k in range(pre):
        sim.random_actions(k, out=act); sim.step(act, obs, rew, done)
    its, ms = [], []
    for k in range(pre, pre + steps):
        sim.random_actions(k, out=act)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sim.step(act, obs, rew, done); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1)); its.append(sim.newton_iterations().cpu().numpy().copy())
    return np.array(its), np.array(ms), sim.info("envs_per_block")

if len(sys.argv) > 1 and sys.argv[1] == "time":
    its, ms, epb = run()
    print(f"RSB_SOLVER_ITERS={os.environ.get('RSB_SOLVER_ITERS')} RSB_EPB={os.environ.get('RSB_EPB')}: kernel ms mean {ms.mean():.3f} min {ms.min():.3f} max {ms.max():.3f}; iterations mean {its.mean():.1f} max {its.max()}")
    sys.exit(0)
its, ms, epb = run()
print(f"kernel ms mean {ms.mean():.3f}; envs/CTA {epb}")
for k in (0, 10, 20, 39):
    v = its[k]
    cta = v[: len(v) // epb * epb].reshape(-1, epb)
    print(f"step {k}: iters/env mean {v.mean():.1f} p50 {np.percentile(v,50):.0f} p90 {np.percentile(v,90):.0f} p99 {np.percentile(v,99):.0f} p99.9 {np.percentile(v,99.9):.0f} max {v.max()} | "
          f"per-CTA max: median {np.median(cta.max(1)):.0f} p90 {np.percentile(cta.max(1),90):.0f} max {cta.max()} | envs >=50: {(v>=50).sum()} >=75: {(v>=75).sum()} >=100: {(v>=100).sum()}; ms {ms[k]:.3f}")
hard = its >= 75
print("persistence of hard envs (>=75 iterations): P(hard at t+1 | hard at t) =", round(float((hard[1:] & hard[:-1]).sum() / max(hard[:-1].sum(), 1)), 3),
      "; P(hard at t+5 | hard at t) =", round(float((hard[5:] & hard[:-5]).sum() / max(hard[:-5].sum(), 1)), 3), "; fraction hard", round(float(hard.mean()), 4))
top = np.argsort(-its.sum(0))[:10]
print("10 hardest envs over the window (env: iterations per step):")
for e in top:
    print("  ", e, its[:, e].tolist())
print("corr(max iterations of the batch, kernel ms) =", round(float(np.corrcoef(its.max(1), ms)[0, 1]), 3))
for cap in (2, 4, 8):
    out = subprocess.run([sys.executable, __file__, "time"], env=dict(os.environ, RSB_SOLVER_ITERS=str(cap)), capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-500:])
