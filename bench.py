#!/usr/bin/env python
"""bench.py -- env control-steps/s of the batched env.step hot path (BASELINE.json metric), one process per GPU.

Default workload at N GPUs (weak scaling): BASELINE.json configs[1] on every GPU -- Lift-Panda-OSC_POSE, 4096 batched envs, synthetic
tanh(N(0,1)) actions from the Philox stream shared with the oracle, horizon 500 with a batch reset at the horizon.  One bench "step" = one
20 Hz control step (25 physics substeps + controller + reward + observation) of every env of the batch.  `value` is measured with
states/actions resident in HBM (CUDA events around each step, L2 flushed between steps); `e2e` goes through the host-buffer C-ABI call
(pinned host actions in, obs/reward/done out) each step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--config lift|door|stack|twoarmlift|pickplacecan|peginhole|nutassemblyround|handoff] [--mode step|train] [--impl ours|reference]

--config selects the other BASELINE.json configs (configs[2] Door-Panda-JOINT_VELOCITY x 16384, configs[3] Stack-Sawyer-OSC_POSE,
configs[4] TwoArmLift-PandaPanda-OSC_POSE); the default (lift) line also carries a short steady-state measurement of each of them
(`other_configs`), the same batch under a trained policy's contact load (`trained_policy`), the mean over one whole 500-step episode (`full_episode`), the SAC updates/s legs (`sac`) and a short run of the
end-to-end training loop (`train`).  --mode train makes the training loop the timed workload: one bench "step" = one epoch of the
reference's loop (util/rlkit_custom.py:215-239: evaluation rollouts, exploration rollouts with per-step policy inference written
straight into the replay ring, add_paths, SAC updates sampled from that ring), under torchrun with the gradient all-reduce.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED, HORIZON = 17, 500
CONFIGS = {   # BASELINE.json configs[1..4]
    "lift": dict(env="Lift", robots=["Panda"], controller="OSC_POSE", envs=4096),
    "door": dict(env="Door", robots=["Panda"], controller="JOINT_VELOCITY", envs=16384),
    "stack": dict(env="Stack", robots=["Sawyer"], controller="OSC_POSE", envs=4096),
    "twoarmlift": dict(env="TwoArmLift", robots=["Panda", "Panda"], controller="OSC_POSE", envs=4096),
    # families of the reference's runs/ beyond BASELINE.json's configs, added last in round 2: parity-tested on the CPU emulator and by
    # tests/test_gpu_zz_pickplace.py; their rates ride along in the default line as `more_families`
    "pickplacecan": dict(env="PickPlaceCan", robots=["Panda"], controller="OSC_POSE", envs=4096),
    "peginhole": dict(env="TwoArmPegInHole", robots=["Panda", "Sawyer"], controller="OSC_POSE", envs=4096),
    "nutassemblyround": dict(env="NutAssemblyRound", robots=["Sawyer"], controller="OSC_POSE", envs=4096),
    "handoff": dict(env="TwoArmHandoff", robots=["Panda", "Panda"], controller="OSC_POSE", envs=4096),
}
UNIT = "control-steps/s"
PREROLL = 100          # untimed control steps after the initial reset (run_ours): the timed region sits on the steady-state part of the episode


def family(cfg):
    return f"{cfg['env']}-{''.join(cfg['robots'])}-{cfg['controller']}"


def metric_name(cfg):
    return "Lift-Panda-OSC env control-steps/s" if cfg is CONFIGS["lift"] else f"{family(cfg)} env control-steps/s"


def algorithmic_bytes_per_step(task, model) -> int:
    """bytes = 4*(2*S + 2*A + O + 2) with S = nq + 2*nv + C + 1 (OSC: C = 21 per arm, JOINT_VELOCITY P-law: 9)."""
    C = sum(21 if r["ctrl_type"] in (0, 1) else 9 for r in task["robot"])
    S = model.nq + 2 * model.nv + C + 1
    return 4 * (2 * S + 2 * task["act_dim"] + task["obs_dim"] + 2)


def measured_traffic(config_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE k_step launch from this round's `ncu --set full` capture (profiles/r2_kstep_traffic.json,
    written from the capture by tools/ncu_summary.py); None when no capture of this config exists -- never a number carried over from elsewhere."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_kstep_traffic.json"))).get(config_name)
    except Exception:
        return None


# ----------------------------------------------------------------------------- CPU baseline (oracle port, test infra)
def _cpu_worker(args):
    wid, steps, seed, cfg = args
    from oracle.oracle import OracleEnv
    from robosuite_benchmark_b200.controllers import load_controller_config
    from robosuite_benchmark_b200.model.tasks import build_task
    m, t = build_task(cfg["env"], cfg["robots"], load_controller_config(default_controller=cfg["controller"]), horizon=HORIZON,
                      control_freq=20, reward_shaping=True, ignore_done=True)
    env = OracleEnv(m, t, ncon_max=t["ncon_max"], nefc_max=t["nefc_max"])
    env.reset(seed=seed, env_id=wid, episode=0)
    for k in range(5):
        env.step(env.random_action(seed, wid, k))
    t0 = time.perf_counter()
    for k in range(5, 5 + steps):
        env.step(env.random_action(seed, wid, k))
    return time.perf_counter() - t0


def _host_cores():
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(cores, 64))


def cpu_baseline(cfg, steps_per_worker=12000):
    """The fp64 C oracle (a PORT of the reference's CPU path; the real robosuite+mujoco cannot be installed here) as one
    process per host core, same workload; steps/s summed over workers."""
    import multiprocessing as mp
    from oracle import oracle as _o
    _o.build()
    cores = _host_cores()
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, [(w, steps_per_worker, SEED, cfg) for w in range(cores)])
        wall = time.perf_counter() - t0
    value = sum(steps_per_worker / t for t in times)
    return dict(value=value, unit=UNIT, cores=cores, kind="port",
                sample=f"{cores} processes x {steps_per_worker} control steps of {family(cfg)} (single env each, fp64 C oracle), wall {wall:.1f}s")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, gpu_index):
        self.samples, self.reasons, self.proc, self.th = [], set(), None, None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((float(p[0]), float(p[1])))
                for nm, v in zip(names, p[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        sm = sorted(s[0] for s in self.samples)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(s[1] for s in self.samples), reasons=sorted(self.reasons))


# ----------------------------------------------------------------------------- arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    # bounded sample: every --steps "step" is `per` control steps on each host core
    per = 40
    t_warm = max(1, args.warmup)
    import multiprocessing as mp
    from oracle import oracle as _o
    _o.build()
    cores = _host_cores()
    total = per * (t_warm + args.steps)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, [(w, per * t_warm, SEED, cfg) for w in range(cores)])          # warm-up
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, [(w, per * args.steps, SEED, cfg) for w in range(cores)])
        wall = time.perf_counter() - t0
    value = sum(per * args.steps / t for t in times)
    line = {"impl": "reference", "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{family(cfg)}, one env per host core, tanh-Gaussian random actions, horizon {HORIZON}",
                       "note": "robosuite+mujoco are not installable here (no network, no wheels): this arm times the fp64 C oracle port of the reference CPU path"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{cores} processes x {per * args.steps} control steps ({total} incl. warm-up)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def _make_env(cfg, E, dev, env_id_base):
    import robosuite_benchmark_b200 as suite
    cc = suite.load_controller_config(default_controller=cfg["controller"])
    return suite.make(cfg["env"], cfg["robots"], controller_configs=cc, num_envs=E, batched=True, device=dev, seed=SEED, env_id_base=env_id_base,
                      horizon=HORIZON, control_freq=20, reward_shaping=True, ignore_done=True)


def steady_state_rate(cfg, E, dev, env_id_base, steps, preroll=PREROLL, full_episode=False):
    """Short measurement of one config: `steps` control steps after `preroll` untimed ones, CUDA events around k_step, L2 flushed between steps;
    optionally the mean over one whole episode (reset at step 0 .. horizon).  -> dict (per-GPU numbers; no cross-rank reduction)."""
    import torch
    env = _make_env(cfg, E, dev, env_id_base)
    sim = env.sim
    obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev)
    done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sim.reset(obs=obs)
    out = {"workload": f"{family(cfg)}, {E} batched envs", "envs_per_gpu": E,
           "algorithmic_bytes_per_env_step": algorithmic_bytes_per_step(env.task, env.model)}
    if full_episode:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(HORIZON):
            sim.random_actions(k, out=act); sim.step(act, obs, rew, done)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / HORIZON
        out["full_episode"] = {"ms_per_step": ms, "steps_per_s": E / ms * 1000.0, "control_steps": HORIZON,
                               "note": "mean over one whole episode from the batch reset (no L2 flush inside)"}
        sim.reset(obs=obs)
    for k in range(preroll):
        sim.random_actions(k, out=act); sim.step(act, obs, rew, done)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.zero_()
        sim.random_actions(preroll + k, out=act)
        ev[k][0].record(); sim.step(act, obs, rew, done); ev[k][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    out.update(kernel_ms=ms, steps_per_s=E / ms * 1000.0, timed_steps=steps, preroll=preroll, mean_reward=float(rew.mean().item()),
               truncation=sim.counters(), envs_per_block=sim.info("envs_per_block"), lanes=sim.info("lanes"), smem_bytes_per_env=sim.info("smem_bytes"))
    out["truncation"]["env_steps"] = E * (preroll + steps + (HORIZON if full_episode else 0))       # events are REPORTED, never silent (tests assert 0 over a 4096-env episode)
    env.close()
    return out


def trained_policy_rate(cfg, E, dev, env_id_base):
    """The same Lift batch driven by a TRAINED policy instead of random actions: the reference's committed Lift-Panda-OSC_POSE-SEED17 policy (weights in
    tests/golden, rolled out by the package's own policy kernel).  A policy that reaches, grasps and lifts keeps 10-20 contacts alive per env, so a control step
    costs more than under random actions; this is the load a late-training collector sees.  One whole episode; k_step timed with CUDA events on steps 100-119."""
    import numpy as np
    import torch
    from robosuite_benchmark_b200.rollout import policy_from_state_dict
    path = os.path.join(ROOT, "tests", "golden", "policy_Lift-Panda-OSC-POSE-SEED17.npz")
    if not os.path.exists(path):
        return None
    d = dict(np.load(path)); logged = d.pop("logged")
    pol = policy_from_state_dict(d)
    env = _make_env(cfg, E, dev, env_id_base)
    sim = env.sim
    obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev); ret = torch.zeros(E, device=dev)
    done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sim.reset(obs=obs)
    ev = []
    for k in range(HORIZON):
        pol.get_actions(obs, deterministic=True, out=act)
        if 100 <= k < 120:
            flush.zero_()
            ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            ev[-1][0].record(); sim.step(act, obs, rew, done); ev[-1][1].record()
        else:
            sim.step(act, obs, rew, done)
        ret += rew
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    out = {"workload": f"{family(cfg)}, {E} batched envs, deterministic actions of the committed SEED17 policy", "kernel_ms": ms, "steps_per_s": E / ms * 1000.0,
           "timed_steps": len(ev), "episode_return_mean": float(ret.mean().item()), "logged_return_last50_mean": float(logged[-50:].mean()),
           "truncation": sim.counters()}
    out["truncation"]["env_steps"] = E * HORIZON
    env.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    world, rank, local, dev = _dist_setup()
    from robosuite_benchmark_b200.parallel import max_over_ranks, shard, whole_job_rate
    cfg = CONFIGS[args.config]
    E = args.envs or cfg["envs"]
    env_id_base, _ = shard(rank, world, E)
    env = _make_env(cfg, E, dev, env_id_base)
    sim = env.sim
    obs = torch.zeros(E, sim.obs_dim, device=dev)
    rew = torch.zeros(E, device=dev)
    done = torch.zeros(E, dtype=torch.uint8, device=dev)
    act = torch.zeros(E, sim.act_dim, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    sim.reset(obs=obs)
    step_idx = 0

    def one_step():
        nonlocal step_idx
        sim.random_actions(step_idx, out=act)
        sim.step(act, obs, rew, done)
        step_idx += 1
        if step_idx % HORIZON == 0:
            sim.reset(obs=obs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pre-roll (untimed): the first ~50 control steps after a batch reset are ~1.6x cheaper than the rest of the episode (no arm has
    # reached the table yet, so no env needs more than 3 Newton iterations); start the timed region on the steady-state plateau
    for _ in range(PREROLL):
        one_step()
    for _ in range(max(3, args.warmup)):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sim.info("launches")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        sim.random_actions(step_idx, out=act)
        kev[k][0].record()
        sim.step(act, obs, rew, done)
        kev[k][1].record()
        step_idx += 1
        if step_idx % HORIZON == 0:
            sim.reset(obs=obs)
        ev[k][1].record()
    barrier()
    launches = sim.info("launches") - l0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    kms = sum(a.elapsed_time(b) for a, b in kev)
    rsum = float(rew.mean().item())
    ms_total = max_over_ranks(ms, dev)
    value = whole_job_rate(E * args.steps, world, ms_total / 1000.0)

    # ---- end-to-end through the host-buffer C-ABI call (pinned host actions in; obs/reward/done out), every step.
    # The host action batches continue the same Philox stream as the device-timed region (a constant action would drive the arms
    # into their limits and time a different workload); they are generated before the timed region and sit in pinned host memory.
    e2e_steps = max(5, min(args.steps, 50))
    h_acts = torch.empty(e2e_steps + 3, E, sim.act_dim, pin_memory=True)
    for k in range(e2e_steps + 3):
        sim.random_actions(step_idx + k, out=act)
        h_acts[k].copy_(act)
    torch.cuda.synchronize()
    h_np = h_acts.numpy()

    def host_step(k):
        nonlocal step_idx
        out = sim.step_host(h_np[k])
        step_idx += 1
        if step_idx % HORIZON == 0:
            sim.reset_host()
        return out

    for k in range(3):
        host_step(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        o_h, r_h, d_h = host_step(3 + k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = whole_job_rate(E * e2e_steps, world, max_over_ranks(e2e_s, dev))
    clocks = sampler.stop() if rank == 0 else None
    truncation = sim.counters()                       # contact / constraint-row truncation events of this rank: reported, never silent
    truncation["env_steps"] = E * step_idx
    if world > 1:
        tt = torch.tensor([truncation["ncon_overflow"], truncation["nefc_overflow"], truncation["env_steps"]], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        truncation = dict(ncon_overflow=int(tt[0]), nefc_overflow=int(tt[1]), steps_after_done=truncation["steps_after_done"], env_steps=int(tt[2]), over="all ranks")
    kinfo = {"regs": sim.info("regs_step"), "smem_bytes_per_env": sim.info("smem_bytes"), "envs_per_block": sim.info("envs_per_block"),
             "blocks_per_sm": sim.info("blocks_per_sm"), "lanes": sim.info("lanes"), "solver_option": list(sim.solver_option())}
    bps = algorithmic_bytes_per_step(env.task, env.model)
    obs_dim, act_dim = sim.obs_dim, sim.act_dim
    env.close()
    del env, sim

    extras = {}
    if not args.quick:
        # one whole episode of the same config (per GPU; the steady-state window above is what `value` reports)
        extras["full_episode"] = steady_state_rate(cfg, E, dev, env_id_base, steps=3, preroll=0, full_episode=True)["full_episode"]
        if args.config == "lift":
            others = {}
            for name in ("door", "stack", "twoarmlift"):
                c = CONFIGS[name]
                r = steady_state_rate(c, c["envs"], dev, shard(rank, world, c["envs"])[0], steps=20)
                r["steps_per_s_all_gpus"] = whole_job_rate(c["envs"], world, max_over_ranks(r["kernel_ms"], dev) / 1000.0)
                others[name] = r
            extras["other_configs"] = others
            tp = trained_policy_rate(cfg, E, dev, env_id_base)
            if tp is not None:
                tp["steps_per_s_all_gpus"] = whole_job_rate(E, world, max_over_ranks(tp["kernel_ms"], dev) / 1000.0)
                extras["trained_policy"] = tp
    sac = None
    if not args.no_sac:
        sac = sac_bench(dev, obs_dim, act_dim, world, rank)
    train = None
    if not args.no_train and not args.quick:
        train = train_loop_bench(cfg, dev, world, rank, envs=E, epochs=3, warm_epochs=1)
    if not args.quick and args.config == "lift" and world == 1:
        # the families added last (CONFIGS): measured after everything the headline line needs, single-GPU line only, and never allowed to take that line down
        more = {}
        for name in ("pickplacecan", "peginhole", "nutassemblyround", "handoff"):
            c = CONFIGS[name]
            try:
                more[name] = steady_state_rate(c, c["envs"], dev, 0, steps=20)
            except Exception as e:          # noqa: BLE001 -- reported in the line instead
                more[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        extras["more_families"] = more

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        kernel_ms = kms / args.steps
        achieved = bps * E / (kernel_ms / 1000.0) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(cfg)
        line = {"metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{family(cfg)}, {E} batched envs per GPU, tanh-Gaussian random actions (Philox), "
                                       f"horizon {HORIZON}, 25 substeps/control step, physics + controller + reward + obs",
                           "envs_per_gpu": E, "l2": "flushed between timed steps (256 MiB write)",
                           "episode_phase": f"timed region starts {PREROLL}+warmup control steps after the batch reset (steady state)",
                           "mean_reward_last_step": rsum},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * act_dim * 4,
                        "d2h_bytes_per_step": E * (obs_dim * 4 + 4 + 1)},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": measured_traffic(args.config) if E == cfg["envs"] else None, "kernel": "k_step", "kernel_ms": kernel_ms,
                             "algorithmic_bytes_per_env_step": bps,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                             "note": "state stays in shared memory for the 25 substeps: the kernel is bound by the dependent-instruction latency of its slowest "
                                     "environment, not by HBM (DESIGN.md 4.2); traffic = dram read+write of one launch from this round's ncu capture "
                                     "(profiles/r2_kstep_traffic.json), null when this config was not captured"},
                "kernel_info": kinfo, "truncation": truncation, "sac": sac, "train": train, "cpu_baseline": cpu}
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sac_bench(dev, obs_dim, act_dim, world, rank, updates=300):
    """SAC updates/s (the second half of BASELINE.json's metric): replay ring pre-filled with 1e6 synthetic transitions, Philox-sampled
    batches, 256x256 twin-Q + tanh-Gaussian policy, reference hyper-parameters; B = 128 (the reference's batch) and B = 4096."""
    import torch
    from robosuite_benchmark_b200 import gemm as _gemm
    from robosuite_benchmark_b200.backend import lib as backend_lib
    from robosuite_benchmark_b200.sac import EnvReplayBuffer, ParamStore, SACTrainer, algorithmic_flops_per_update
    n = 1_000_000
    rb = EnvReplayBuffer(n, obs_dim=obs_dim, action_dim=act_dim, device=dev, seed=SEED + rank)
    g = torch.Generator(device=dev); g.manual_seed(SEED + rank)
    for lo in range(0, n, 250_000):
        m = 250_000
        obs = torch.randn(m, obs_dim, device=dev, generator=g) * 0.5
        rb.add_batch(obs, torch.tanh(torch.randn(m, act_dim, device=dev, generator=g)), torch.rand(m, device=dev, generator=g) * 0.1,
                     torch.zeros(m, dtype=torch.uint8, device=dev), obs + 0.05 * torch.randn(m, obs_dim, device=dev, generator=g))
    out = {"gemm": "tcgen05 TF32, fp32 accumulate in tensor memory (csrc/rsb_tc_gemm.cu; bias/ReLU/ReLU-backward fused in the epilogue)",
           "ring_transitions": n, "world": world}
    arms = [(128, "tcgen05", "fused"), (4096, "tcgen05", "fused"), (128, "cublas", "fused"), (4096, "cublas", "fused")]
    if world > 1:
        arms.append((128, "tcgen05", "nccl"))           # comparison arm of the collective: graph -> dist.all_reduce -> graph
    out["allreduce"] = ("gradient mean fused into the optimizer kernel over NVLink symmetric memory (csrc/rsb_dp.cu)" if world > 1 else "none (one rank)")
    for B, gemm, ar in arms:
        store = ParamStore(obs_dim, act_dim, dev, seed=SEED, symmetric=(world > 1 and ar == "fused"))
        tr = SACTrainer(store=store, replay_buffer=rb, batch_size=B, discount=0.99, reward_scale=1.0, policy_lr=1e-3, qf_lr=5e-4,
                        soft_target_tau=0.005, target_update_period=5, seed=SEED, use_graph=True, world_size=world, rank=rank, gemm=gemm, allreduce=ar)
        for _ in range(10):
            tr.train_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(updates):
            tr.train_step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / updates
        fl = algorithmic_flops_per_update(obs_dim, act_dim, B)
        out[(f"b{B}" if gemm == "tcgen05" else f"b{B}_cublas_tf32") + ("_nccl_allreduce" if ar == "nccl" else "")] = {"updates_per_s": 1000.0 / ms, "samples_per_s": world * B * 1000.0 / ms, "us_per_update": 1000.0 * ms,
                        "algorithmic_gflop_per_update": fl / 1e9, "achieved_tflops": fl / (ms / 1000.0) / 1e12}
        del tr, store
    out["gemm_timeouts"] = _gemm.timeouts()
    out["dp_timeouts"] = int(backend_lib().rsb_dp_timeouts()) if world > 1 else 0
    assert out["gemm_timeouts"] == 0 and out["dp_timeouts"] == 0
    return out


def train_variant(cfg, envs, T, updates, batch, epochs):
    """The committed runs' variant (runs/*/variant.json: SAC 256x256, lr 1e-3 / 5e-4, tau 0.005, period 5, auto-entropy, replay 1e6) with the
    step counts scaled to a batch of `envs` environments: per epoch one evaluation round and one exploration round of T control steps each."""
    ek = dict(env_name=cfg["env"], robots=list(cfg["robots"]), horizon=HORIZON, control_freq=20, controller=cfg["controller"], reward_scale=1.0,
              hard_reset=False, ignore_done=True)
    return dict(algorithm="SAC", seed=SEED, version="normal", replay_buffer_size=1_000_000, qf_kwargs=dict(hidden_sizes=[256, 256]),
                policy_kwargs=dict(hidden_sizes=[256, 256]),
                algorithm_kwargs=dict(batch_size=batch, num_epochs=epochs, num_eval_steps_per_epoch=envs * T, num_expl_steps_per_train_loop=envs * T,
                                      num_trains_per_train_loop=updates, min_num_steps_before_training=envs * T, expl_max_path_length=T, eval_max_path_length=T),
                trainer_kwargs=dict(discount=0.99, policy_lr=1e-3, qf_lr=5e-4, reward_scale=1.0, soft_target_tau=0.005, target_update_period=5,
                                    use_automatic_entropy_tuning=True),
                expl_environment_kwargs=dict(ek), eval_environment_kwargs=dict(ek))


def train_loop_bench(cfg, dev, world, rank, envs, epochs, warm_epochs, T=25, updates=1000, batch=128):
    """The reference's training loop (util/rlkit_custom.py:215-239) on the batched backend, timed as a whole: per epoch an evaluation round and
    an exploration round of T control steps over `envs` envs per GPU (policy kernel -> step kernel -> replay ring, nothing else), add_paths
    (a pointer advance), `updates` SAC updates of `batch` rows sampled from that ring (gradient all-reduce over the ranks).  Device time by
    CUDA events at the epoch boundaries, max over ranks."""
    import torch
    from robosuite_benchmark_b200.algorithm import build_experiment
    from robosuite_benchmark_b200.parallel import max_over_ranks
    v = train_variant(cfg, envs * world, T, updates, batch, warm_epochs + epochs)
    algo = build_experiment(v, num_envs=envs, device=str(dev), seed=SEED, rank=rank, world_size=world)
    marks = []

    def mark(_algo, epoch):
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)

    algo.post_epoch_funcs.append(mark)
    l0 = [algo.expl_env.sim.info("launches"), algo.eval_env.sim.info("launches")]
    algo.train()
    torch.cuda.synchronize()
    ms = max_over_ranks(marks[warm_epochs - 1].elapsed_time(marks[-1]), dev) / epochs
    rows = algo.epoch_times[warm_epochs:]
    mean = lambda k: sum(r[k] for r in rows) / len(rows)
    env_steps = 2 * envs * T * world
    launches = (algo.expl_env.sim.info("launches") - l0[0] + algo.eval_env.sim.info("launches") - l0[1])
    out = {"workload": f"{family(cfg)} SAC training loop, {envs} exploration + {envs} evaluation envs per GPU, per epoch {T}+{T} control steps with per-step "
                       f"policy inference, {updates} updates of batch {batch} from the HBM replay ring",
           "world": world, "epochs_timed": epochs, "ms_per_epoch": ms,
           "env_steps_per_s": env_steps / ms * 1000.0, "updates_per_s_in_loop": updates / ms * 1000.0,
           "sampling_env_steps_per_s": env_steps / max(mean("time/exploration sampling (s)") + mean("time/evaluation sampling (s)"), 1e-9),
           "training_updates_per_s": updates / max(mean("time/training (s)"), 1e-9),
           "phase_s": {k.replace("time/", "").replace(" (s)", ""): mean(k) for k in rows[0]},
           "env_kernel_launches_total": int(launches), "eval_return_mean": algo.last_eval_info.get("Returns Mean"),
           "truncation": {"exploration": algo.expl_env.sim.counters(), "evaluation": algo.eval_env.sim.counters()}}
    algo.expl_env.close(); algo.eval_env.close()
    return out


def run_train(args):
    """--mode train: the training loop is the timed workload; one bench step = one epoch."""
    import torch.distributed as dist
    world, rank, local, dev = _dist_setup()
    cfg = CONFIGS[args.config]
    E = args.envs or cfg["envs"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = train_loop_bench(cfg, dev, world, rank, envs=E, epochs=args.steps, warm_epochs=max(1, args.warmup), T=args.train_steps_per_epoch,
                         updates=args.updates_per_epoch, batch=args.batch)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        line = {"metric": f"{family(cfg)} SAC training env control-steps/s (in-loop: policy inference + env.step + replay + SAC updates)",
                "value": r["env_steps_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup),
                "ms_per_step": r["ms_per_epoch"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 env / tf32 update",
                "data": "synthetic (random-init networks, transitions produced by the loop itself)", "config": {"workload": r["workload"], "envs_per_gpu": E,
                "updates_per_epoch": args.updates_per_epoch, "batch": args.batch, "control_steps_per_round": args.train_steps_per_epoch},
                "clocks": clocks, "gpu_launches": r["env_kernel_launches_total"], "train": r}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the config's BASELINE.json size)")
    ap.add_argument("--config", default="lift", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="step", choices=["step", "train"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sac", action="store_true", help="skip the SAC updates/s leg")
    ap.add_argument("--no-train", action="store_true", help="skip the short training-loop leg")
    ap.add_argument("--quick", action="store_true", help="only the timed region (no full-episode / other-config / training-loop legs)")
    ap.add_argument("--train-steps-per-epoch", type=int, default=25, help="--mode train: control steps per collection round")
    ap.add_argument("--updates-per-epoch", type=int, default=1000, help="--mode train: SAC updates per epoch (reference: 1000)")
    ap.add_argument("--batch", type=int, default=128, help="--mode train: SAC batch size (reference: 128)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "train":
        if args.steps == 100:
            args.steps = 10
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
