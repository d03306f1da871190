#!/usr/bin/env python3
"""Per-stage cycle profile of k_step from a -DRSB_PROFILE build (developer tool, run under gpurun).
Builds csrc/librsb_cuda_prof.so here (CPU), loads it instead of the product library, runs to a steady-state control step and prints,
per stage, the mean over warps and the slowest warp's / slowest CTA's cycles."""
import sys, os, subprocess, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from robosuite_benchmark_b200 import backend
PROF = os.path.join(os.path.dirname(backend.LIB_PATH), "librsb_cuda_prof.so")
def build():
    srcs = [os.path.join(os.path.dirname(backend.LIB_PATH), f) for f in ("rsb_cuda.cu", "rsb_cuda16.cu", "rsb_sac.cu", "rsb_tc_gemm.cu")]
    if not os.path.exists(PROF) or os.path.getmtime(PROF) < max(os.path.getmtime(s) for s in backend.sources()):
        subprocess.check_call(["nvcc"] + backend.NVCC_FLAGS + ["-DRSB_PROFILE", "-o", PROF] + srcs)
if len(sys.argv) > 1 and sys.argv[1] == "build":
    build(); sys.exit(0)
backend.LIB_PATH = PROF
import numpy as np, torch
import robosuite_benchmark_b200 as suite
E = 4096; at = int(sys.argv[1]) if len(sys.argv) > 1 else 150
dev = torch.device("cuda", 0)
cfg = suite.load_controller_config(default_controller="OSC_POSE")
env = suite.make("Lift", "Panda", controller_configs=cfg, num_envs=E, batched=True, device=dev, seed=17, horizon=500, control_freq=20, reward_shaping=True, ignore_done=True)
sim = env.sim; L = backend.lib()
obs = torch.zeros(E, sim.obs_dim, device=dev); rew = torch.zeros(E, device=dev); done = torch.zeros(E, dtype=torch.uint8, device=dev); act = torch.zeros(E, sim.act_dim, device=dev)
sim.reset(obs=obs)
names = ["kinematics", "inertia+crb", "collision", "bias", "controller", "actuation", "constraint", "solve (all)", "euler", "smooth_acc", " solve:warmstart", " solve:grad", " solve:hessian", " solve:cholesky", " solve:linesearch", "barrier wait"]
for k in range(at + 1):
    sim.random_actions(k, out=act)
    if k == at:
        torch.cuda.synchronize(); L.rsb_prof_reset()
    sim.step(act, obs, rew, done)
torch.cuda.synchronize()
epb = sim.info("envs_per_block"); wpb = epb // 2; nblk = (E + epb - 1) // epb; nw = nblk * wpb
buf = np.zeros(nw * 16, np.uint64); L.rsb_prof_read(buf.ctypes.data_as(C.c_void_p), buf.size)
cnt = (buf >> np.uint64(40)).reshape(nblk, wpb, 16).astype(np.float64); p = (buf & np.uint64((1 << 40) - 1)).reshape(nblk, wpb, 16).astype(np.float64)
tot = p[..., :10].sum(-1) + p[..., 15]      # slots 10-14 are a breakdown of slot 7
print(f"control step {at}: {nblk} CTAs x {wpb} warps; per-warp total cycles mean {tot.mean():.0f} max {tot.max():.0f} (= {tot.max()/1.965e6:.2f} ms at 1.965 GHz)")
worst = np.unravel_index(np.argmax(tot), tot.shape)[0]
print(f"{'stage':14s} {'mean/warp':>10s} {'share':>6s} {'max warp':>10s} {'slowest CTA (mean of its warps)':>32s}")
for i, nm in enumerate(names):
    if not nm: continue
    print(f"{nm:18s} {p[..., i].mean():10.0f} {100*p[..., i].mean()/tot.mean():5.1f}% {p[..., i].max():10.0f} {p[worst, :, i].mean():14.0f} {p[worst, :, i].max():14.0f}")
print("per-CTA total (max over its warps): min %.0f median %.0f p90 %.0f max %.0f" % tuple(np.percentile(tot.max(1), [0, 50, 90, 100])))
wi = np.unravel_index(np.argmax(p[..., 7]), p[..., 7].shape)
print(f"Newton iterations per warp per control step: mean {cnt[..., 12].mean():.1f} max {cnt[..., 12].max():.0f}; line-search evaluations: mean {cnt[..., 14].mean():.1f} max {cnt[..., 14].max():.0f}")
print(f"slowest-solve warp: {cnt[wi][12]:.0f} iterations, {cnt[wi][14]:.0f} line-search evaluations; cycles per iteration: grad {p[wi][11]/cnt[wi][12]:.0f} hessian {p[wi][12]/cnt[wi][12]:.0f} cholesky {p[wi][13]/cnt[wi][12]:.0f} linesearch {p[wi][14]/cnt[wi][12]:.0f} (per evaluation {p[wi][14]/max(cnt[wi][14],1):.0f})")
print(f"typical warp: cycles per iteration: grad {p[..., 11].sum()/cnt[..., 12].sum():.0f} hessian {p[..., 12].sum()/cnt[..., 12].sum():.0f} cholesky {p[..., 13].sum()/cnt[..., 12].sum():.0f} linesearch {p[..., 14].sum()/cnt[..., 12].sum():.0f} (per evaluation {p[..., 14].sum()/cnt[..., 14].sum():.0f})")
