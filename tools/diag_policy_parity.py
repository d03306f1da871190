#!/usr/bin/env python3
"""Per-env deviations CUDA vs oracle vs host emulator on policy-driven states (developer tool, run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from robosuite_benchmark_b200.policy_io import DeterministicPolicy
from robosuite_benchmark_b200.model import tasks
from robosuite_benchmark_b200 import controllers
from robosuite_benchmark_b200.backend import BatchSim
from oracle.oracle import OracleEnv
from tests.emu.emu import EmuEnv
cc = controllers.load_controller_config(default_controller="OSC_POSE")
m, t = tasks.build_task("Lift", ["Panda"], cc)
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "policy_Lift-Panda-OSC-POSE-SEED17.npz"))); d.pop("logged")
pol = DeterministicPolicy({k: v.astype(np.float64) for k, v in d.items()})
n = 32
sim = BatchSim(m, t, n, device="cuda:0", seed=17, ncon_max=16, nefc_max=64)
e = EmuEnv(m, t, ncon_max=16, nefc_max=64, lanes=16)
rows, acts, envs = [], [], []
for i in range(n):
    orc = OracleEnv(m, t, ncon_max=16, nefc_max=64); o = orc.reset(seed=17, env_id=i, episode=0)
    for k in range(20 + 6 * i): o, _, _ = orc.step(pol(np.asarray(o, np.float64)))
    qpos, qvel, warm, cs = orc.get_state()
    rows.append(sim.pack_state(qpos, qvel, warm, cs, timestep=k + 1, episode=1)[0]); acts.append(pol(np.asarray(o, np.float64))); envs.append(orc)
sim.set_state(torch.as_tensor(np.stack(rows)))
a = torch.as_tensor(np.stack(acts), dtype=torch.float32, device=sim.device)
obs, rew, done = sim.step(a)
st = sim.unpack_state(sim.get_state().cpu().numpy())
for i, orc in enumerate(envs):
    e.set_raw_state(rows[i]); e.step(acts[i]); q2, v2, _, _ = e.get_state()
    orc.step(acts[i]); q1, v1, _, _ = orc.get_state()
    print(i, "cuda-oracle dv %.2e dq %.2e | emu-oracle dv %.2e | cuda-emu dv %.2e  argmax %d ncon %d" % (np.abs(v1 - st["qvel"][i]).max(), np.abs(q1 - st["qpos"][i]).max(),
          np.abs(v1 - v2).max(), np.abs(v2 - st["qvel"][i]).max(), int(np.abs(v1 - st["qvel"][i]).argmax()), int(orc.get("counts")[0])))
