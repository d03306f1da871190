#!/usr/bin/env python3
"""Substep-by-substep comparison CUDA vs host emulator for one policy-driven state (developer tool, run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from robosuite_benchmark_b200.policy_io import DeterministicPolicy
from robosuite_benchmark_b200.model import tasks
from robosuite_benchmark_b200 import controllers
from robosuite_benchmark_b200.backend import BatchSim
from oracle.oracle import OracleEnv
from tests.emu.emu import EmuEnv, split_debug
i = int(sys.argv[1]) if len(sys.argv) > 1 else 15
cc = controllers.load_controller_config(default_controller="OSC_POSE")
m, t = tasks.build_task("Lift", ["Panda"], cc)
d = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "policy_Lift-Panda-OSC-POSE-SEED17.npz"))); d.pop("logged")
pol = DeterministicPolicy({k: v.astype(np.float64) for k, v in d.items()})
sim = BatchSim(m, t, 2, device="cuda:0", seed=17, ncon_max=16, nefc_max=64)
e = EmuEnv(m, t, ncon_max=16, nefc_max=64, lanes=16)
orc = OracleEnv(m, t, ncon_max=16, nefc_max=64); o = orc.reset(seed=17, env_id=i, episode=0)
for k in range(20 + 6 * i): o, _, _ = orc.step(pol(np.asarray(o, np.float64)))
qpos, qvel, warm, cs = orc.get_state(); a = pol(np.asarray(o, np.float64))
row = sim.pack_state(qpos, qvel, warm, cs, timestep=k + 1, episode=1)[0]
sim.set_state(torch.as_tensor(np.stack([row, row]))); e.set_raw_state(row)
at = torch.as_tensor(np.stack([a, a]), dtype=torch.float32, device=sim.device)
for sub in range(25):
    g = split_debug(sim.debug_substep(at, sub == 0).cpu().numpy()[0], m.nv, 16, 64)
    h = e.debug_substep(a, sub == 0)
    same = g["contact_geoms"].tolist() == h["contact_geoms"].tolist()
    dn = np.abs(g["contact_frame"][:, :3] - h["contact_frame"][:, :3]).max() if same and g["ncon"] else -1
    dd = np.abs(g["contact_dist"] - h["contact_dist"]).max() if same and g["ncon"] else -1
    print(f"sub {sub:2d} ncon {g['ncon']}/{h['ncon']} nefc {g['nefc']}/{h['nefc']} iters {g['iters']}/{h['iters']} same_pairs {same} dnormal {dn:.2e} ddist {dd:.2e} "
          f"dqacc {np.abs(g['qacc']-h['qacc']).max():.2e} dqas {np.abs(g['qacc_smooth']-h['qacc_smooth']).max():.2e} dqfc {np.abs(g['qfrc_constraint']-h['qfrc_constraint']).max():.2e} dtau {np.abs(g['torques']-h['torques']).max():.2e}")
    if not same: print("   cuda", g["contact_geoms"].tolist(), "emu", h["contact_geoms"].tolist())
