#!/usr/bin/env python3
"""Kernel timeline of ONE SAC update (graph replay) from CUPTI activity records (torch.profiler): start offset, duration, stream of every kernel.
Shows where the update's dependent chain spends its time (developer tool, run under gpurun).   python tools/sac_timeline.py [B] [gemm]"""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from robosuite_benchmark_b200.sac import EnvReplayBuffer, ParamStore, SACTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gemm = sys.argv[2] if len(sys.argv) > 2 else "tcgen05"
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))); torch.cuda.set_device(dev); O, A = 42, 7
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
rb = EnvReplayBuffer(200000, obs_dim=O, action_dim=A, device=dev, seed=1)
g = torch.Generator(device=dev); g.manual_seed(1)
obs = torch.randn(200000, O, device=dev, generator=g) * 0.5
rb.add_batch(obs, torch.tanh(torch.randn(200000, A, device=dev, generator=g)), torch.rand(200000, device=dev, generator=g) * 0.1, torch.zeros(200000, dtype=torch.uint8, device=dev), obs)
tr = SACTrainer(store=ParamStore(O, A, dev, seed=1, symmetric=world > 1), world_size=world, rank=rank, replay_buffer=rb, batch_size=B, discount=0.99, policy_lr=1e-3, qf_lr=5e-4, soft_target_tau=0.005, target_update_period=5, seed=1, gemm=gemm)
for _ in range(20):
    tr.train_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        tr.train_step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
# split into updates at k_replay_sample_dev
starts = [i for i, e in enumerate(ev) if "k_sac_begin" in e.name or "replay_sample" in e.name]
import collections
print("kernel names:", dict(collections.Counter(e.name.replace("(anonymous namespace)::", "").split("(")[0][:28] for e in ev)))
if rank != 0:
    starts = []
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0)
if len(starts) >= 4:
    a, b = starts[2], starts[3]
    t0 = ev[a].time_range.start
    print(f"update of batch {B} ({gemm}): {b - a} kernels, span {(ev[b].time_range.start - t0):.1f} us (start of sampling to start of the next update's sampling)")
    for e in ev[a:b]:
        nm = e.name.replace("(anonymous namespace)::", "").split("(")[0][:28]
        print(f"  +{e.time_range.start - t0:7.1f} us  dur {e.time_range.end - e.time_range.start:6.1f} us  end {e.time_range.end - t0:6.1f}  {nm}")
else:
    print("could not split updates;", len(ev), "kernels")

if world > 1:
    dist.barrier(); dist.destroy_process_group()
