"""Phase times of the fused all-reduce + Adam kernel (developer tool; torchrun with 2 or 8 ranks under gpurun)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
from robosuite_benchmark_b200.sac import EnvReplayBuffer, ParamStore, SACTrainer
from robosuite_benchmark_b200.backend import lib
world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
O, A = 42, 7
rb = EnvReplayBuffer(100000, obs_dim=O, action_dim=A, device=dev, seed=1 + rank)
obs = torch.randn(100000, O, device=dev) * 0.5
rb.add_batch(obs, torch.tanh(torch.randn(100000, A, device=dev)), torch.rand(100000, device=dev) * 0.1, torch.zeros(100000, dtype=torch.uint8, device=dev), obs)
tr = SACTrainer(store=ParamStore(O, A, dev, seed=1, symmetric=True), world_size=world, rank=rank, replay_buffer=rb, batch_size=128, discount=0.99, policy_lr=1e-3, qf_lr=5e-4,
                soft_target_tau=0.005, target_update_period=5, seed=1)
for _ in range(50):
    tr.train_step()
torch.cuda.synchronize(); dist.barrier()
acc = [0.0] * 6
N = 30
for _ in range(N):
    for _ in range(7):
        tr.train_step()
    clk = (C.c_ulonglong * 8)(); lib().rsb_dp_debug_clocks(clk)
    for i in range(6):
        acc[i] += (clk[i + 1] - clk[i]) / 1000.0
names = ["dependency wait", "signal READY", "wait peers READY", "go flag", "peer loads + Adam (+ CTA sync)", "tail (done signal)"]
print(f"rank {rank}/{world}: " + ", ".join(f"{n} {a / N:.2f} us" for n, a in zip(names, acc)) + f"; timeouts {tr.dp_timeouts()}", flush=True)
dist.barrier(); dist.destroy_process_group()
