"""SAC updates/s for both GEMM providers (tcgen05 kernel / cuBLAS TF32) at the reference batch and a large batch -- bench.py's SAC leg alone.
    python tools/sac_rate.py [--updates 300]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--updates", type=int, default=300)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    out = bench.sac_bench(dev, 42, 7, world, rank, updates=a.updates)
    if rank == 0:
        for k, v in out.items():
            print(k, json.dumps(v) if isinstance(v, dict) else v)
    if world > 1:
        dist.destroy_process_group()
