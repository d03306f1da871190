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
    out = bench.sac_bench(torch.device("cuda:0"), 42, 7, 1, 0, updates=a.updates)
    for k, v in out.items():
        print(k, json.dumps(v) if isinstance(v, dict) else v)
