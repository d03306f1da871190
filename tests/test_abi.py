"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsb_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from robosuite_benchmark_b200 import backend
    so = backend.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    declared = _declared("rsb.h") + _declared("rsb_sac.h") + _declared("rsb_gemm.h")
    assert len(declared) >= 20
    for name in declared:
        assert name in exported, name
    assert sorted(declared) == sorted(backend.EXPORTS)


def test_struct_sizes_match_header():
    from robosuite_benchmark_b200 import backend
    from robosuite_benchmark_b200.model.cstruct import RsbModel, RsbTask
    L = ctypes.CDLL(backend.build())
    assert L.rsb_sizeof_model() == ctypes.sizeof(RsbModel) and L.rsb_sizeof_task() == ctypes.sizeof(RsbTask)


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import robosuite_benchmark_b200 as suite
    from robosuite_benchmark_b200.backend import RsbError
    with pytest.raises(RsbError):
        suite.make("Lift", "Panda")
    from robosuite_benchmark_b200.sac import EnvReplayBuffer
    with pytest.raises(RsbError):
        EnvReplayBuffer(10, obs_dim=3, action_dim=2)


def test_gemm_launch_plan_invariants():
    """rsb_gemm_plan is the host arithmetic behind every tcgen05 product (tile width, split-K factor, stages, shared memory): no GPU needed."""
    from robosuite_benchmark_b200 import backend
    L = ctypes.CDLL(backend.build())
    plan = (ctypes.c_int * 8)()
    for m in (1, 42, 128, 130, 256, 4096, 8192):
        for n in (1, 14, 49, 256, 300):
            for k in (1, 49, 64, 128, 256, 512, 4096):
                for batch in (1, 2):
                    for nt, sp in ((0, 0), (16, 4), (64, 2), (128, 4), (128, 1), (32, 1)):
                        assert L.rsb_gemm_plan(m, n, k, batch, nt, sp, plan) == 0
                        n_tile, splits, cps, stages, recv_off, smem, gx, ctas = list(plan)
                        chunks = -(-k // 64)
                        assert n_tile in (16, 32, 64, 128) and splits in (1, 2, 4) and (nt == 0 or n_tile == nt)
                        assert splits * cps >= chunks and (splits - 1) * cps < chunks                  # the slices cover K, none is empty
                        assert 1 <= stages <= 4 and stages <= cps
                        stage = 128 * 64 * 4 + n_tile * 64 * 4
                        assert recv_off == stages * stage
                        panels = splits * 128 * (n_tile // splits + 4) * 4
                        assert smem >= 1024 + panels and smem >= 1024 + recv_off + (panels if splits > 1 else 0) and smem <= 232448
                        assert gx == -(-n // n_tile) * splits and ctas == gx * -(-m // 128) * batch
                        assert (n_tile // splits) % 4 == 0                                             # an owner's columns come in groups of 4
                        if nt == 0 and sp == 0 and splits > 1:
                            assert ctas <= 148                                                         # split-K only inside one wave
    assert L.rsb_gemm_plan(0, 1, 1, 1, 0, 0, plan) != 0 and L.rsb_gemm_plan(8, 8, 8, 1, 48, 0, plan) != 0
