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
