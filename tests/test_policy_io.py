"""params.pkl importer: shapes and the deterministic tanh-MLP against a torch restatement."""
import glob
import os

import numpy as np
import pytest

from robosuite_benchmark_b200.policy_io import DeterministicPolicy, load_params_pkl, mlp_weights

RUN = "/root/reference/runs/Lift-Panda-OSC-POSE-SEED17"


@pytest.mark.skipif(not os.path.isdir(RUN), reason="the reference's committed runs are only present in the build container")
def test_load_committed_lift_policy():
    import torch
    path = glob.glob(RUN + "/*/params.pkl")[0]
    snap = load_params_pkl(path)
    for key in ("trainer/policy", "trainer/qf1", "trainer/qf2", "trainer/target_qf1", "trainer/target_qf2", "evaluation/policy", "exploration/policy"):
        assert key in snap                                                      # the 7 keys of _get_snapshot [REF util/rlkit_custom.py:68-82]
    w = mlp_weights(snap["trainer/policy"])
    assert w["fc0.weight"].shape == (256, 42) and w["fc1.weight"].shape == (256, 256) and w["last_fc.weight"].shape == (7, 256)
    assert w["last_fc_log_std.weight"].shape == (7, 256)
    q = mlp_weights(snap["trainer/qf1"])
    assert q["fc0.weight"].shape == (256, 49) and q["last_fc.weight"].shape == (1, 256)
    det = mlp_weights(snap["evaluation/policy"])                                 # MakeDeterministic wraps the same network
    assert np.array_equal(det["fc0.weight"], w["fc0.weight"])
    pol = DeterministicPolicy(w)
    obs = np.random.default_rng(0).normal(size=(5, 42))
    lin = snap["trainer/policy"]._state["_modules"]
    with torch.no_grad():
        x = torch.tensor(obs, dtype=torch.float32)
        ref = torch.tanh(lin["last_fc"](torch.relu(lin["fc1"](torch.relu(lin["fc0"](x)))))).numpy()
    assert np.abs(pol(obs) - ref).max() < 1e-5
