"""Read the reference's committed `params.pkl` snapshots without rlkit (§8f-2).

The files are legacy `torch.save` pickles of whole rlkit modules [REF util/rlkit_custom.py:54-82 `_get_snapshot`]: the classes
`rlkit.torch.sac.policies.{TanhGaussianPolicy, MakeDeterministic}` and `rlkit.torch.networks.FlattenMlp` do not exist here, so
unresolvable classes are replaced by stubs that keep their pickled state; the `torch.nn.Linear` layers inside load normally.
"""
from __future__ import annotations

import pickle
from typing import Dict

import numpy as np


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.__dict__["_state"] = s


class _Unpickler(pickle.Unpickler):
    def find_class(self, mod, name):
        try:
            return super().find_class(mod, name)
        except Exception:
            return type(name, (_Stub,), {"__module__": mod})


class _PickleModule:
    __name__ = "rsb_params_pkl"
    Unpickler = _Unpickler

    @staticmethod
    def load(f, **k):
        return _Unpickler(f, **k).load()


def load_params_pkl(path: str) -> Dict[str, object]:
    """The snapshot dict (`trainer/policy`, `trainer/qf1`, ..., `evaluation/policy`, ...) with stubbed rlkit modules."""
    import torch
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_PickleModule)


def mlp_weights(module) -> Dict[str, np.ndarray]:
    """`fc0/fc1/last_fc[/last_fc_log_std]` weights and biases of a stubbed rlkit Mlp / TanhGaussianPolicy as float64 arrays."""
    st = module._state
    if "stochastic_policy" in st:                     # MakeDeterministic wrapper
        st = st["stochastic_policy"]._state
    out = {}
    for name, lin in st["_modules"].items():
        if hasattr(lin, "weight"):
            out[name + ".weight"] = lin.weight.detach().cpu().numpy().astype(np.float64)
            out[name + ".bias"] = lin.bias.detach().cpu().numpy().astype(np.float64)
    return out


class DeterministicPolicy:
    """`MakeDeterministic(TanhGaussianPolicy)`: a = tanh(last_fc(relu(fc1(relu(fc0(obs))))))  [rlkit sac/policies.py]."""

    def __init__(self, weights: Dict[str, np.ndarray]):
        self.w = weights
        self.obs_dim = weights["fc0.weight"].shape[1]
        self.act_dim = weights["last_fc.weight"].shape[0]

    @classmethod
    def from_params_pkl(cls, path: str, key: str = "trainer/policy") -> "DeterministicPolicy":
        return cls(mlp_weights(load_params_pkl(path)[key]))

    def __call__(self, obs: np.ndarray) -> np.ndarray:
        w = self.w
        h = np.maximum(obs @ w["fc0.weight"].T + w["fc0.bias"], 0.0)
        k = 1
        while f"fc{k}.weight" in w:
            h = np.maximum(h @ w[f"fc{k}.weight"].T + w[f"fc{k}.bias"], 0.0)
            k += 1
        return np.tanh(h @ w["last_fc.weight"].T + w["last_fc.bias"])
