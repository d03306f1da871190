"""ctypes front-end of oracle/liboracle.so -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  PARITY UNPINNED against the real robosuite/MuJoCo (absent here): see rsb_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rsb_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "rsb_model.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_cost.restype = C.c_double
        for name in ("orc_destroy", "orc_forward", "orc_fwd_actuation", "orc_fwd_constraint", "orc_euler"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """One scalar environment stepping in float64."""

    def __init__(self, model, task, ncon_max: int = 0, nefc_max: int = 0):
        from robosuite_benchmark_b200.model.cstruct import model_to_c, task_to_c
        self.L = lib()
        self.model, self.task = model, task
        self._cm, self._keep = model_to_c(model)
        self._ct = task_to_c(task)
        assert self.L.orc_sizeof_model() == C.sizeof(self._cm), "rsb_model ABI mismatch"
        assert self.L.orc_sizeof_task() == C.sizeof(self._ct), "rsb_task ABI mismatch"
        self.h = self.L.orc_create(C.byref(self._cm), C.byref(self._ct), ncon_max)
        if not self.h:
            raise RuntimeError("orc_create failed (model exceeds oracle limits)")
        if nefc_max:
            self.L.orc_set_nefc_max(C.c_void_p(self.h), int(nefc_max))
        self.nq, self.nv, self.nu = model.nq, model.nv, model.nu
        self.obs_dim, self.act_dim, self.nrobot = task["obs_dim"], task["act_dim"], task["nrobot"]

    def __del__(self):
        try:
            self.L.orc_destroy(C.c_void_p(self.h))
        except Exception:
            pass

    # -- env API
    def reset(self, seed=0, env_id=0, episode=0):
        self.L.orc_reset(C.c_void_p(self.h), C.c_uint64(seed), C.c_uint64(env_id), C.c_uint64(episode))
        return self.observe()[0]

    def step(self, action):
        a = np.ascontiguousarray(action, np.float64)
        assert a.size == self.act_dim
        obs = np.zeros(self.obs_dim)
        rew = C.c_double(0)
        done = self.L.orc_step(C.c_void_p(self.h), _p(a), _p(obs), C.byref(rew))
        if done < 0:
            raise ValueError("executing action in terminated episode")
        return obs, rew.value, bool(done)

    def substep(self, action, policy_step):
        a = np.ascontiguousarray(action, np.float64)
        self.L.orc_substep(C.c_void_p(self.h), _p(a), int(policy_step))

    def observe(self):
        obs = np.zeros(self.obs_dim)
        rew = C.c_double(0)
        self.L.orc_observe(C.c_void_p(self.h), _p(obs), C.byref(rew))
        return obs, rew.value

    def random_action(self, seed, env_id, step):
        a = np.zeros(self.act_dim)
        self.L.orc_random_action(C.c_void_p(self.h), C.c_uint64(seed), C.c_uint64(env_id), C.c_uint64(step), _p(a))
        return a

    # -- state
    def get_state(self):
        qpos, qvel, warm = np.zeros(self.nq), np.zeros(self.nv), np.zeros(self.nv)
        cs = np.zeros(80 * self.nrobot)
        self.L.orc_get_state(C.c_void_p(self.h), _p(qpos), _p(qvel), _p(warm), _p(cs))
        return qpos, qvel, warm, cs

    def set_state(self, qpos, qvel, warm=None, cs=None):
        qpos = np.ascontiguousarray(qpos, np.float64)
        qvel = np.ascontiguousarray(qvel, np.float64)
        warm = None if warm is None else np.ascontiguousarray(warm, np.float64)
        cs = None if cs is None else np.ascontiguousarray(cs, np.float64)
        self.L.orc_set_state(C.c_void_p(self.h), _p(qpos), _p(qvel), None if warm is None else _p(warm),
                             None if cs is None else _p(cs))

    def get_bpose(self):
        b = np.zeros(7)
        self.L.orc_get_bpose(C.c_void_p(self.h), _p(b))
        return b

    def set_bpose(self, b):
        b = np.ascontiguousarray(b, np.float64)
        self.L.orc_set_bpose(C.c_void_p(self.h), _p(b))

    def set_timestep(self, t):
        self.L.orc_set_timestep(C.c_void_p(self.h), int(t))

    def set_ctrl(self, ctrl):
        c = np.ascontiguousarray(ctrl, np.float64)
        self.L.orc_set_ctrl(C.c_void_p(self.h), _p(c))

    # -- stages
    def forward(self):
        self.L.orc_forward(C.c_void_p(self.h))

    def fwd_actuation(self):
        self.L.orc_fwd_actuation(C.c_void_p(self.h))

    def fwd_constraint(self):
        self.L.orc_fwd_constraint(C.c_void_p(self.h))

    def euler(self):
        self.L.orc_euler(C.c_void_p(self.h))

    def cost(self, qacc):
        q = np.ascontiguousarray(qacc, np.float64)
        return self.L.orc_cost(C.c_void_p(self.h), _p(q))

    def get(self, name, shape=None):
        buf = np.zeros(640 * 40)
        n = self.L.orc_get(C.c_void_p(self.h), name.encode(), _p(buf))
        if n < 0:
            raise KeyError(name)
        out = buf[:n].copy()
        return out.reshape(shape) if shape is not None else out


def philox(seed, env_id, stream, index):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(C.c_uint64(seed), C.c_uint64(env_id), C.c_uint32(stream), C.c_uint32(index), out)
    return np.array(list(out), np.uint32)


def box_box(pa, Ra, ha, pb, Rb, hb, margin=0.0):
    out = np.zeros((8, 7))
    args = [np.ascontiguousarray(x, np.float64) for x in (pa, Ra, ha, pb, Rb, hb)]
    n = lib().orc_box_box(*[_p(a) for a in args], C.c_double(margin), _p(out))
    return out[:n]
