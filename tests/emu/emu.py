"""ctypes front-end of tests/emu/librsb_emu.so -- TEST INFRASTRUCTURE (host emulation of the device code).

The product path never imports this: robosuite_benchmark_b200.backend loads only the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = None


def build(force=False, lanes=32):
    so = os.path.join(_HERE, "librsb_emu.so" if lanes == 32 else f"librsb_emu{lanes}.so")
    srcs = [os.path.join(_HERE, "rsb_emu.cpp"), os.path.join(_ROOT, "robosuite_benchmark_b200", "csrc", "rsb_dev.h"),
            os.path.join(_ROOT, "robosuite_benchmark_b200", "csrc", "rsb_devmodel.h"), os.path.join(_ROOT, "include", "rsb_model.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fno-strict-aliasing", "-Wall",
                               "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable", "-Wno-misleading-indentation",
                               f"-DRSB_LANES={lanes}", "-o", so, srcs[0]])
    return so


_LIBS = {}


def lib(lanes=32):
    if lanes not in _LIBS:
        L = C.CDLL(build(lanes=lanes))
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _LIBS[lanes] = L
    return _LIBS[lanes]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def split_debug(dbg, nv, ncon_max, nefc_max):
    """Unpack the debug record written by dump_debug (rsb_dev.h)."""
    out = {}
    ncon, nefc = int(dbg[0]), int(dbg[1])
    out["ncon"], out["nefc"], out["iters"] = ncon, nefc, int(dbg[2])
    o = 8
    out["M"] = dbg[o:o + nv * nv].reshape(nv, nv); o += nv * nv
    for k, name in enumerate(["qfrc_bias", "qfrc_passive", "qfrc_actuator", "qacc_smooth", "qacc", "qfrc_constraint",
                              "qfrc_smooth", "qacc_warmstart"]):
        out[name] = dbg[o + k * nv:o + (k + 1) * nv]
    o += 8 * nv
    out["torques"] = dbg[o:o + 14]; o += 14
    con = dbg[o:o + ncon_max * 16].reshape(ncon_max, 16)[:ncon]; o += ncon_max * 16
    out["contact_pos"], out["contact_frame"], out["contact_dist"] = con[:, :3], con[:, 3:12], con[:, 12]
    out["contact_geoms"] = con[:, 13:15].astype(int)
    out["contact_mu"] = con[:, 15]
    for k, name in enumerate(["efc_aref", "efc_R", "efc_force", "efc_pos", "efc_jar", "efc_type"]):
        out[name] = dbg[o + k * nefc_max:o + k * nefc_max + nefc]
    o += 6 * nefc_max
    out["efc_J"] = dbg[o:o + nefc_max * nv].reshape(nefc_max, nv)[:nefc]
    return out


class EmuEnv:
    def __init__(self, model, task, ncon_max=16, nefc_max=64, lanes=32):
        from robosuite_benchmark_b200.model.cstruct import model_to_c, task_to_c
        self.L = lib(lanes)
        self.model, self.task = model, task
        self._cm, self._keep = model_to_c(model)
        self._ct = task_to_c(task)
        self.h = self.L.emu_create(C.byref(self._cm), C.byref(self._ct), ncon_max, nefc_max)
        if not self.h:
            raise RuntimeError("emu_create failed")
        self.ncon_max, self.nefc_max = ncon_max, nefc_max
        self.nq, self.nv, self.nrobot = model.nq, model.nv, task["nrobot"]
        self.obs_dim, self.act_dim = task["obs_dim"], task["act_dim"]
        self.st_words = self.L.emu_state_words(C.c_void_p(self.h))
        self.dbg_words = self.L.emu_dbg_words(C.c_void_p(self.h))
        self.smem_words = self.L.emu_smem_words(C.c_void_p(self.h))

    def __del__(self):
        try:
            self.L.emu_destroy(C.c_void_p(self.h))
        except Exception:
            pass

    def set_order(self, order):
        self.L.emu_set_order(int(order))

    def set_solver(self, iters, ls_iters, tol):
        self.L.emu_set_solver(C.c_void_p(self.h), int(iters), int(ls_iters), C.c_float(tol))

    def solver_option(self):
        """(iterations, tolerance, ls_iterations, ls_tolerance) the device code runs with -- read from the model's <option>."""
        out = (C.c_double * 4)()
        self.L.emu_get_solver(C.c_void_p(self.h), out)
        return int(out[0]), float(out[1]), int(out[2]), float(out[3])

    def counters(self):
        """(ncon overflow, nefc overflow, steps after done) event counters of the device code since emu_create."""
        out = (C.c_uint * 3)()
        self.L.emu_counters(C.c_void_p(self.h), out)
        return tuple(int(v) for v in out)

    def raw_state(self):
        st = np.zeros(self.st_words, np.float32)
        self.L.emu_get_state(C.c_void_p(self.h), _p(st))
        return st

    def set_raw_state(self, st):
        st = np.ascontiguousarray(st, np.float32)
        self.L.emu_set_state(C.c_void_p(self.h), _p(st))

    def get_state(self):
        st = self.raw_state()
        nq, nv = self.nq, self.nv
        return st[:nq].copy(), st[nq:nq + nv].copy(), st[nq + nv:nq + 2 * nv].copy(), st[nq + 2 * nv:nq + 2 * nv + 80 * self.nrobot].copy()

    def set_state(self, qpos, qvel, warm=None, cs=None, timestep=0, episode=0, bpose=None):
        nq, nv = self.nq, self.nv
        st = self.raw_state()
        st[:nq], st[nq:nq + nv] = qpos, qvel
        st[nq + nv:nq + 2 * nv] = 0 if warm is None else warm
        if cs is not None:
            st[nq + 2 * nv:nq + 2 * nv + 80 * self.nrobot] = cs
        if bpose is not None:
            st[-9:-2] = bpose
        ints = st.view(np.int32)
        ints[-2], ints[-1] = timestep, episode
        self.set_raw_state(st)

    def reset(self, seed=0, env_id=0, episode=0):
        st = self.raw_state()
        st.view(np.int32)[-1] = episode
        self.set_raw_state(st)
        obs = np.zeros(self.obs_dim, np.float32)
        self.L.emu_reset(C.c_void_p(self.h), C.c_uint64(seed), C.c_uint64(env_id), _p(obs))
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, np.float32)
        obs = np.zeros(self.obs_dim, np.float32)
        rew = C.c_float(0)
        done = self.L.emu_step(C.c_void_p(self.h), _p(a), _p(obs), C.byref(rew))
        if done == 2:
            raise ValueError("executing action in terminated episode")
        return obs, rew.value, bool(done)

    def debug_substep(self, action, policy_step):
        a = np.ascontiguousarray(action, np.float32)
        dbg = np.zeros(self.dbg_words, np.float32)
        self.L.emu_debug_substep(C.c_void_p(self.h), _p(a), int(policy_step), _p(dbg))
        return split_debug(dbg, self.nv, self.ncon_max, self.nefc_max)

    def random_action(self, seed, env_id, step):
        a = np.zeros(self.act_dim, np.float32)
        self.L.emu_random_action(C.c_void_p(self.h), C.c_uint64(seed), C.c_uint64(env_id), C.c_uint64(step), _p(a))
        return a
