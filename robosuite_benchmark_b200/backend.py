"""ctypes binding of the C-ABI in include/rsb.h (librsb_cuda.so) + torch-tensor plumbing.

This is the ONLY compute path of the package: there is no CPU fallback.  If the CUDA library is missing or no
CUDA device is visible, construction raises -- the oracle under oracle/ is test infrastructure and is never imported
from here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_CSRC, "librsb_cuda.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]
CU_SOURCES = ("rsb_cuda.cu", "rsb_cuda16.cu", "rsb_sac.cu", "rsb_sac_fused.cu", "rsb_dp.cu", "rsb_collect.cu", "rsb_tc_gemm.cu")
_LIB = None

INFO = dict(nenvs=0, obs_dim=1, act_dim=2, state_words=3, smem_bytes=4, dbg_words=5, nq=6, nv=7, envs_per_block=8,
            launches=9, ncon_max=10, nefc_max=11, regs_step=12, blocks_per_sm=13, lanes=14,
            ncon_overflow=15, nefc_overflow=16, steps_after_done=17, solver_iterations=18, ls_iterations=19)


def sources():
    return [os.path.join(_CSRC, f) for f in ("rsb_cuda.cu", "rsb_cuda16.cu", "rsb_kernels.inl", "rsb_ktable.h", "rsb_sac.cu", "rsb_sac_fused.cu", "rsb_dp.cu", "rsb_pdl.h", "rsb_collect.cu", "rsb_tc_gemm.cu", "rsb_dev.h", "rsb_devmodel.h")] + \
           [os.path.join(_HERE, "..", "include", f) for f in ("rsb.h", "rsb_sac.h", "rsb_gemm.h", "rsb_model.h")]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/rsb_cuda.cu for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    stale = force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in sources())
    if stale:
        cmd = ["nvcc", "--threads", "0"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + [os.path.join(_CSRC, f) for f in CU_SOURCES]
        subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the batched env.step path has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.rsb_last_error.restype = C.c_char_p
        L.rsb_info.restype = C.c_int64
        L.rsb_info.argtypes = [C.c_void_p, C.c_int]
        L.rsb_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        L.rsb_destroy.argtypes = [C.c_void_p]
        L.rsb_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rsb_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rsb_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rsb_reset_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rsb_random_actions.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.rsb_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rsb_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rsb_debug_substep.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.rsb_reset_ring.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.rsb_step_ring.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.rsb_get_iters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rsb_get_option.argtypes = [C.c_void_p, C.c_void_p]
        L.rsb_clear_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.rsb_sac_last_error.restype = C.c_char_p
        V, I, F, U64, L_ = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_long
        L.rsb_policy_act.argtypes = [V, V, V, V, V, V, I, I, I, V, L_, V, L_, C.c_int64, C.c_int64, I, I, U64, U64, U64, V]
        L.rsb_path_stats_words.argtypes = [I]
        L.rsb_path_stats.argtypes = [V, V, C.c_int64, C.c_int64, I, I, I, I, V, V]
        L.rsb_replay_sample.argtypes = [V, V, V, V, V, I, I, I, U64, U64, I, V, I, V, V, V, V, I, V, V]
        L.rsb_sac_prepare.argtypes = [V, V, V, V, V, I, V, I, I, I, V]
        L.rsb_normal.argtypes = [U64, U64, C.c_uint32, I, V, V]
        L.rsb_bias_relu.argtypes = [V, V, I, I, I, I, L_, I, V]
        L.rsb_relu_bwd.argtypes = [V, V, L_, V]
        L.rsb_colsum.argtypes = [V, I, I, I, V, I, L_, I, V]
        L.rsb_head_fwd.argtypes = [V, V, I, I, V, V, V, I, I, I, V, I, I, I, V]
        L.rsb_head_bwd.argtypes = [V, V, V, I, I, I, V, V, I, V, V]
        L.rsb_sac_losses.argtypes = [V, V, V, V, V, V, F, F, F, I, V, V, V, V, V]
        L.rsb_adam_polyak.argtypes = [V, V, V, V, L_, C.c_double, C.c_double, F, F, F, V, V, L_, L_, F, I, V, L_, I, V]
        L.rsb_adam_tick.argtypes = [V, F, F, V, V]
        L.rsb_policy_head_fwd.argtypes = [V, V, V, V, I, I, V, V, V, V, I, I, I, V, I, I, I, V]
        L.rsb_q_losses.argtypes = [V, V, V, V, V, V, V, V, V, V, F, F, F, I, V, V, V, V, V, V, V, V]
        L.rsb_policy_head_bwd.argtypes = [V, V, V, V, V, I, I, V, V, I, V, V, V]
        L.rsb_replay_sample_dev.argtypes = [V, V, V, V, V, V, I, I, U64, I, V, I, V, V, V, V, I, V, V]
        L.rsb_normal_dev.argtypes = [U64, V, C.c_uint32, I, V, V]
        L.rsb_counter_add.argtypes = [V, C.c_int64, V]
        L.rsb_dp_wait_peers_done.argtypes = [V, V, V, I, I, V]
        L.rsb_adam_polyak_allreduce.argtypes = [V, V, V, I, I, V, V, V, L_, C.c_double, C.c_double, F, F, F, V, V, L_, L_, F, I, V, L_, V]
        L.rsb_sac_begin.argtypes = [V, V, V, V, V, V, I, I, U64, I, I, V, I, V, V, V, V, V, V, I, V, I, V, I, U64, C.c_uint32, V, V, V, I, I, V]
        L.rsb_gemm_tf32.argtypes = [V, L_, L_, L_, V, L_, L_, L_, V, L_, L_, I, I, I, I, V, L_, V, L_, L_, I, I, I, L_, L_, V]
        L.rsb_gemm_debug_swap_offsets.argtypes = [I]
        L.rsb_gemm_debug_swap_offsets.restype = None
        L.rsb_gemm_debug_splits.argtypes = [I]
        L.rsb_gemm_debug_splits.restype = None
        L.rsb_gemm_debug_tma.argtypes = [I]
        L.rsb_gemm_debug_tma.restype = None
        L.rsb_gemm_debug_mn_swap.argtypes = [I]
        L.rsb_gemm_debug_mn_swap.restype = None
        _LIB = L
    return _LIB


EXPORTS = ["rsb_last_error", "rsb_sizeof_model", "rsb_sizeof_task", "rsb_create", "rsb_destroy", "rsb_info", "rsb_reset",
           "rsb_step", "rsb_step_host", "rsb_reset_host", "rsb_random_actions", "rsb_get_state", "rsb_set_state",
           "rsb_debug_substep", "rsb_reset_ring", "rsb_step_ring", "rsb_get_iters", "rsb_get_option", "rsb_clear_counters",
           "rsb_policy_act", "rsb_path_stats", "rsb_path_stats_words", "rsb_sac_last_error", "rsb_sac_prepare", "rsb_replay_sample", "rsb_normal", "rsb_bias_relu", "rsb_relu_bwd",
           "rsb_colsum", "rsb_head_fwd", "rsb_head_bwd", "rsb_sac_losses", "rsb_adam_polyak", "rsb_adam_tick", "rsb_policy_head_fwd", "rsb_q_losses",
           "rsb_policy_head_bwd", "rsb_replay_sample_dev", "rsb_normal_dev", "rsb_counter_add", "rsb_sac_begin", "rsb_dp_wait_peers_done", "rsb_adam_polyak_allreduce", "rsb_dp_timeouts", "rsb_dp_debug_clocks", "rsb_gemm_tf32", "rsb_gemm_timeouts",
           "rsb_gemm_debug_swap_offsets", "rsb_gemm_debug_clocks", "rsb_gemm_debug_splits", "rsb_gemm_plan", "rsb_gemm_debug_tma", "rsb_gemm_debug_last_tma", "rsb_gemm_debug_mn_swap"]


class RsbError(RuntimeError):
    pass


class RsbRing(C.Structure):
    """include/rsb.h `rsb_ring`: the five arrays of rlkit's EnvReplayBuffer, resident in HBM."""
    _fields_ = [("observations", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p), ("terminals", C.c_void_p),
                ("next_obs", C.c_void_p), ("capacity", C.c_int64)]


def _check(rc):
    if rc != 0:
        raise RsbError(lib().rsb_last_error().decode())


def _stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class BatchSim:
    """N environments of one compiled (model, task) on one GPU.  Tensors in, tensors out, stream-ordered on torch's current stream."""

    def __init__(self, model, task, num_envs: int, device="cuda:0", seed: int = 0, env_id_base: int = 0,
                 ncon_max: int = 0, nefc_max: int = 0):
        import torch
        from .model.cstruct import model_to_c, task_to_c
        self.L = lib()
        self.torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RsbError("the batched env.step path runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RsbError("no CUDA device visible: the batched env.step path has no CPU fallback")
        self.model, self.task = model, task
        self._cm, self._keep = model_to_c(model)
        self._ct = task_to_c(task)
        assert self.L.rsb_sizeof_model() == C.sizeof(self._cm), "rsb_model ABI mismatch"
        assert self.L.rsb_sizeof_task() == C.sizeof(self._ct), "rsb_task ABI mismatch"
        ncon_max = ncon_max or int(task.get("ncon_max", 0))
        nefc_max = nefc_max or int(task.get("nefc_max", 0))
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _check(self.L.rsb_create(C.byref(self._cm), C.byref(self._ct), int(num_envs), int(idx), C.c_uint64(seed),
                                 C.c_uint64(env_id_base), int(ncon_max), int(nefc_max), C.byref(h)))
        self.h = h
        self.device = torch.device("cuda", idx)
        self.num_envs = int(num_envs)
        self.obs_dim, self.act_dim = self.info("obs_dim"), self.info("act_dim")
        self.state_words, self.dbg_words = self.info("state_words"), self.info("dbg_words")
        self.nq, self.nv, self.nrobot = self.info("nq"), self.info("nv"), task["nrobot"]

    def info(self, what: str) -> int:
        return int(self.L.rsb_info(self.h, INFO[what]))

    def close(self):
        if getattr(self, "h", None):
            self.L.rsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- tensors
    def _f32(self, *shape):
        return self.torch.empty(*shape, dtype=self.torch.float32, device=self.device)

    def reset(self, mask=None, obs=None):
        t = self.torch
        if obs is None:
            obs = t.zeros(self.num_envs, self.obs_dim, dtype=t.float32, device=self.device)
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=t.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        _check(self.L.rsb_reset(self.h, mp, C.c_void_p(obs.data_ptr()), _stream_ptr(self.device)))
        return obs

    def step(self, actions, obs=None, reward=None, done=None):
        t = self.torch
        assert actions.is_cuda and actions.dtype == t.float32 and actions.is_contiguous() and tuple(actions.shape) == (self.num_envs, self.act_dim)
        obs = self._f32(self.num_envs, self.obs_dim) if obs is None else obs
        reward = self._f32(self.num_envs) if reward is None else reward
        done = t.empty(self.num_envs, dtype=t.uint8, device=self.device) if done is None else done
        _check(self.L.rsb_step(self.h, C.c_void_p(actions.data_ptr()), C.c_void_p(obs.data_ptr()), C.c_void_p(reward.data_ptr()),
                               C.c_void_p(done.data_ptr()), _stream_ptr(self.device)))
        return obs, reward, done

    # -- ring mode: the env kernels read actions from / write transitions into the replay ring (include/rsb.h rsb_step_ring)
    def reset_ring(self, ring: "RsbRing", slot0: int):
        _check(self.L.rsb_reset_ring(self.h, C.byref(ring), C.c_int64(slot0), _stream_ptr(self.device)))

    def step_ring(self, ring: "RsbRing", slot0: int, write_next_row: bool):
        _check(self.L.rsb_step_ring(self.h, C.byref(ring), C.c_int64(slot0), int(bool(write_next_row)), _stream_ptr(self.device)))

    def newton_iterations(self):
        """Newton iterations every env spent in its last control step (uint32 device tensor stored as int32)."""
        out = self.torch.empty(self.num_envs, dtype=self.torch.int32, device=self.device)
        _check(self.L.rsb_get_iters(self.h, C.c_void_p(out.data_ptr()), _stream_ptr(self.device)))
        return out

    def solver_option(self):
        """(iterations, tolerance, ls_iterations, ls_tolerance) the kernels run with: the model's <option>, as uploaded."""
        out = (C.c_double * 4)()
        _check(self.L.rsb_get_option(self.h, out))
        return int(out[0]), float(out[1]), int(out[2]), float(out[3])

    def counters(self):
        """Event counters since creation / clear_counters (SYNCHRONISES): contact truncations, constraint-row truncations, steps asked of
        terminated episodes.  Truncation is never silent: tests and bench assert the first two are 0."""
        return dict(ncon_overflow=self.info("ncon_overflow"), nefc_overflow=self.info("nefc_overflow"), steps_after_done=self.info("steps_after_done"))

    def clear_counters(self):
        _check(self.L.rsb_clear_counters(self.h, _stream_ptr(self.device)))

    def step_host(self, actions: np.ndarray):
        """The reference-facing call with HOST buffers (numpy in, numpy out): H2D + step + D2H inside."""
        a = np.ascontiguousarray(actions, np.float32).reshape(self.num_envs, self.act_dim)
        obs = np.empty((self.num_envs, self.obs_dim), np.float32)
        rew = np.empty(self.num_envs, np.float32)
        done = np.empty(self.num_envs, np.uint8)
        _check(self.L.rsb_step_host(self.h, a.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
                                    rew.ctypes.data_as(C.c_void_p), done.ctypes.data_as(C.c_void_p)))
        return obs, rew, done

    def reset_host(self, mask: Optional[np.ndarray] = None, obs: Optional[np.ndarray] = None):
        if obs is None:
            obs = np.zeros((self.num_envs, self.obs_dim), np.float32)
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            mp = mask.ctypes.data_as(C.c_void_p)
        _check(self.L.rsb_reset_host(self.h, mp, obs.ctypes.data_as(C.c_void_p)))
        return obs

    def random_actions(self, step: int, out=None):
        out = self._f32(self.num_envs, self.act_dim) if out is None else out
        _check(self.L.rsb_random_actions(self.h, C.c_uint64(step), C.c_void_p(out.data_ptr()), _stream_ptr(self.device)))
        return out

    def get_state(self):
        st = self._f32(self.num_envs, self.state_words)
        _check(self.L.rsb_get_state(self.h, C.c_void_p(st.data_ptr()), _stream_ptr(self.device)))
        return st

    def set_state(self, st):
        st = st.to(device=self.device, dtype=self.torch.float32).contiguous()
        assert tuple(st.shape) == (self.num_envs, self.state_words)
        _check(self.L.rsb_set_state(self.h, C.c_void_p(st.data_ptr()), _stream_ptr(self.device)))
        self.torch.cuda.current_stream(self.device).synchronize()   # `st` may be a temporary

    def debug_substep(self, actions, policy_step: bool):
        dbg = self.torch.zeros(self.num_envs, self.dbg_words, dtype=self.torch.float32, device=self.device)
        _check(self.L.rsb_debug_substep(self.h, C.c_void_p(actions.data_ptr()), int(policy_step), C.c_void_p(dbg.data_ptr()),
                                        _stream_ptr(self.device)))
        return dbg

    # -- state record helpers (layout: include/rsb.h rsb_get_state)
    def pack_state(self, qpos, qvel, warm=None, cs=None, timestep=0, episode=0, bpose=None) -> np.ndarray:
        """numpy rows [n, state_words] from per-env arrays (float64 accepted)."""
        qpos = np.atleast_2d(qpos)
        n = qpos.shape[0]
        st = np.zeros((n, self.state_words), np.float32)
        nq, nv = self.nq, self.nv
        st[:, :nq] = qpos
        st[:, nq:nq + nv] = np.atleast_2d(qvel)
        if warm is not None:
            st[:, nq + nv:nq + 2 * nv] = np.atleast_2d(warm)
        if cs is not None:
            st[:, nq + 2 * nv:nq + 2 * nv + 80 * self.nrobot] = np.atleast_2d(cs)
        o = nq + 2 * nv + 80 * self.nrobot
        st[:, o:o + 7] = [0, 0, 0, 1, 0, 0, 0] if bpose is None else np.atleast_2d(bpose)
        ints = st.view(np.int32)
        ints[:, -2] = timestep
        ints[:, -1] = episode
        return st

    def unpack_state(self, st: np.ndarray):
        nq, nv = self.nq, self.nv
        ints = st.view(np.int32)
        return dict(qpos=st[:, :nq], qvel=st[:, nq:nq + nv], warm=st[:, nq + nv:nq + 2 * nv],
                    cs=st[:, nq + 2 * nv:nq + 2 * nv + 80 * self.nrobot], bpose=st[:, -9:-2], timestep=ints[:, -2], episode=ints[:, -1])
