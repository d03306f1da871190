"""rlkit's SAC pieces as the reference wires them (util/rlkit_utils.py:64-106,139-150; util/rlkit_custom.py:233-240),
re-built on the CUDA library: `EnvReplayBuffer` (HBM-resident ring, Philox-indexed sampling), `FlattenMlp`,
`TanhGaussianPolicy`, `MakeDeterministic`, `SACTrainer`.

All trainable parameters live in ONE flat fp32 buffer  [policy | Q1,Q2 (stacked per layer) | log_alpha]  with a matching flat
gradient buffer, so that (a) the twin Q networks run as batched GEMMs, (b) one kernel does the four Adam steps and the
Polyak update, (c) data-parallel training all-reduces a single bucket.  The dense products run on the hand-written tcgen05 TF32 kernel
(csrc/rsb_tc_gemm.cu via gemm.gemm_tf32, bias / ReLU / ReLU-backward in its epilogue; `gemm="cublas"` keeps torch.mm / bmm as the
comparison arm and as the fp32 strict-parity mode); everything else is the fused kernels of csrc/rsb_sac.cu.  One whole update is captured in a CUDA graph.
There is no CPU path: constructing any of these without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np

from .backend import RsbError, lib
from .gemm import gemm_tf32

HID = 256


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _chk(rc):
    if rc != 0:
        raise RsbError(lib().rsb_sac_last_error().decode())


def _stream(dev):
    import torch
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# ----------------------------------------------------------------------------- replay ring
class EnvReplayBuffer:
    """rlkit EnvReplayBuffer layout (observations, actions, rewards, terminals, next_obs; ring pointer `_top`, `_size`)
    resident in HBM.  `add_batch` appends N transitions written by the env kernels; `random_batch` draws indices with
    replacement from Philox keyed (seed, draw counter, row) -- the same integers the numpy oracle produces."""

    def __init__(self, max_replay_buffer_size, env=None, env_info_sizes=None, obs_dim=None, action_dim=None, device=None, seed=0):
        import torch
        if not torch.cuda.is_available():
            raise RsbError("the replay ring lives in HBM: no CUDA device visible")
        self.torch, self.L = torch, lib()
        self.device = torch.device(device or default_device())
        if env is not None:
            obs_dim = env.observation_space.low.size
            action_dim = env.action_space.low.size
        self.obs_dim, self.action_dim = int(obs_dim), int(action_dim)
        self.capacity = int(max_replay_buffer_size)
        f32 = dict(dtype=torch.float32, device=self.device)
        self._observations = torch.zeros(self.capacity, self.obs_dim, **f32)
        self._next_obs = torch.zeros(self.capacity, self.obs_dim, **f32)
        self._actions = torch.zeros(self.capacity, self.action_dim, **f32)
        self._rewards = torch.zeros(self.capacity, **f32)
        self._terminals = torch.zeros(self.capacity, dtype=torch.uint8, device=self.device)
        self._top, self._size, self.seed, self.draws = 0, 0, int(seed), 0
        from .backend import RsbRing
        # the C-ABI view of the five arrays (include/rsb.h rsb_ring): in ring mode the env kernels are the producer of these rows
        self.ring = RsbRing(self._observations.data_ptr(), self._actions.data_ptr(), self._rewards.data_ptr(), self._terminals.data_ptr(),
                            self._next_obs.data_ptr(), self.capacity)

    @property
    def top(self):
        return self._top

    def commit(self, n):
        """The n rows from `top` on were written in place by the env kernels (rsb_step_ring): advance the ring pointer, saturate the
        size -- rlkit's add_sample bookkeeping (`_advance`) for n samples, with no data movement."""
        if n > self.capacity:
            raise ValueError("round larger than the ring")
        self._advance(int(n))

    def _pieces(self, n):
        """Ring rows for the next n transitions as (ring_slice, batch_slice) pieces (two when the pointer wraps)."""
        n = int(n)
        if n > self.capacity:
            raise ValueError("batch larger than the ring")
        first = min(n, self.capacity - self._top)
        out = [(slice(self._top, self._top + first), slice(0, first))]
        if first < n:
            out.append((slice(0, n - first), slice(first, n)))
        return out

    def _advance(self, n):
        self._top = (self._top + n) % self.capacity
        self._size = min(self.capacity, self._size + n)

    def add_batch(self, obs, actions, rewards, terminals, next_obs):
        """Append n transitions (device tensors).  rlkit add_sample semantics applied row by row: ring pointer advances, size saturates."""
        n = obs.shape[0]
        rewards, terminals = rewards.reshape(-1), terminals.reshape(-1)
        for rs, bs in self._pieces(n):
            self._observations[rs].copy_(obs[bs])
            self._actions[rs].copy_(actions[bs])
            self._rewards[rs].copy_(rewards[bs])
            self._terminals[rs].copy_(terminals[bs])
            self._next_obs[rs].copy_(next_obs[bs])
        self._advance(n)

    def add_paths(self, paths):
        """rlkit path dicts (numpy, [T, .]) -- the single-env protocol."""
        t = self.torch
        for p in paths:
            self.add_batch(t.as_tensor(p["observations"], dtype=t.float32, device=self.device),
                           t.as_tensor(p["actions"], dtype=t.float32, device=self.device),
                           t.as_tensor(p["rewards"], dtype=t.float32, device=self.device),
                           t.as_tensor(np.asarray(p["terminals"]).astype(np.uint8), device=self.device),
                           t.as_tensor(p["next_observations"], dtype=t.float32, device=self.device))

    def num_steps_can_sample(self):
        return self._size

    def sample_into(self, batch_size, step, b_obs, ld_obs, b_act, b_rew, b_term, b_next, ld_next, b_idx=None):
        _chk(self.L.rsb_replay_sample(_ptr(self._observations), _ptr(self._actions), _ptr(self._rewards), _ptr(self._terminals),
                                      _ptr(self._next_obs), self._size, self.obs_dim, self.action_dim, C.c_uint64(self.seed),
                                      C.c_uint64(step), int(batch_size), _ptr(b_obs), int(ld_obs), _ptr(b_act), _ptr(b_rew), _ptr(b_term),
                                      _ptr(b_next), int(ld_next), None if b_idx is None else _ptr(b_idx), _stream(self.device)))

    def random_batch(self, batch_size):
        t = self.torch
        f32 = dict(dtype=t.float32, device=self.device)
        b = dict(observations=t.empty(batch_size, self.obs_dim, **f32), actions=t.empty(batch_size, self.action_dim, **f32),
                 rewards=t.empty(batch_size, **f32), terminals=t.empty(batch_size, **f32),
                 next_observations=t.empty(batch_size, self.obs_dim, **f32),
                 indices=t.empty(batch_size, dtype=t.int32, device=self.device))
        self.sample_into(batch_size, self.draws, b["observations"], self.obs_dim, b["actions"], b["rewards"], b["terminals"],
                         b["next_observations"], self.obs_dim, b["indices"])
        self.draws += 1
        b["rewards"], b["terminals"] = b["rewards"].unsqueeze(1), b["terminals"].unsqueeze(1)
        return b

    def get_diagnostics(self):
        return OrderedDict([("size", self._size)])

    def get_snapshot(self):
        return {}

    def end_epoch(self, epoch):
        return


def replay_indices_reference(seed, step, batch, size):
    """Host restatement of the index rule (numpy) -- used by tests; mirrors oracle/sac_oracle.py."""
    from .philox import philox4x32
    out = np.empty(batch, np.int64)
    for b in range(batch):
        w = philox4x32([b, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, 0xB0FFE7], [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
        out[b] = (int(w[0]) * int(size)) >> 32
    return out


# ----------------------------------------------------------------------------- parameter store + network handles
class ParamStore:
    """Flat fp32 parameter / gradient / Adam-moment buffers and the views the GEMMs read."""

    def __init__(self, obs_dim, act_dim, device, hidden=HID, seed=0, policy_init_w=1e-3, qf_init_w=3e-3, b_init=0.1, symmetric=False):
        """symmetric: allocate the gradient bucket in NVLink symmetric memory (data-parallel runs: every rank then reads every rank's bucket
        directly, csrc/rsb_dp.cu); needs an initialised NCCL process group."""
        import torch
        self.torch, self.device = torch, torch.device(device)
        self.symmetric = bool(symmetric)
        self.O, self.A, self.H = int(obs_dim), int(act_dim), int(hidden)
        O, A, H, QI = self.O, self.A, self.H, self.O + self.A
        shapes = OrderedDict([
            ("p_W0", (O, H)), ("p_b0", (H,)), ("p_W1", (H, H)), ("p_b1", (H,)), ("p_W2", (H, 2 * A)), ("p_b2", (2 * A,)),
            ("q_W0", (2, QI, H)), ("q_b0", (2, H)), ("q_W1", (2, H, H)), ("q_b1", (2, H)), ("q_W2", (2, H, 1)), ("q_b2", (2, 1)),
            ("log_alpha", (1,))])
        self.shapes, self.offsets, n = shapes, OrderedDict(), 0
        for k, s in shapes.items():
            n = (n + 3) & ~3                            # every tensor starts on a 16-byte boundary: TMA-stageable as a GEMM operand (the gaps stay zero)
            self.offsets[k] = n
            n += int(np.prod(s))
        self.n = (n + 3) & ~3
        self.q_begin, self.q_end = self.offsets["q_W0"], self.offsets["log_alpha"]
        f32 = dict(dtype=torch.float32, device=self.device)
        self.flat = torch.zeros(n, **f32)
        if self.symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.grad = symm_mem.empty(n, dtype=torch.float32, device=self.device)
            self.grad.zero_()
        else:
            self.grad = torch.zeros(n, **f32)
        self.m, self.v = torch.zeros(n, **f32), torch.zeros(n, **f32)
        self.target = torch.zeros(self.q_end - self.q_begin, **f32)
        self.P, self.G, self.T = self._views(self.flat, 0), self._views(self.grad, 0), self._views(self.target, self.q_begin, only_q=True)
        self.init_weights(seed, policy_init_w, qf_init_w, b_init)

    def _views(self, buf, base, only_q=False):
        v = {}
        for k, s in self.shapes.items():
            if only_q and not k.startswith("q_"):
                continue
            o = self.offsets[k] - base
            v[k] = buf[o:o + int(np.prod(s))].view(*s)
        return v

    def init_weights(self, seed, policy_init_w, qf_init_w, b_init):
        self.load_host(init_host_params(self.O, self.A, self.H, seed, policy_init_w, qf_init_w, b_init))

    def load_host(self, host, targets=None):
        t = self.torch
        for k, a in host.items():
            self.P[k].copy_(t.as_tensor(np.asarray(a, np.float32).reshape(self.shapes[k])))
        if targets is None:
            self.target.copy_(self.flat[self.q_begin:self.q_end])
        else:
            for k, a in targets.items():
                self.T[k].copy_(t.as_tensor(np.asarray(a, np.float32).reshape(self.shapes[k])))

    def to_host(self):
        return {k: v.detach().cpu().numpy().copy() for k, v in self.P.items()}, {k: v.detach().cpu().numpy().copy() for k, v in self.T.items()}


def init_host_params(O, A, H=HID, seed=0, policy_init_w=1e-3, qf_init_w=3e-3, b_init=0.1):
    """rlkit Mlp init: hidden W ~ U(+-1/sqrt(fan_in)) (fanin_init), hidden b = 0.1, last layers U(+-init_w).
    Host numpy, in the store's layout (weights [in, out], twin Q stacked on axis 0)."""
    rng = np.random.default_rng(seed)
    QI = O + A

    def u(shape, bound):
        return rng.uniform(-bound, bound, size=shape).astype(np.float32)

    return {"p_W0": u((O, H), 1 / np.sqrt(O)), "p_b0": np.full(H, b_init, np.float32), "p_W1": u((H, H), 1 / np.sqrt(H)),
            "p_b1": np.full(H, b_init, np.float32), "p_W2": u((H, 2 * A), policy_init_w), "p_b2": u((2 * A,), policy_init_w),
            "q_W0": u((2, QI, H), 1 / np.sqrt(QI)), "q_b0": np.full((2, H), b_init, np.float32), "q_W1": u((2, H, H), 1 / np.sqrt(H)),
            "q_b1": np.full((2, H), b_init, np.float32), "q_W2": u((2, H, 1), qf_init_w), "q_b2": u((2, 1), qf_init_w),
            "log_alpha": np.zeros(1, np.float32)}


def default_device():
    """cuda:LOCAL_RANK (one process per GPU under torchrun), cuda:0 otherwise."""
    import os
    return "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))


def _fanin_uniform(shape, fan_in):
    bound = 1.0 / np.sqrt(fan_in)
    return np.random.uniform(-bound, bound, size=shape).astype(np.float32)


class _Net:
    """Common part of the network handles.  A handle is either BOUND to a ParamStore (its weights are views into the store's flat
    buffer: the form the fused update needs) or UNBOUND (host numpy weights: what the reference's constructors produce before the
    trainer exists, util/rlkit_utils.py:64-92, and what a snapshot unpickles to); SACTrainer binds unbound handles into one store."""
    KEYS = ("W0", "b0", "W1", "b1", "W2", "b2")
    store = None

    def to(self, device=None):                       # rlkit: net.to(ptu.device) -- parameters already live where the store lives
        return self

    def cuda(self, device=None):
        return self

    def cpu(self):
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def host_params(self):
        """name -> numpy [in, out] weights / [out] biases."""
        if self.store is None:
            return {k: v.copy() for k, v in self._host.items()}
        return {k: v.detach().cpu().numpy().copy() for k, v in self.views().items()}

    def _torch_state(self, host):
        import torch
        return OrderedDict((k, torch.from_numpy(np.ascontiguousarray(v))) for k, v in host.items())


class FlattenMlp(_Net):
    """rlkit FlattenMlp: cat(inputs, dim=1) -> 256 -> 256 -> output_size, with the reference's constructor
    `FlattenMlp(input_size=obs_dim + action_dim, output_size=1, hidden_sizes=[256, 256])` (util/rlkit_utils.py:64-83).
    rlkit init: hidden W ~ U(+-1/sqrt(fan_in)), hidden b = 0.1, last layer U(+-init_w), drawn from numpy's global generator."""

    def __init__(self, hidden_sizes=(HID, HID), output_size=1, input_size=None, init_w=3e-3, b_init_value=0.1, **unsupported):
        if unsupported:
            raise NotImplementedError(f"FlattenMlp options {sorted(unsupported)} are not on the benchmark path")
        if list(hidden_sizes) != [HID, HID] or output_size != 1 or input_size is None:
            raise NotImplementedError("the fused update is built for the benchmark's Q networks: input -> 256 -> 256 -> 1")
        self.input_size, self.output_size, self.hidden_sizes = int(input_size), 1, [HID, HID]
        self.store, self.index, self.target = None, None, False
        QI = self.input_size
        self._host = {"W0": _fanin_uniform((QI, HID), QI), "b0": np.full(HID, b_init_value, np.float32),
                      "W1": _fanin_uniform((HID, HID), HID), "b1": np.full(HID, b_init_value, np.float32),
                      "W2": np.random.uniform(-init_w, init_w, (HID, 1)).astype(np.float32),
                      "b2": np.random.uniform(-init_w, init_w, (1,)).astype(np.float32)}

    @classmethod
    def of(cls, store: "ParamStore", index: int, target=False):
        """Handle on Q network `index` of an existing store (its target copy when target=True)."""
        self = cls.__new__(cls)
        self.store, self.index, self.target, self._host = store, int(index), bool(target), None
        self.input_size, self.output_size, self.hidden_sizes = store.O + store.A, 1, [store.H, store.H]
        return self

    def views(self):
        src = self.store.T if self.target else self.store.P
        return {k: src["q_" + k][self.index] for k in self.KEYS}

    def __call__(self, *inputs):
        """Q(obs, act) on the device (diagnostics / tests; the update itself runs the batched twin-Q products of SACTrainer)."""
        if self.store is None:
            raise RsbError("this Q network is not bound to a trainer yet (SACTrainer(...) places it in the parameter store)")
        t = self.store.torch
        v, out = self.views(), None
        x = t.cat([t.as_tensor(i, dtype=t.float32, device=self.store.device) for i in inputs], dim=1).contiguous()
        h1 = t.empty(x.shape[0], HID, device=x.device); h2 = t.empty_like(h1); out = t.empty(x.shape[0], 1, device=x.device)
        gemm_tf32(x, v["W0"], h1, bias=v["b0"], relu=True); gemm_tf32(h1, v["W1"], h2, bias=v["b1"], relu=True); gemm_tf32(h2, v["W2"], out, bias=v["b2"])
        return out

    def state_dict(self):
        v = self.host_params()
        return self._torch_state(OrderedDict([("fc0.weight", v["W0"].T), ("fc0.bias", v["b0"]), ("fc1.weight", v["W1"].T), ("fc1.bias", v["b1"]),
                                              ("last_fc.weight", v["W2"].T), ("last_fc.bias", v["b2"])]))

    def __getstate__(self):                          # snapshots hold the weights, not the store
        return dict(input_size=self.input_size, params=self._torch_state(self.host_params()))

    def __setstate__(self, st):
        self.input_size, self.output_size, self.hidden_sizes = int(st["input_size"]), 1, [HID, HID]
        self.store, self.index, self.target = None, None, False
        self._host = {k: v.numpy().copy() for k, v in st["params"].items()}


class TanhGaussianPolicy(_Net):
    """rlkit TanhGaussianPolicy: obs -> 256 -> 256 -> (mean, log_std); a = tanh(mean + std * eps), with the reference's constructor
    `TanhGaussianPolicy(obs_dim=, action_dim=, hidden_sizes=[256, 256])` (util/rlkit_utils.py:88-92).  get_action / get_actions run
    the hand-written batched forward kernel (csrc/rsb_collect.cu k_policy_act) -- there is no torch/cuBLAS forward and no CPU path."""

    def __init__(self, hidden_sizes=(HID, HID), obs_dim=None, action_dim=None, std=None, init_w=1e-3, b_init_value=0.1, **unsupported):
        if unsupported or std is not None:
            raise NotImplementedError("TanhGaussianPolicy: only the learned-std form of the benchmark is implemented")
        if list(hidden_sizes) != [HID, HID] or obs_dim is None or action_dim is None:
            raise NotImplementedError("the policy kernels are built for the benchmark's policy: obs -> 256 -> 256 -> 2 * action_dim")
        self.obs_dim, self.action_dim = int(obs_dim), int(action_dim)
        self.input_size, self.output_size, self.hidden_sizes = self.obs_dim, self.action_dim, [HID, HID]
        self.store, self._dev, self.seed, self.env_id_base, self._steps = None, None, 0, 0, 0
        O, A = self.obs_dim, self.action_dim
        self._host = {"W0": _fanin_uniform((O, HID), O), "b0": np.full(HID, b_init_value, np.float32),
                      "W1": _fanin_uniform((HID, HID), HID), "b1": np.full(HID, b_init_value, np.float32),
                      "W2": np.random.uniform(-init_w, init_w, (HID, 2 * A)).astype(np.float32),
                      "b2": np.random.uniform(-init_w, init_w, (2 * A,)).astype(np.float32)}

    @classmethod
    def of(cls, store: "ParamStore", seed=0, env_id_base=0):
        self = cls.__new__(cls)
        self.store, self._host, self._dev = store, None, None
        self.obs_dim, self.action_dim = store.O, store.A
        self.input_size, self.output_size, self.hidden_sizes = store.O, store.A, [store.H, store.H]
        self.seed, self.env_id_base, self._steps = int(seed), int(env_id_base), 0
        return self

    def views(self):
        return {k: self.store.P["p_" + k] for k in self.KEYS}

    # -- device weights for the forward kernel
    def _weights(self):
        if self.store is not None:
            return self.views(), self.store.device
        if self._dev is None:                        # an unbound policy (e.g. unpickled from params.pkl) uploads its weights on first use
            import torch
            if not torch.cuda.is_available():
                raise RsbError("the policy forward runs on a CUDA device only (no CPU fallback)")
            dev = torch.device(default_device())
            self._dev = ({k: torch.as_tensor(v, device=dev).contiguous() for k, v in self._host.items()}, dev)
        return self._dev

    def act_into(self, obs, obs_ld, act, act_ld, n, slot0=0, cap=0, deterministic=False, step=None, stream_device=None):
        """Launch k_policy_act on raw row-addressed arrays (the replay ring in ring mode).  `step` keys the exploration noise
        (default: this policy's own call counter)."""
        W, dev = self._weights()
        if step is None:
            step = self._steps
            self._steps += 1
        _chk(lib().rsb_policy_act(_ptr(W["W0"]), _ptr(W["b0"]), _ptr(W["W1"]), _ptr(W["b1"]), _ptr(W["W2"]), _ptr(W["b2"]),
                                  self.obs_dim, self.action_dim, HID, obs, int(obs_ld), act, int(act_ld), C.c_int64(slot0), C.c_int64(cap), int(n),
                                  int(bool(deterministic)), C.c_uint64(self.seed), C.c_uint64(self.env_id_base), C.c_uint64(step), _stream(dev)))

    def get_actions(self, obs, deterministic=False, out=None, step=None):
        """obs: fp32 CUDA tensor [N, obs_dim] -> actions [N, action_dim] (env i of the batch draws noise under global id env_id_base + i)."""
        _, dev = self._weights()
        import torch
        obs = obs.to(device=dev, dtype=torch.float32)
        if obs.stride(-1) != 1:
            obs = obs.contiguous()
        n = obs.shape[0]
        out = torch.empty(n, self.action_dim, dtype=torch.float32, device=dev) if out is None else out
        self.act_into(_ptr(obs), obs.stride(0), _ptr(out), out.stride(0), n, deterministic=deterministic, step=step)
        return out

    def get_action(self, obs_np, deterministic=False):
        """rlkit Policy.get_action: one observation (numpy) -> (action float64 numpy, agent_info {})."""
        import torch
        _, dev = self._weights()
        o = torch.as_tensor(np.asarray(obs_np, np.float32).reshape(1, -1), device=dev)
        return self.get_actions(o, deterministic)[0].cpu().numpy().astype(np.float64), {}

    def reset(self):
        pass

    def state_dict(self):
        v, A = self.host_params(), self.action_dim
        return self._torch_state(OrderedDict([("fc0.weight", v["W0"].T), ("fc0.bias", v["b0"]), ("fc1.weight", v["W1"].T), ("fc1.bias", v["b1"]),
                                              ("last_fc.weight", v["W2"][:, :A].T), ("last_fc.bias", v["b2"][:A]),
                                              ("last_fc_log_std.weight", v["W2"][:, A:].T), ("last_fc_log_std.bias", v["b2"][A:])]))

    def __getstate__(self):
        return dict(obs_dim=self.obs_dim, action_dim=self.action_dim, seed=self.seed, params=self._torch_state(self.host_params()))

    def __setstate__(self, st):
        self.obs_dim, self.action_dim = int(st["obs_dim"]), int(st["action_dim"])
        self.input_size, self.output_size, self.hidden_sizes = self.obs_dim, self.action_dim, [HID, HID]
        self.store, self._dev, self.seed, self.env_id_base, self._steps = None, None, int(st.get("seed", 0)), 0, 0
        self._host = {k: v.numpy().copy() for k, v in st["params"].items()}


class MakeDeterministic(_Net):
    """rlkit MakeDeterministic: the evaluation policy, a = tanh(mean)."""

    def __init__(self, stochastic_policy):
        self.stochastic_policy = stochastic_policy

    def get_action(self, obs_np):
        return self.stochastic_policy.get_action(obs_np, deterministic=True)

    def get_actions(self, obs, deterministic=True, out=None, step=None):
        return self.stochastic_policy.get_actions(obs, deterministic=True, out=out, step=step)

    def act_into(self, *a, **k):
        k["deterministic"] = True
        return self.stochastic_policy.act_into(*a, **k)

    def reset(self):
        pass

    def state_dict(self):
        return self.stochastic_policy.state_dict()

    def cuda(self, device=None):
        self.stochastic_policy.cuda(device)
        return self


def register_safe_globals():
    """torch >= 2.6 unpickles with weights_only=True by default: allow-list the snapshot classes so that the reference's
    `torch.load(params.pkl)['evaluation/policy']` (util/rlkit_utils.py:173-174,241-242) keeps working on snapshots written here."""
    try:
        import torch
        torch.serialization.add_safe_globals([FlattenMlp, TanhGaussianPolicy, MakeDeterministic, OrderedDict])
    except Exception:
        pass


# ----------------------------------------------------------------------------- trainer
class SACTrainer:
    """rlkit SACTrainer (twin Q, no V net, automatic entropy tuning) -- update order and semantics per SURVEY.md A.4:
    all losses are built from the pre-update weights and the PRE-update alpha; Adam on policy / Q1 / Q2 / log_alpha;
    Polyak update when n_train_steps % target_update_period == 0 (so the very first step already updates)."""

    def __init__(self, env=None, policy=None, qf1=None, qf2=None, target_qf1=None, target_qf2=None, discount=0.99, reward_scale=1.0,
                 policy_lr=1e-3, qf_lr=1e-3, optimizer_class=None, soft_target_tau=1e-2, target_update_period=1, plotter=None,
                 render_eval_paths=False, use_automatic_entropy_tuning=True, target_entropy=None, *, store: ParamStore = None,
                 replay_buffer: EnvReplayBuffer = None, batch_size=None, seed=0, use_graph=True, world_size=1, rank=0,
                 parallel_branches=True, gemm="tcgen05", device=None, allreduce="fused"):
        """Positional / keyword arguments up to `target_entropy` are rlkit's SACTrainer signature as the reference calls it
        (util/rlkit_utils.py:98-106); the keyword-only ones belong to this backend.  Networks built with the reference's constructors
        (unbound handles) are placed into ONE flat parameter store here; `store=` passes an existing one."""
        import torch
        if optimizer_class is not None:
            raise NotImplementedError("the fused optimizer kernel implements torch.optim.Adam (rlkit's default); optimizer_class is not supported")
        if store is None and policy is not None and policy.store is not None:
            store = policy.store
        if store is None:
            if policy is None:
                raise ValueError("SACTrainer needs a policy (or store=)")
            O, A = policy.obs_dim, policy.action_dim
            store = ParamStore(O, A, device or default_device(), seed=seed)
            host, targets = {}, {}
            for k, v in policy.host_params().items():
                host["p_" + k] = v
            q1, q2 = (qf1 or FlattenMlp(input_size=O + A)).host_params(), (qf2 or FlattenMlp(input_size=O + A)).host_params()
            # rlkit's target networks are separately initialised modules (the reference builds four FlattenMlp's) and are NOT copied from
            # the Q networks before training: keep that
            t1 = (target_qf1 or FlattenMlp(input_size=O + A)).host_params()
            t2 = (target_qf2 or FlattenMlp(input_size=O + A)).host_params()
            for k in _Net.KEYS:
                host["q_" + k] = np.stack([q1[k], q2[k]])
                targets["q_" + k] = np.stack([t1[k], t2[k]])
            store.load_host(host, targets)
            for net, idx, tgt in ((qf1, 0, False), (qf2, 1, False), (target_qf1, 0, True), (target_qf2, 1, True)):
                if net is not None:
                    net.store, net.index, net.target, net._host = store, idx, tgt, None
            policy.store, policy._host, policy._dev = store, None, None
        self.torch, self.L, self.store, self.device = torch, lib(), store, store.device
        self.policy = policy or TanhGaussianPolicy.of(store, seed=seed)
        self.qf1, self.qf2 = qf1 or FlattenMlp.of(store, 0), qf2 or FlattenMlp.of(store, 1)
        self.target_qf1, self.target_qf2 = target_qf1 or FlattenMlp.of(store, 0, True), target_qf2 or FlattenMlp.of(store, 1, True)
        self.env, self.replay = env, replay_buffer
        self.discount, self.reward_scale = float(discount), float(reward_scale)
        self.tau, self.period = float(soft_target_tau), int(target_update_period)
        if not use_automatic_entropy_tuning:
            raise NotImplementedError("fixed-alpha SAC is not on the benchmark path (all committed variants use automatic entropy tuning)")
        A = store.A
        self.target_entropy = float(-A if target_entropy is None else target_entropy)   # -prod(action_space.shape)
        self.seed, self.use_graph, self.world, self.rank = int(seed), bool(use_graph), int(world_size), int(rank)
        # GEMM back-end of the update, always explicit: "tcgen05" = the hand-written TF32 tensor-core kernel (csrc/rsb_tc_gemm.cu, the product
        # path); "cublas" = torch.mm/bmm with TF32 (comparison arm); "cublas_fp32" = torch.mm/bmm in fp32 (reference arm of the parity tests)
        if gemm not in ("tcgen05", "cublas", "cublas_fp32"):
            raise ValueError(f"gemm must be 'tcgen05', 'cublas' or 'cublas_fp32', got {gemm!r}")
        self.gemm, self.tf32 = gemm, gemm != "cublas_fp32"
        # exploration-noise key of the update's reparameterised actions: the rank is folded in, so data-parallel ranks draw independent eps
        self.noise_stream = 7 + 16 * self.rank
        # the one collective of the path (world > 1): "fused" = peer-memory reduction inside the optimizer kernel (csrc/rsb_dp.cu; the gradient bucket must
        # be symmetric memory: ParamStore(symmetric=True)), "nccl" = dist.all_reduce between the gradient graph and the optimizer graph (comparison arm)
        if allreduce not in ("fused", "nccl"):
            raise ValueError(f"allreduce must be 'fused' or 'nccl', got {allreduce!r}")
        self.allreduce = allreduce if self.world > 1 else "none"
        self._dp = None
        if self.allreduce == "fused":
            if not getattr(store, "symmetric", False):
                raise RsbError("allreduce='fused' needs ParamStore(symmetric=True) (gradient bucket in NVLink symmetric memory)")
            self._dp = self._setup_dp()
        self._n_train_steps_total = 0
        self._need_to_update_eval_statistics = True
        self.eval_statistics = OrderedDict()
        f32 = dict(dtype=torch.float32, device=self.device)
        self.policy_lr, self.qf_lr = float(policy_lr), float(qf_lr)
        self.bc = torch.tensor([0.0, 0.0, 1.0, 1.0], dtype=torch.float64, device=self.device)
        self.alpha = torch.tensor([1.0, 0.0], **f32)                 # [alpha, log_alpha] (device copy read by the kernels)
        self.refresh_alpha()
        self.ctr = torch.zeros(2, dtype=torch.int64, device=self.device)     # device-resident [filled ring rows, update counter]
        self._ring_size_seen = -1
        self.B = None
        if batch_size is not None:
            self.set_batch_size(batch_size)
        self._graphs = {}
        # side streams: the target-Q forward and the weight-gradient GEMMs do not lie on the update's dependency chain; inside the captured
        # graph they become parallel branches (the update is launch-latency bound: ~55 kernels of 2-4 us each)
        self._sT, self._sB = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self._sWs = [torch.cuda.Stream(self.device) for _ in range(3)]
        self.parallel_branches = bool(parallel_branches)          # False: everything on one stream (the reference order; tests compare the two)

    def _setup_dp(self):
        """Rendezvous of the symmetric buffers: every rank learns the device addresses of every rank's gradient bucket and flag words."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        try:
            symm_mem.enable_symm_mem_for_group(dist.group.WORLD.group_name)      # needed on older torch; a no-op / deprecated on newer
        except Exception:
            pass
        flags = symm_mem.empty(64, dtype=torch.int32, device=self.device)
        flags.zero_()
        hg = symm_mem.rendezvous(self.store.grad, dist.group.WORLD.group_name)
        hf = symm_mem.rendezvous(flags, dist.group.WORLD.group_name)
        local = torch.zeros(4, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier()                                 # every rank's flag words are zero before anyone signals
        return dict(flags=flags, hg=hg, hf=hf, local=local, grads_dev=C.c_void_p(hg.buffer_ptrs_dev), flags_dev=C.c_void_p(hf.buffer_ptrs_dev))

    def dp_timeouts(self):
        return int(self.L.rsb_dp_timeouts()) if self._dp is not None else 0

    def refresh_alpha(self):
        """[alpha, log_alpha] device copy from the store's log_alpha (after load_host / a checkpoint restore)."""
        la = self.store.P["log_alpha"]
        self.alpha[1:2].copy_(la)
        self.alpha[0:1].copy_(la.double().exp().float())

    def set_batch_size(self, B):
        """(Re)allocate the update's activation buffers for batches of B rows; captured graphs of another size are dropped."""
        if self.B != int(B):
            self.B = int(B)
            self._alloc(self.B)
            self._graphs = {}

    # -- buffers
    def _alloc(self, B):
        t, s = self.torch, self.store
        O, A, H, QI = s.O, s.A, s.H, s.O + s.A
        f32 = dict(dtype=t.float32, device=self.device)
        Op, QIp = (O + 3) & ~3, (QI + 3) & ~3          # row pitches padded to 16 bytes: the first layers' inputs are TMA-stageable (csrc/rsb_tc_gemm.cu)
        self.ldO, self.ldQ = Op, QIp
        self.Xp = t.zeros(2 * B, Op, **f32)[:, :O]     # policy input: obs rows, then next_obs rows
        self.XQ = t.zeros(2 * B, QIp, **f32)[:, :QI]   # Q input: (obs, a_new) rows, then (obs, act) rows
        self.XT = t.zeros(B, QIp, **f32)[:, :QI]       # target-Q input: (next_obs, a')
        self.act, self.rew, self.term = t.zeros(B, A, **f32), t.zeros(B, **f32), t.zeros(B, **f32)
        self.idx = t.zeros(B, dtype=t.int32, device=self.device)
        self.eps = t.zeros(2 * B, A, **f32)
        self.H1p, self.H2p, self.OUT = t.zeros(2 * B, H, **f32), t.zeros(2 * B, H, **f32), t.zeros(2 * B, 2 * A, **f32)
        self.a_store, self.logpi = t.zeros(2 * B, A, **f32), t.zeros(2 * B, **f32)
        self.H1q, self.H2q, self.q = t.zeros(2, 2 * B, H, **f32), t.zeros(2, 2 * B, H, **f32), t.zeros(2, 2 * B, 1, **f32)
        self.H1t, self.H2t, self.qt = t.zeros(2, B, H, **f32), t.zeros(2, B, H, **f32), t.zeros(2, B, 1, **f32)
        self.dq, self.y, self.sums = t.zeros(2, 2 * B, 1, **f32), t.zeros(B, **f32), t.zeros(8, **f32)
        self.dH2q, self.dH1q = t.zeros(2, 2 * B, H, **f32), t.zeros(2, 2 * B, H, **f32)
        self.gX = t.zeros(B, QIp, **f32)[:, :QI]
        self.dOUT, self.dH2p, self.dH1p = t.zeros(2 * B, 2 * A, **f32), t.zeros(B, H, **f32), t.zeros(B, H, **f32)

    # -- kernels
    def _bias_relu(self, x, bias, relu, nmat=1):
        rows, cols = x.shape[-2], x.shape[-1]
        _chk(self.L.rsb_bias_relu(_ptr(x), _ptr(bias), rows, cols, int(relu), nmat, rows * cols, cols, _stream(self.device)))

    def _relu_bwd(self, dy, y):
        _chk(self.L.rsb_relu_bwd(_ptr(dy), _ptr(y), dy.numel(), _stream(self.device)))

    def _colsum(self, dy, r0, r1, db, nmat=1):
        rows, cols = dy.shape[-2], dy.shape[-1]
        _chk(self.L.rsb_colsum(_ptr(dy), r0, r1, cols, _ptr(db), nmat, rows * cols, cols, _stream(self.device)))

    def _mm(self, a, b, out, bias=None, relu=False, mask=None, accumulate=False):
        """out = epilogue(a @ b) (2-D, or 3-D batched over the twin networks): ONE launch of the tcgen05 TF32 kernel with the bias / ReLU /
        ReLU-backward mask / accumulation in its epilogue (gemm="tcgen05"), or the cuBLAS product followed by the elementwise kernels
        (gemm="cublas": fp32 when tf32=False -- the mode the strict oracle-parity test runs in)."""
        if self.gemm == "tcgen05":
            return gemm_tf32(a, b, out, bias=bias, relu=relu, mask=mask, accumulate=accumulate)
        t = self.torch
        if accumulate:
            out.addmm_(a, b)
        elif a.dim() == 3:
            t.bmm(a, b, out=out)
        else:
            t.mm(a, b, out=out)
        if bias is not None:
            self._bias_relu(out, bias, int(relu), out.shape[0] if out.dim() == 3 else 1)
        if mask is not None:
            self._relu_bwd(out, mask)

    def _sync_ring_size(self):
        """ctr[0] = filled rows of the replay ring (read by the in-graph sampling kernel); refreshed only when the ring grew."""
        n = self.replay._size if self.replay is not None else 0
        if n != self._ring_size_seen:
            self.ctr[0:1].fill_(n)
            self._ring_size_seen = n

    def _sample(self):
        """EnvReplayBuffer.random_batch (util/rlkit_custom.py:235) straight into the update's input buffers: row b takes ring row
        mulhi(Philox(seed; b, update counter), size) -- size and counter are read from device memory (graph-capturable)."""
        s, R = self.store, self.replay
        B, O = self.B, s.O
        _chk(self.L.rsb_replay_sample_dev(_ptr(R._observations), _ptr(R._actions), _ptr(R._rewards), _ptr(R._terminals), _ptr(R._next_obs), _ptr(self.ctr),
                                          R.obs_dim, R.action_dim, C.c_uint64(R.seed), B, _ptr(self.Xp), self.ldO, _ptr(self.act), _ptr(self.rew), _ptr(self.term),
                                          C.c_void_p(self.Xp.data_ptr() + 4 * B * self.ldO), self.ldO, _ptr(self.idx), _stream(self.device)))

    def load_batch(self, batch):
        """Use an explicit batch (dict of tensors/arrays, rlkit keys) instead of sampling -- parity tests and `train(batch)`."""
        t = self.torch
        self.set_batch_size(len(batch["rewards"]))
        B = self.B
        g = lambda k: t.as_tensor(np.asarray(batch[k]) if not t.is_tensor(batch[k]) else batch[k], dtype=t.float32, device=self.device)
        self.Xp[:B].copy_(g("observations")); self.Xp[B:].copy_(g("next_observations"))
        self.act.copy_(g("actions")); self.rew.copy_(g("rewards").reshape(-1)); self.term.copy_(g("terminals").reshape(-1))

    def _update_body(self, sample, noise, tick_early):
        """Gradients of one update into the flat gradient buffer.  sample / noise: draw the batch / the policy noise here (device counters);
        tick_early: advance Adam's bias corrections and the update counter on a side stream (the single-graph form: `_apply(tick=False)` follows)."""
        t, s, L, B = self.torch, self.store, self.L, self.B
        O, A, H, QI = s.O, s.A, s.H, s.O + s.A
        P, G, T = s.P, s.G, s.T
        st = _stream(self.device)
        main = t.cuda.current_stream(self.device)
        sT, sB = (self._sT, self._sB) if self.parallel_branches else (main, main)
        sWs = self._sWs if self.parallel_branches else [main]
        self._wi = 0
        # head of the update in one kernel: batch rows from the ring (or the explicit batch already loaded), the update's input layout, cleared
        # accumulators, policy noise -- all keyed by the device-resident counters
        R, dp = self.replay, self._dp                # (data parallel: this kernel also waits until the peers are done with the previous gradient bucket)
        ring = (_ptr(R._observations), _ptr(R._actions), _ptr(R._rewards), _ptr(R._terminals), _ptr(R._next_obs), C.c_uint64(R.seed)) if sample else (None,) * 5 + (C.c_uint64(0),)
        _chk(L.rsb_sac_begin(ring[0], ring[1], ring[2], ring[3], ring[4], _ptr(self.ctr), O, A, ring[5], B, int(sample), _ptr(self.Xp), self.ldO, _ptr(self.act), _ptr(self.rew),
                             _ptr(self.term), _ptr(self.idx), _ptr(self.XQ), _ptr(self.XT), self.ldQ, _ptr(self.sums), self.sums.numel(), _ptr(G["log_alpha"]), int(noise),
                             C.c_uint64(self.seed), self.noise_stream, _ptr(self.eps),
                             dp["flags_dev"] if dp else None, _ptr(dp["local"]) if dp else None, self.rank, self.world if dp else 0, st))
        if tick_early:                               # after the kernel that reads the update counter; nothing else of the update reads what it writes before _apply
            sB.wait_stream(main)
            with t.cuda.stream(sB):
                _chk(L.rsb_adam_tick(_ptr(self.bc), 0.9, 0.999, C.c_void_p(self.ctr.data_ptr() + 8), _stream(self.device)))
        # policy forward on [obs; next_obs]: two tensor-core layers, then last layer + tanh-Gaussian head in one kernel; the actions go straight into
        # the action columns of the Q inputs (rows [0, B) -> XQ, rows [B, 2B) -> XT)
        mm = self._mm
        mm(self.Xp, P["p_W0"], self.H1p, bias=P["p_b0"], relu=True)
        mm(self.H1p, P["p_W1"], self.H2p, bias=P["p_b1"], relu=True)
        _chk(L.rsb_policy_head_fwd(_ptr(self.H2p), _ptr(P["p_W2"]), _ptr(P["p_b2"]), _ptr(self.eps), 2 * B, A, _ptr(self.OUT), _ptr(self.a_store), _ptr(self.logpi),
                                   C.c_void_p(self.XQ.data_ptr() + 4 * O), self.ldQ, 0, B, C.c_void_p(self.XT.data_ptr() + 4 * O), self.ldQ, B, 2 * B, st))
        # twin Q hidden layers (batched over the two networks) on [(obs,a_new); (obs,act)]; the target nets' on (next_obs, a') on a side stream
        XQ2, XT2 = self.XQ.unsqueeze(0).expand(2, 2 * B, QI), self.XT.unsqueeze(0).expand(2, B, QI)
        sT.wait_stream(main)
        with t.cuda.stream(sT):
            mm(XT2, T["q_W0"], self.H1t, bias=T["q_b0"], relu=True)
            mm(self.H1t, T["q_W1"], self.H2t, bias=T["q_b1"], relu=True)
        mm(XQ2, P["q_W0"], self.H1q, bias=P["q_b0"], relu=True)
        mm(self.H1q, P["q_W1"], self.H2q, bias=P["q_b1"], relu=True)
        main.wait_stream(sT)
        # last Q layers (dot products), TD target, losses, dq and the gradient w.r.t. the second hidden layer: one kernel
        _chk(L.rsb_q_losses(_ptr(self.H2q), _ptr(P["q_W2"]), _ptr(P["q_b2"]), _ptr(self.H2t), _ptr(T["q_W2"]), _ptr(T["q_b2"]), _ptr(self.logpi), _ptr(self.rew),
                            _ptr(self.term), _ptr(self.alpha), self.reward_scale, self.discount, self.target_entropy, B, _ptr(self.q), _ptr(self.qt), _ptr(self.dq),
                            _ptr(self.dH2q), _ptr(self.y), _ptr(self.sums), _ptr(G["log_alpha"]), st))
        # twin-Q backward: weight grads from the Bellman rows [B, 2B) only, input grads for the policy rows [0, B).  The input-gradient
        # chain (dH2q -> dH1q -> gX -> policy head -> dH2p -> dH1p) runs on the main stream; every weight/bias gradient only needs the
        # activation gradient of its own layer and goes to the side streams as soon as that exists.  `mask=` is the ReLU backward of the
        # layer that produced the mask tensor, fused into the product's epilogue.
        def weight_grads(product, bias_sum):      # the weight-gradient product and the bias column sum of one layer: off the chain, on side streams
            sW = sWs[self._wi % len(sWs)]             # (the six products are independent of each other: round-robin over three streams, or the
            self._wi += 1                             #  last ones queue up behind the first and end the update late)
            sW.wait_stream(main); sB.wait_stream(main)
            with t.cuda.stream(sW):
                product()
            with t.cuda.stream(sB):
                bias_sum()
        weight_grads(lambda: mm(self.H2q[:, B:].transpose(1, 2), self.dq[:, B:], G["q_W2"]), lambda: self._colsum(self.dq, B, 2 * B, G["q_b2"], 2))
        weight_grads(lambda: mm(self.H1q[:, B:].transpose(1, 2), self.dH2q[:, B:], G["q_W1"]), lambda: self._colsum(self.dH2q, B, 2 * B, G["q_b1"], 2))
        mm(self.dH2q, P["q_W1"].transpose(1, 2), self.dH1q, mask=self.H1q)
        weight_grads(lambda: mm(XQ2[:, B:].transpose(1, 2), self.dH1q[:, B:], G["q_W0"]), lambda: self._colsum(self.dH1q, B, 2 * B, G["q_b0"], 2))
        if self.gemm == "tcgen05":       # dX = dH_1 W_1^T + dH_2 W_2^T as ONE product: the two networks' blocks side by side along the contraction
            gemm_tf32(self.dH1q[:, :B], P["q_W0"].transpose(1, 2), self.gX, stack_k=True)
        else:
            mm(self.dH1q[0, :B], P["q_W0"][0].t(), self.gX); mm(self.dH1q[1, :B], P["q_W0"][1].t(), self.gX, accumulate=True)
        # policy backward: head backward + the last layer's input gradient in one kernel
        _chk(L.rsb_policy_head_bwd(_ptr(self.OUT), _ptr(self.eps), _ptr(self.a_store), _ptr(self.H2p), _ptr(P["p_W2"]), B, A, _ptr(self.alpha),
                                   C.c_void_p(self.gX.data_ptr() + 4 * O), self.ldQ, _ptr(self.dOUT), _ptr(self.dH2p), st))
        weight_grads(lambda: mm(self.H2p[:B].t(), self.dOUT[:B], G["p_W2"]), lambda: self._colsum(self.dOUT, 0, B, G["p_b2"]))
        weight_grads(lambda: mm(self.H1p[:B].t(), self.dH2p, G["p_W1"]), lambda: self._colsum(self.dH2p, 0, B, G["p_b1"]))
        mm(self.dH2p, P["p_W1"].t(), self.dH1p, mask=self.H1p[:B])
        weight_grads(lambda: mm(self.Xp[:B].t(), self.dH1p, G["p_W0"]), lambda: self._colsum(self.dH1p, 0, B, G["p_b0"]))
        for sW in sWs:
            main.wait_stream(sW)
        main.wait_stream(sB)

    def _apply(self, do_soft, tick=True, count=True):
        """Adam on policy / Q1 / Q2 / log_alpha + Polyak in one kernel.  tick: advance the bias corrections first (unless `_update_body(tick_early=True)`
        did); count: advance the device-resident update counter (the tick kernel does it in the tick_early form)."""
        s = self.store
        _chk(self.L.rsb_adam_polyak(_ptr(s.flat), _ptr(s.grad), _ptr(s.m), _ptr(s.v), s.n, self.policy_lr, self.qf_lr, 0.9, 0.999, 1e-8, _ptr(self.bc),
                                    _ptr(s.target), s.q_begin, s.q_end, self.tau, int(do_soft), _ptr(self.alpha), s.offsets["log_alpha"], int(tick),
                                    _stream(self.device)))
        if count:
            _chk(self.L.rsb_counter_add(C.c_void_p(self.ctr.data_ptr() + 8), 1, _stream(self.device)))

    def _apply_allreduce(self, do_soft):
        """Gradient mean over the ranks + Adam + Polyak in ONE kernel over NVLink peer memory (csrc/rsb_dp.cu); the caller ticked (tick_early)."""
        s, d = self.store, self._dp
        _chk(self.L.rsb_adam_polyak_allreduce(d["grads_dev"], d["flags_dev"], _ptr(d["local"]), self.rank, self.world, _ptr(s.flat), _ptr(s.m), _ptr(s.v), s.n,
                                              self.policy_lr, self.qf_lr, 0.9, 0.999, 1e-8, _ptr(self.bc), _ptr(s.target), s.q_begin, s.q_end, self.tau, int(do_soft),
                                              _ptr(self.alpha), s.offsets["log_alpha"], _stream(self.device)))

    def _allreduce(self):
        from .parallel import allreduce_mean_
        allreduce_mean_(self.store.grad, self.world)                # ONE flat bucket per update

    # -- public
    def train_step(self, batch=None, eps=None):
        """One SAC update.  batch=None samples from the replay ring with the Philox rule; `eps` overrides the policy noise.
        Without overrides the whole update -- sampling, noise, forward, backward, Adam, Polyak -- is ONE CUDA-graph replay (world size 1), or
        graph / gradient all-reduce / graph (data parallel)."""
        t = self.torch
        step = self._n_train_steps_total
        do_soft = (step % self.period) == 0
        if self.B is None and batch is None:
            raise RsbError("SACTrainer.train_step(): no batch size set (pass batch_size= or call set_batch_size)")
        old = t.backends.cuda.matmul.allow_tf32
        t.backends.cuda.matmul.allow_tf32 = self.tf32
        try:
            sample = batch is None
            if batch is not None:
                self.load_batch(batch)
            else:
                self._sync_ring_size()
                if self._ring_size_seen <= 0:
                    raise RsbError("replay_sample: the replay ring is empty")
            if eps is not None:
                self.eps.copy_(t.as_tensor(eps, dtype=t.float32, device=self.device).reshape(self.eps.shape))
            if self.use_graph and eps is None:
                if self.world == 1 or self._dp is not None:
                    # no host-side collective between the gradients and the optimizer: body + (all-reduce +) Adam/Polyak are ONE graph per Polyak flavour
                    key = ("update", do_soft, sample)
                    if key not in self._graphs:
                        self._warm(lambda: self._update_body(sample, True, False))       # (in data-parallel mode the body only READS peer flags: no rank waits here)
                        g = t.cuda.CUDAGraph()
                        with t.cuda.graph(g):
                            self._update_body(sample, True, True)
                            if self._dp is None:
                                self._apply(do_soft, tick=False, count=False)
                            else:
                                self._apply_allreduce(do_soft)
                        self._graphs[key] = g
                    self._graphs[key].replay()
                else:
                    key = ("body", sample)
                    if key not in self._graphs:
                        self._warm(lambda: self._update_body(sample, True, False))
                        g = t.cuda.CUDAGraph()
                        with t.cuda.graph(g):
                            self._update_body(sample, True, False)
                        self._graphs[key] = g
                    if "apply" not in self._graphs:
                        ga, gb = t.cuda.CUDAGraph(), t.cuda.CUDAGraph()
                        with t.cuda.graph(ga):
                            self._apply(True)
                        with t.cuda.graph(gb):
                            self._apply(False)
                        self._graphs["apply"] = (ga, gb)
                    self._graphs[key].replay()
                    self._allreduce()
                    ga, gb = self._graphs["apply"]
                    (ga if do_soft else gb).replay()
            elif self._dp is not None:
                self._update_body(sample, eps is None, True)
                self._apply_allreduce(do_soft)
            else:
                self._update_body(sample, eps is None, False)
                self._allreduce()
                self._apply(do_soft)
        finally:
            t.backends.cuda.matmul.allow_tf32 = old
        if self._need_to_update_eval_statistics:
            self._need_to_update_eval_statistics = False
            self._fill_statistics()
        self._n_train_steps_total += 1

    def _warm(self, fn):
        """Graph capture must not be the first time cuBLAS sees these shapes (workspace allocation): run the body once eagerly
        on a side stream (it only writes scratch buffers and the gradient buffer, never the parameters)."""
        t = self.torch
        side = t.cuda.Stream(self.device)
        side.wait_stream(t.cuda.current_stream(self.device))
        with t.cuda.stream(side):
            fn()
        t.cuda.current_stream(self.device).wait_stream(side)
        t.cuda.synchronize(self.device)

    def train(self, np_batch):
        """rlkit TorchTrainer.train(np_batch) (util/rlkit_custom.py:238)."""
        self.train_step(batch=np_batch)

    train_from_torch = train

    def _fill_statistics(self):
        t, B, A = self.torch, self.B, self.store.A
        s = self.sums.cpu().numpy()
        st = OrderedDict()
        st["QF1 Loss"], st["QF2 Loss"], st["Policy Loss"] = float(s[0]), float(s[1]), float(s[2])

        def add(name, x):
            x = x.detach().float().reshape(-1)
            st[name + " Mean"], st[name + " Std"] = float(x.mean()), float(x.std(unbiased=False))
            st[name + " Max"], st[name + " Min"] = float(x.max()), float(x.min())

        add("Q1 Predictions", self.q[0, B:]); add("Q2 Predictions", self.q[1, B:]); add("Q Targets", self.y)
        add("Log Pis", self.logpi[:B]); add("Policy mu", self.OUT[:B, :A]); add("Policy log std", self.OUT[:B, A:].clamp(-20, 2))
        st["Alpha"], st["Alpha Loss"] = float(self.alpha[0].item()), float(s[3])
        self.eval_statistics = st

    def get_diagnostics(self):
        return self.eval_statistics

    def end_epoch(self, epoch):
        self._need_to_update_eval_statistics = True

    @property
    def networks(self):
        return [self.policy, self.qf1, self.qf2, self.target_qf1, self.target_qf2]

    def get_snapshot(self):
        return dict(policy=self.policy, qf1=self.qf1, qf2=self.qf2, target_qf1=self.target_qf1, target_qf2=self.target_qf2)


def algorithmic_flops_per_update(obs_dim, act_dim, batch, hidden=HID):
    """2*B*[6*MAC_Q + 2*MAC_pi (fwd) + 2*(2*MAC_Q + MAC_pi) + 4*MAC_Q (bwd)]."""
    mq = (obs_dim + act_dim) * hidden + hidden * hidden + hidden
    mp = obs_dim * hidden + hidden * hidden + 2 * act_dim * hidden
    return 2 * batch * (6 * mq + 2 * mp + 2 * (2 * mq + mp) + 4 * mq)
