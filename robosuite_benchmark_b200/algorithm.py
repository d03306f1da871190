"""The outer hot loop of the reference -- `CustomBatchRLAlgorithm._train` (util/rlkit_custom.py:199-242), its collectors
(rlkit MdpPathCollector + rollout), path statistics (`get_custom_generic_path_information`, util/rlkit_custom.py:315-377), the
epoch log (`_log_stats`, :244-301: the 83 `progress.csv` columns) and the snapshot (`_get_snapshot`, :68-82) -- on the batched
CUDA backend.  `experiment(variant)` mirrors util/rlkit_utils.py:31-165.

Two collectors share one interface:
  * `MdpPathCollector`  : rlkit's single-env protocol (numpy paths through GymWrapper/RobosuiteEnv) -- the reference's own shape;
  * `BatchedPathCollector`: N envs stepped together on the GPU; a "path" is one env's episode segment, all paths of a round are
    kept as stacked device tensors and appended to the replay ring in bulk.
"""
from __future__ import annotations

import csv
import json
import os
import pickle
import time
from collections import OrderedDict, deque

import numpy as np


# ----------------------------------------------------------------------------- statistics (rlkit eval_util)
def create_stats_ordered_dict(name, data, stat_prefix=None, always_show_all_stats=True):
    if stat_prefix is not None:
        name = "{}{}".format(stat_prefix, name)
    if isinstance(data, (int, float)):
        return OrderedDict({name: data})
    data = np.asarray(data) if not isinstance(data, (list, tuple)) else np.concatenate([np.asarray(d).reshape(-1) for d in data]) if len(data) else np.zeros(0)
    if data.size == 0:
        return OrderedDict()
    return OrderedDict([(name + " Mean", float(np.mean(data))), (name + " Std", float(np.std(data))),
                        (name + " Max", float(np.max(data))), (name + " Min", float(np.min(data)))])


def get_average_returns(paths):
    return float(np.mean([np.sum(p["rewards"]) for p in paths]))


def get_generic_path_information(paths, stat_prefix=""):
    """rlkit.core.eval_util.get_generic_path_information (exploration statistics)."""
    st = OrderedDict()
    returns = [float(np.sum(p["rewards"])) for p in paths]
    rewards = np.vstack([np.asarray(p["rewards"]).reshape(-1, 1) for p in paths])
    st.update(create_stats_ordered_dict("Rewards", rewards, stat_prefix=stat_prefix))
    st.update(create_stats_ordered_dict("Returns", np.asarray(returns), stat_prefix=stat_prefix))
    actions = np.vstack([np.asarray(p["actions"]).reshape(len(p["actions"]), -1) for p in paths])
    st.update(create_stats_ordered_dict("Actions", actions, stat_prefix=stat_prefix))
    st["Num Paths"] = len(paths)
    st[stat_prefix + "Average Returns"] = get_average_returns(paths)
    return st


def get_custom_generic_path_information(paths, path_length, reward_scale, stat_prefix=""):
    """util/rlkit_custom.py:315-377: adds ExplReturns = return accumulated up to `path_length` steps."""
    st = OrderedDict()
    returns = [float(np.sum(p["rewards"])) for p in paths]
    expl_returns = [float(np.sum(np.asarray(p["rewards"])[:path_length])) for p in paths]
    rewards = np.vstack([np.asarray(p["rewards"]).reshape(-1, 1) for p in paths])
    st.update(create_stats_ordered_dict("Rewards", rewards, stat_prefix=stat_prefix))
    st.update(create_stats_ordered_dict("Returns", np.asarray(returns), stat_prefix=stat_prefix))
    st.update(create_stats_ordered_dict("ExplReturns", np.asarray(expl_returns), stat_prefix=stat_prefix))
    actions = np.vstack([np.asarray(p["actions"]).reshape(len(p["actions"]), -1) for p in paths])
    st.update(create_stats_ordered_dict("Actions", actions, stat_prefix=stat_prefix))
    st["Num Paths"] = len(paths)
    st[stat_prefix + "Average Returns"] = get_average_returns(paths)
    return st


# ----------------------------------------------------------------------------- single-env collectors (rlkit protocol)
class NormalizedBoxEnv:
    """rlkit.envs.wrappers.NormalizedBoxEnv: actions in [-1, 1] mapped affinely to [lb, ub] and clipped; reward_scale 1."""

    def __init__(self, env, reward_scale=1.0):
        self._wrapped_env, self._reward_scale = env, reward_scale
        self.action_space, self.observation_space = env.action_space, env.observation_space

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return getattr(self._wrapped_env, k)

    def reset(self, *a, **k):
        return self._wrapped_env.reset(*a, **k)

    def step(self, action):
        lb, ub = self._wrapped_env.action_space.low, self._wrapped_env.action_space.high
        if hasattr(action, "clamp"):                               # batched torch actions
            import torch
            # in float64 like rlkit's numpy arithmetic: for robosuite's (-1, 1) bounds the map is then EXACTLY the identity on fp32 actions
            # (in fp32, (a + 1) - 1 != a), which is what the fused collector assumes when it bypasses this wrapper
            lo, hi = torch.as_tensor(lb, device=action.device, dtype=torch.float64), torch.as_tensor(ub, device=action.device, dtype=torch.float64)
            scaled = torch.minimum(torch.maximum(lo + (action.double() + 1.0) * 0.5 * (hi - lo), lo), hi).float().contiguous()
        else:
            scaled = np.clip(lb + (np.asarray(action) + 1.0) * 0.5 * (ub - lb), lb, ub)
        o, r, d, info = self._wrapped_env.step(scaled)
        return o, r * self._reward_scale, d, info


def rollout(env, agent, max_path_length=np.inf, render=False, render_kwargs=None):
    """rlkit.samplers.rollout_functions.rollout: the episode loop the collectors run (and the path-dict layout that
    util/rlkit_custom.py:380-495 also returns)."""
    observations, actions, rewards, terminals, agent_infos, env_infos = [], [], [], [], [], []
    o = env.reset()
    agent.reset()
    next_o, path_length = None, 0
    while path_length < max_path_length:
        a, agent_info = agent.get_action(o)
        next_o, r, d, env_info = env.step(a)
        observations.append(o); rewards.append(r); terminals.append(d); actions.append(a)
        agent_infos.append(agent_info); env_infos.append(env_info)
        path_length += 1
        if d:
            break
        o = next_o
    actions = np.array(actions)
    if len(actions.shape) == 1:
        actions = np.expand_dims(actions, 1)
    observations = np.array(observations)
    if len(observations.shape) == 1:
        observations = np.expand_dims(observations, 1)
        next_o = np.array([next_o])
    next_observations = np.vstack((observations[1:, :], np.expand_dims(next_o, 0)))
    return dict(observations=observations, actions=actions, rewards=np.array(rewards).reshape(-1, 1), next_observations=next_observations,
                terminals=np.array(terminals).reshape(-1, 1), agent_infos=agent_infos, env_infos=env_infos)


class MdpPathCollector:
    def __init__(self, env, policy, max_num_epoch_paths_saved=None):
        self._env, self._policy = env, policy
        self._epoch_paths = deque(maxlen=max_num_epoch_paths_saved)
        self._num_steps_total = self._num_paths_total = 0

    def collect_new_paths(self, max_path_length, num_steps, discard_incomplete_paths):
        paths, collected = [], 0
        while collected < num_steps:
            length_this = min(max_path_length, num_steps - collected)
            path = rollout(self._env, self._policy, max_path_length=length_this)
            plen = len(path["actions"])
            if plen != max_path_length and not path["terminals"][-1] and discard_incomplete_paths:
                break
            collected += plen
            paths.append(path)
        self._num_paths_total += len(paths)
        self._num_steps_total += collected
        self._epoch_paths.extend(paths)
        return paths

    def get_epoch_paths(self):
        return self._epoch_paths

    def end_epoch(self, epoch):
        self._epoch_paths = deque(maxlen=self._epoch_paths.maxlen)

    def get_diagnostics(self):
        st = OrderedDict([("num steps total", self._num_steps_total), ("num paths total", self._num_paths_total)])
        st.update(create_stats_ordered_dict("path length", np.asarray([len(p["actions"]) for p in self._epoch_paths])))
        return st

    def get_snapshot(self):
        return dict(env=self._env, policy=self._policy)


# ----------------------------------------------------------------------------- batched collectors
def _stats_dict(name, v4, count, stat_prefix=""):
    """Mean / Std / Max / Min (numpy's population std) from {sum, sum of squares, max, min} and the sample count."""
    mean = v4[0] / count
    var = max(v4[1] / count - mean * mean, 0.0)
    n = stat_prefix + name
    return OrderedDict([(n + " Mean", float(mean)), (n + " Std", float(np.sqrt(var))), (n + " Max", float(v4[2])), (n + " Min", float(v4[3]))])


class BatchedPaths:
    """One collection round of N environments x T control steps.  The transitions live in a replay ring (rlkit EnvReplayBuffer layout) at
    rows (slot0 + t N + i) mod capacity -- written there by the env kernels themselves in the fused collector -- and the round's
    statistics are reduced on the device (`stats`: 16 doubles, csrc/rsb_collect.cu k_path_stats).  `in_place`: the ring is the replay
    buffer the algorithm trains from, so `add_paths` only advances its pointer.  Iterating yields numpy path dicts (diagnostics only)."""

    def __init__(self, ring, slot0, T, N, stats, in_place, expl_len):
        self.ring, self.slot0, self.T, self.N, self.stats, self.in_place, self.expl_len = ring, int(slot0), int(T), int(N), stats, bool(in_place), int(expl_len)
        self.committed = False
        self._host_stats = None

    def host_stats(self):
        if self._host_stats is None:
            self._host_stats = self.stats[:16].cpu().numpy().astype(np.float64)
        return self._host_stats

    def _rows(self):
        import torch
        r = torch.arange(self.T * self.N, device=self.ring.device)
        return (r + self.slot0) % self.ring.capacity

    def tensors(self):
        """(obs, act, rew, term, next_obs) stacked [T, N, .] device tensors (a gather: used by tests and the copy path, not by training)."""
        rows, T, N, R = self._rows(), self.T, self.N, self.ring
        return (R._observations[rows].view(T, N, -1), R._actions[rows].view(T, N, -1), R._rewards[rows].view(T, N),
                R._terminals[rows].view(T, N), R._next_obs[rows].view(T, N, -1))

    def __len__(self):
        return self.N

    def __iter__(self):
        _, act, rew, _, _ = self.tensors()
        rew, act = rew.cpu().numpy(), act.cpu().numpy()
        for n in range(self.N):
            yield dict(rewards=rew[:, n].reshape(-1, 1), actions=act[:, n])


def batched_path_information(rounds, stat_prefix="", custom=False, reduce_fn=None):
    """get_generic_path_information / get_custom_generic_path_information (util/rlkit_custom.py:315-377) for rounds of BatchedPaths:
    the per-round device reductions are combined on the host (16 doubles per round) -- and over the ranks by `reduce_fn` in
    data-parallel runs -- into the same statistic names."""
    acc = np.zeros(16)
    acc[2::4], acc[3::4] = -np.inf, np.inf
    n_rew = n_ret = n_act = 0
    for r in rounds:
        h = r.host_stats()
        acc[0::4] += h[0::4]; acc[1::4] += h[1::4]
        acc[2::4] = np.maximum(acc[2::4], h[2::4]); acc[3::4] = np.minimum(acc[3::4], h[3::4])
        n_rew += r.T * r.N; n_ret += r.N; n_act += r.T * r.N * r.ring.action_dim
    counts = np.array([n_rew, n_ret, n_act], np.float64)
    if reduce_fn is not None:
        acc, counts = reduce_fn(acc, counts)
    st = OrderedDict()
    if counts[0] == 0:
        return st
    st.update(_stats_dict("Rewards", acc[0:4], counts[0], stat_prefix))
    st.update(_stats_dict("Returns", acc[4:8], counts[1], stat_prefix))
    if custom:
        st.update(_stats_dict("ExplReturns", acc[8:12], counts[1], stat_prefix))
    st.update(_stats_dict("Actions", acc[12:16], counts[2], stat_prefix))
    st["Num Paths"] = int(counts[1])
    st[stat_prefix + "Average Returns"] = float(acc[4] / counts[1])
    return st


def _unwrap_batched(env):
    """(BatchSim, GymWrapper) under NormalizedBoxEnv(GymWrapper(BatchedEnv)); checks that the wrappers are pass-through on the device path."""
    gw = getattr(env, "_wrapped_env", env)
    lo, hi = gw.action_space.low, gw.action_space.high
    if not (np.all(lo == -1.0) and np.all(hi == 1.0)):
        raise NotImplementedError("the fused collector needs robosuite's (-1, 1) action spec (NormalizedBoxEnv is then the identity)")
    if not getattr(gw, "_order_is_native", True):
        raise NotImplementedError("the fused collector writes observations in the kernels' native key order; GymWrapper(keys=...) with another "
                                  "order needs CopyingPathCollector")
    return gw.env.sim, gw.env


class BatchedPathCollector:
    """The fused collector: N envs stepped together with ZERO torch ops and zero copies per control step.

        rsb_reset_ring(top)                                   first observations -> ring.observations[top + i]
        for t < T:  k_policy_act(ring rows top + t N ..)      actions -> ring.actions[row]            (csrc/rsb_collect.cu)
                    k_step(ring rows top + t N ..)            next_obs, reward, terminal -> ring[row]; obs -> ring.observations[row + N]

    collect_new_paths(max_path_length, num_steps, discard) keeps rlkit's MdpPathCollector contract (util/rlkit_custom.py:202-227): rounds of
    N fresh episodes (every rollout starts with env.reset()) until num_steps transitions are gathered; a round's length is
    min(max_path_length, ceil(remaining / N)); incomplete rounds are dropped when discard_incomplete_paths is set.  With `replay=` the
    ring is the training replay buffer itself (exploration); otherwise the collector owns a scratch ring (evaluation)."""

    def __init__(self, env, policy, replay=None, deterministic=False, expl_len=None):
        self._env, self._policy, self._det = env, policy, bool(deterministic)
        self._sim, self._benv = _unwrap_batched(env)
        self._replay, self._own = replay, None
        self._epoch_paths, self._uncommitted = [], []
        self._num_steps_total = self._num_paths_total = 0
        self._expl_len = expl_len
        self._policy_step = 0                       # keys the exploration noise: one value per control step, monotonic over the run

    def _ring_for(self, rows):
        if self._replay is not None:
            return self._replay, True
        from .sac import EnvReplayBuffer
        if self._own is None or self._own.capacity < rows:
            self._own = EnvReplayBuffer(rows, obs_dim=self._sim.obs_dim, action_dim=self._sim.act_dim, device=self._sim.device)
        return self._own, False

    def collect_new_paths(self, max_path_length, num_steps, discard_incomplete_paths):
        import torch
        from .sac import _chk, _ptr, _stream
        from .backend import lib
        sim, N = self._sim, self._sim.num_envs
        O, A = sim.obs_dim, sim.act_dim
        horizon = self._benv.horizon if not self._benv.ignore_done else None
        self._uncommitted = [r for r in self._uncommitted if not r.committed]
        rounds, collected = [], 0
        plan, left = [], num_steps                   # round lengths first: an evaluation scratch ring is sized for the whole call
        while left > 0:
            T = int(min(max_path_length, -(-left // N)))
            if T != max_path_length and discard_incomplete_paths:
                break
            if horizon is not None:
                T = min(T, horizon)                 # an episode ends at the horizon (done = 1): never step a terminated episode
            plan.append(T)
            left -= T * N
        ring, in_place = self._ring_for(sum(plan) * N if plan else N)
        if not in_place:
            ring._top, ring._size = 0, 0
        cap = ring.capacity
        pending = sum(r.T * r.N for r in self._uncommitted) if in_place else 0
        for T in plan:
            if (pending + T * N) > cap:
                raise ValueError(f"a collection round of {T} x {N} transitions does not fit the ring ({cap} rows)")
            top = (ring.top + pending) % cap
            sim.reset_ring(ring.ring, top)
            for t in range(T):
                slot = (top + t * N) % cap
                self._policy.act_into(_ptr(ring._observations), O, _ptr(ring._actions), A, N, slot0=slot, cap=cap,
                                      deterministic=self._det, step=self._policy_step)
                self._policy_step += 1
                sim.step_ring(ring.ring, slot, t + 1 < T)
            stats = torch.empty(lib().rsb_path_stats_words(N), dtype=torch.float64, device=sim.device)
            _chk(lib().rsb_path_stats(_ptr(ring._rewards), _ptr(ring._actions), top, cap, N, T, A,
                                      int(self._expl_len if self._expl_len is not None else T), _ptr(stats), _stream(sim.device)))
            r = BatchedPaths(ring, top, T, N, stats, in_place, self._expl_len if self._expl_len is not None else T)
            rounds.append(r)
            if in_place:
                self._uncommitted.append(r); pending += T * N
            else:
                ring.commit(T * N); r.committed = True
            collected += T * N
        if self._sim.info("steps_after_done") > 0:
            raise ValueError("executing action in terminated episode")
        self._num_paths_total += sum(r.N for r in rounds)
        self._num_steps_total += collected
        self._epoch_paths.extend(rounds)
        return rounds

    def get_epoch_paths(self):
        return list(self._epoch_paths)

    def end_epoch(self, epoch):
        self._epoch_paths = []

    def get_diagnostics(self):
        st = OrderedDict([("num steps total", self._num_steps_total), ("num paths total", self._num_paths_total)])
        lens = np.concatenate([np.full(r.N, r.T) for r in self._epoch_paths]) if self._epoch_paths else np.zeros(0)
        st.update(create_stats_ordered_dict("path length", lens))
        return st

    def get_snapshot(self):
        return dict(policy=self._policy)


class CopyingPathCollector(BatchedPathCollector):
    """The round-1 data flow, kept as the comparison arm of the fused collector (tests/test_gpu_collector.py holds the two to bit-equal ring
    contents) and for GymWrapper(keys=...) orders the kernels do not write natively: observations go through a scratch row, torch
    `copy_` stacks them per step, `add_paths` copies the stack into the ring.  Same kernels, same noise keys."""

    def __init__(self, env, policy, replay=None, deterministic=False, expl_len=None):
        self._env, self._policy, self._det = env, policy, bool(deterministic)
        gw = getattr(env, "_wrapped_env", env)
        self._sim, self._benv = gw.env.sim, gw.env
        self._replay, self._own = None, None
        self._epoch_paths, self._uncommitted = [], []
        self._num_steps_total = self._num_paths_total = 0
        self._expl_len, self._policy_step = expl_len, 0

    def collect_new_paths(self, max_path_length, num_steps, discard_incomplete_paths):
        import torch
        from .sac import EnvReplayBuffer, _chk, _ptr, _stream
        from .backend import lib
        env, N = self._env, self._sim.num_envs
        rounds, collected = [], 0
        horizon = self._benv.horizon if not self._benv.ignore_done else None
        while collected < num_steps:
            T = int(min(max_path_length, -(-(num_steps - collected) // N)))
            if T != max_path_length and discard_incomplete_paths:
                break
            if horizon is not None:
                T = min(T, horizon)
            o = env.reset()
            O, A = o.shape[1], env.action_space.low.size
            obs = torch.empty(T + 1, N, O, device=o.device)
            act, rew = torch.empty(T, N, A, device=o.device), torch.empty(T, N, device=o.device)
            term = torch.empty(T, N, dtype=torch.uint8, device=o.device)
            obs[0].copy_(o)
            for t in range(T):
                self._policy.get_actions(obs[t], deterministic=self._det, out=act[t], step=self._policy_step)
                self._policy_step += 1
                o2, r, d, _ = env.step(act[t])
                if bool((d == 2).any()):
                    raise ValueError("executing action in terminated episode")
                obs[t + 1].copy_(o2); rew[t].copy_(r); term[t].copy_(d)
            ring = EnvReplayBuffer(T * N, obs_dim=O, action_dim=A, device=o.device)      # the round as a private ring (same statistics kernel)
            ring.add_batch(obs[:-1].reshape(T * N, O), act.reshape(T * N, A), rew.reshape(-1), term.reshape(-1), obs[1:].reshape(T * N, O))
            stats = torch.empty(lib().rsb_path_stats_words(N), dtype=torch.float64, device=o.device)
            el = int(self._expl_len if self._expl_len is not None else T)
            _chk(lib().rsb_path_stats(_ptr(ring._rewards), _ptr(ring._actions), 0, ring.capacity, N, T, A, el, _ptr(stats), _stream(o.device)))
            rounds.append(BatchedPaths(ring, 0, T, N, stats, False, el))
            collected += T * N
        self._num_paths_total += sum(r.N for r in rounds)
        self._num_steps_total += collected
        self._epoch_paths.extend(rounds)
        return rounds


def add_paths(replay, paths):
    """replay_buffer.add_paths (util/rlkit_custom.py:207,230) for every collector: rounds the env kernels wrote in place only advance the
    ring pointer; other rounds are copied row-block by row-block; rlkit path dicts go through EnvReplayBuffer.add_paths."""
    if paths and isinstance(paths[0], BatchedPaths):
        for r in paths:
            if r.in_place and r.ring is replay:
                assert r.slot0 == replay.top, "in-place rounds must be committed in collection order"
                replay.commit(r.T * r.N)
                r.committed = True
            else:
                T, N = r.T, r.N
                obs, act, rew, term, nxt = r.tensors()
                replay.add_batch(obs.reshape(T * N, -1), act.reshape(T * N, -1), rew.reshape(-1), term.reshape(-1), nxt.reshape(T * N, -1))
    else:
        replay.add_paths(paths)


# ----------------------------------------------------------------------------- logger (progress.csv / variant.json / params.pkl)
class Logger:
    def __init__(self, log_dir, variant=None):
        self.dir = log_dir
        os.makedirs(log_dir, exist_ok=True)
        self._row, self._header, self._f = OrderedDict(), None, None
        if variant is not None:
            with open(os.path.join(log_dir, "variant.json"), "w") as f:
                json.dump(variant, f, indent=2, sort_keys=True)
        self._dbg = open(os.path.join(log_dir, "debug.log"), "a")

    def log(self, msg):
        self._dbg.write(f"{time.strftime('%Y-%m-%d %H:%M:%S')} | {msg}\n"); self._dbg.flush()

    def record_dict(self, d, prefix=""):
        for k, v in d.items():
            self._row[prefix + k] = v

    def record_tabular(self, k, v):
        self._row[k] = v

    def dump_tabular(self):
        if self._header is None:
            self._header = list(self._row.keys())
            self._f = open(os.path.join(self.dir, "progress.csv"), "w", newline="")
            self._w = csv.DictWriter(self._f, fieldnames=self._header)
            self._w.writeheader()
        self._w.writerow({k: self._row.get(k, "") for k in self._header}); self._f.flush()
        self._row = OrderedDict()

    def save_itr_params(self, epoch, snapshot):
        """rlkit logger.save_itr_params (snapshot_mode "last"): params.pkl is a torch-pickled dict of network objects, so the reference's
        consumers read it with `torch.load(path)['evaluation/policy'].get_action(obs)` (util/rlkit_utils.py:173-174,241-242)."""
        import torch
        tmp = os.path.join(self.dir, "params.pkl.tmp")
        torch.save(snapshot, tmp)
        os.replace(tmp, os.path.join(self.dir, "params.pkl"))


# ----------------------------------------------------------------------------- the algorithm
class BatchRLAlgorithm:
    """CustomTorchBatchRLAlgorithm (util/rlkit_custom.py:163-312): same constructor arguments, same `_train` order.

    Data parallel (`world_size` > 1, one process per GPU): every rank runs this same loop on its own env slice and replay
    shard; the ranks meet in the SAC update's gradient all-reduce (inside trainer.train_step) and once per epoch in the reduction of the
    logged statistics; rank 0 writes progress.csv / params.pkl."""

    def __init__(self, trainer, exploration_env, evaluation_env, exploration_data_collector, evaluation_data_collector, replay_buffer,
                 batch_size, expl_max_path_length, eval_max_path_length, num_epochs, num_eval_steps_per_epoch,
                 num_expl_steps_per_train_loop, num_trains_per_train_loop, num_train_loops_per_epoch=1,
                 min_num_steps_before_training=0, logger=None, rank=0, world_size=1):
        self.trainer, self.expl_env, self.eval_env = trainer, exploration_env, evaluation_env
        self.expl_data_collector, self.eval_data_collector, self.replay_buffer = exploration_data_collector, evaluation_data_collector, replay_buffer
        self.batch_size, self.expl_max_path_length, self.eval_max_path_length = batch_size, expl_max_path_length, eval_max_path_length
        self.num_epochs, self.num_eval_steps_per_epoch = num_epochs, num_eval_steps_per_epoch
        self.num_expl_steps_per_train_loop, self.num_trains_per_train_loop = num_expl_steps_per_train_loop, num_trains_per_train_loop
        self.num_train_loops_per_epoch, self.min_num_steps_before_training = num_train_loops_per_epoch, min_num_steps_before_training
        self._start_epoch, self.logger, self._t0 = 0, logger, time.time()
        self.rank, self.world = int(rank), int(world_size)
        self.trainer.replay = replay_buffer
        self.trainer.set_batch_size(batch_size)
        self.epoch_times = []                        # per epoch: the time/* columns (also kept without a logger: bench.py --mode train reads them)
        self.post_epoch_funcs = []                   # util/rlkit_custom.py:42,65-66: called as f(algorithm, epoch) at the end of every epoch

    def to(self, device):
        return self

    def training_mode(self, mode):
        pass

    def train(self, start_epoch=0):
        self._start_epoch = start_epoch
        self._train()

    def _sync(self):
        import torch
        torch.cuda.synchronize()

    def _train(self):
        times = OrderedDict()

        def stamp(name, t0):
            self._sync()
            times[name] = times.get(name, 0.0) + time.time() - t0
            return time.time()

        if self.min_num_steps_before_training > 0:
            init = self.expl_data_collector.collect_new_paths(self.expl_max_path_length, self.min_num_steps_before_training, discard_incomplete_paths=False)
            add_paths(self.replay_buffer, init)
            self.expl_data_collector.end_epoch(-1)
        for epoch in range(self._start_epoch, self.num_epochs):
            times.clear()
            e0 = t = time.time()
            self.eval_data_collector.collect_new_paths(self.eval_max_path_length, self.num_eval_steps_per_epoch, discard_incomplete_paths=True)
            t = stamp("evaluation sampling", t)
            for _ in range(self.num_train_loops_per_epoch):
                new = self.expl_data_collector.collect_new_paths(self.expl_max_path_length, self.num_expl_steps_per_train_loop, discard_incomplete_paths=False)
                t = stamp("exploration sampling", t)
                add_paths(self.replay_buffer, new)
                t = stamp("data storing", t)
                for _ in range(self.num_trains_per_train_loop):
                    self.trainer.train_step()                      # samples its batch from the ring with the Philox rule
                t = stamp("training", t)
            self._end_epoch(epoch, times, e0)

    def _get_snapshot(self):
        """util/rlkit_custom.py:68-82: trainer/, exploration/, evaluation/ entries (`env` entries skipped, as there).  The values are the
        network objects themselves; they pickle as their weights (sac._Net.__getstate__) and come back as objects with get_action()."""
        snap = {}
        for k, v in self.trainer.get_snapshot().items():
            snap["trainer/" + k] = v
        for k, v in self.expl_data_collector.get_snapshot().items():
            if k != "env":
                snap["exploration/" + k] = v
        for k, v in self.eval_data_collector.get_snapshot().items():
            if k != "env":
                snap["evaluation/" + k] = v
        for k, v in self.replay_buffer.get_snapshot().items():
            snap["replay_buffer/" + k] = v
        return snap

    def _check_device_health(self):
        """Once per epoch (synchronises): a tensor-core GEMM whose bounded wait gave up has poisoned its tile with NaN -- stop here with the
        cause instead of training on; contact / constraint-row truncation of the env batches is reported, never silent."""
        from . import gemm
        n = gemm.timeouts()
        if n:
            raise RuntimeError(f"{n} mbarrier waits of the tcgen05 GEMM timed out during this epoch (results poisoned with NaN)")
        n = self.trainer.dp_timeouts() if hasattr(self.trainer, "dp_timeouts") else 0
        if n:
            raise RuntimeError(f"{n} cross-rank waits of the fused gradient all-reduce timed out during this epoch (a rank is missing or stalled)")
        out = {}
        for name, env in (("exploration", self.expl_env), ("evaluation", self.eval_env)):
            sim = getattr(env, "sim", None)
            if sim is not None and hasattr(sim, "counters"):
                out[name] = sim.counters()
        return out

    def _end_epoch(self, epoch, times, e0):
        t = time.time()
        if self.logger is not None:
            self.logger.save_itr_params(epoch, self._get_snapshot())
        times["saving"] = time.time() - t
        health = self._check_device_health()
        if self.logger is not None and any(v["ncon_overflow"] or v["nefc_overflow"] for v in health.values()):
            self.logger.log(f"contact / constraint-row truncation events so far: {health}")
        self._log_stats(epoch, times, e0)
        self.expl_data_collector.end_epoch(epoch)
        self.eval_data_collector.end_epoch(epoch)
        self.replay_buffer.end_epoch(epoch)
        self.trainer.end_epoch(epoch)
        for f in self.post_epoch_funcs:
            f(self, epoch)

    def _path_information(self, collector, custom):
        paths = collector.get_epoch_paths()
        if paths and isinstance(paths[0], BatchedPaths):
            reduce_fn = None
            if self.world > 1:
                from .parallel import reduce_path_stats
                dev = paths[0].ring.device
                reduce_fn = lambda acc, counts: reduce_path_stats(acc, counts, dev)
            return batched_path_information(paths, custom=custom, reduce_fn=reduce_fn)
        if custom:
            return get_custom_generic_path_information(paths, self.expl_max_path_length, self.trainer.reward_scale)
        return get_generic_path_information(paths)

    def _log_stats(self, epoch, times, e0):
        lg, t = self.logger, time.time()
        expl_info = self._path_information(self.expl_data_collector, custom=False)       # all ranks: holds the per-epoch collective
        eval_info = self._path_information(self.eval_data_collector, custom=True)
        times["logging"] = time.time() - t
        row = OrderedDict((f"time/{k} (s)", times.get(k, 0.0)) for k in ("data storing", "evaluation sampling", "exploration sampling", "logging", "saving", "training"))
        row["time/epoch (s)"], row["time/total (s)"] = time.time() - e0, time.time() - self._t0
        self.epoch_times.append(row)
        self.last_eval_info, self.last_expl_info = eval_info, expl_info
        if lg is None:
            return
        lg.log("Epoch {} finished".format(epoch))
        lg.record_dict(self.replay_buffer.get_diagnostics(), prefix="replay_buffer/")
        lg.record_dict(self.trainer.get_diagnostics(), prefix="trainer/")
        lg.record_dict(self.expl_data_collector.get_diagnostics(), prefix="exploration/")
        lg.record_dict(expl_info, prefix="exploration/")
        lg.record_dict(self.eval_data_collector.get_diagnostics(), prefix="evaluation/")
        lg.record_dict(eval_info, prefix="evaluation/")
        for k, v in row.items():
            lg.record_tabular(k, v)
        lg.record_tabular("Epoch", epoch)
        lg.dump_tabular()


# ----------------------------------------------------------------------------- experiment wiring
def make_env(env_config, controller=None, num_envs=1, device="cuda:0", seed=0, env_id_base=0):
    """util/rlkit_utils.py:39-59: pop `controller`, load its config (name in ALL_CONTROLLERS or a JSON path), suite.make, wrap."""
    import robosuite_benchmark_b200 as suite
    from .controllers import ALL_CONTROLLERS, load_controller_config
    from .wrappers import GymWrapper
    cfg = dict(env_config)
    controller = cfg.pop("controller", controller)
    if controller in ALL_CONTROLLERS:
        ccfg = load_controller_config(default_controller=controller)
    else:
        ccfg = load_controller_config(custom_fpath=controller)
    env = suite.make(**cfg, has_renderer=False, has_offscreen_renderer=False, use_object_obs=True, use_camera_obs=False,
                     reward_shaping=True, controller_configs=ccfg, num_envs=num_envs, device=device, seed=seed, env_id_base=env_id_base,
                     batched=num_envs > 1)
    return NormalizedBoxEnv(GymWrapper(env))


def distributed_context():
    """(rank, world, local_rank) of a torchrun launch (RANK / WORLD_SIZE / LOCAL_RANK), (0, 1, 0) otherwise."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def build_experiment(variant, agent="SAC", num_envs=1, device=None, log_dir=None, seed=None, gemm="tcgen05", rank=None, world_size=None, fused=True,
                     allreduce="fused"):
    """util/rlkit_utils.py:31-165 on the batched backend -> the algorithm object, not yet run.

    num_envs = 1 reproduces the reference's single-env data flow (rlkit path dicts through the host entry points).  num_envs = N > 1 steps N
    exploration and N evaluation envs together on the GPU with the fused collector (policy kernel -> step kernel -> replay ring).
    Under torchrun (world_size W > 1) every rank owns N envs of each kind (global env ids rank N .. rank N + N - 1), a replay shard of
    replay_buffer_size / W rows and a batch of batch_size rows; the step counts of algorithm_kwargs are totals over the job and are split
    evenly over the ranks; gradients are averaged over the ranks every update."""
    from .sac import EnvReplayBuffer, MakeDeterministic, ParamStore, SACTrainer, TanhGaussianPolicy, default_device
    if agent != "SAC":
        raise NotImplementedError("only the SAC agent is on the benchmark path (TD3 is named by the reference but out of scope)")
    r0, w0, _ = distributed_context()
    rank, world = (r0 if rank is None else rank), (w0 if world_size is None else world_size)
    device = device or default_device()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            torch.cuda.set_device(torch.device(device))
            dist.init_process_group("nccl", device_id=torch.device(device))
    seed = variant.get("seed", 0) if seed is None else seed
    expl_env = make_env(variant["expl_environment_kwargs"], num_envs=num_envs, device=device, seed=seed, env_id_base=rank * num_envs)
    eval_env = make_env(variant["eval_environment_kwargs"], num_envs=num_envs, device=device, seed=seed, env_id_base=(1 << 20) + rank * num_envs)
    obs_dim, action_dim = expl_env.observation_space.low.size, expl_env.action_space.low.size
    assert list(variant["policy_kwargs"]["hidden_sizes"]) == [256, 256] and list(variant["qf_kwargs"]["hidden_sizes"]) == [256, 256], \
        "the fused update is built for the benchmark's 256x256 networks"
    store = ParamStore(obs_dim, action_dim, device, seed=seed, symmetric=world > 1 and allreduce == "fused")   # same seed on every rank: replicated parameters
    policy = TanhGaussianPolicy.of(store, seed=seed, env_id_base=rank * num_envs)
    eval_policy = TanhGaussianPolicy.of(store, seed=seed, env_id_base=(1 << 20) + rank * num_envs)
    replay = EnvReplayBuffer(max(1, variant["replay_buffer_size"] // world), expl_env, device=device, seed=seed + 7919 * rank)
    ak = dict(variant["algorithm_kwargs"])
    if world > 1:
        for k in ("num_eval_steps_per_epoch", "num_expl_steps_per_train_loop", "min_num_steps_before_training"):
            ak[k] = -(-ak[k] // world)
    trainer = SACTrainer(env=eval_env, store=store, policy=policy, replay_buffer=replay, batch_size=ak["batch_size"], seed=seed, gemm=gemm,
                         world_size=world, rank=rank, allreduce=allreduce, **variant["trainer_kwargs"])
    if num_envs == 1:
        expl_c, eval_c = MdpPathCollector(expl_env, policy), MdpPathCollector(eval_env, MakeDeterministic(eval_policy))
    elif fused:
        expl_c = BatchedPathCollector(expl_env, policy, replay=replay)
        eval_c = BatchedPathCollector(eval_env, MakeDeterministic(eval_policy), deterministic=True, expl_len=ak["expl_max_path_length"])
    else:
        expl_c = CopyingPathCollector(expl_env, policy)
        eval_c = CopyingPathCollector(eval_env, MakeDeterministic(eval_policy), deterministic=True, expl_len=ak["expl_max_path_length"])
    logger = Logger(log_dir, variant) if (log_dir and rank == 0) else None
    return BatchRLAlgorithm(trainer=trainer, exploration_env=expl_env, evaluation_env=eval_env, exploration_data_collector=expl_c,
                            evaluation_data_collector=eval_c, replay_buffer=replay, logger=logger, rank=rank, world_size=world, **ak)


def experiment(variant, agent="SAC", **kw):
    algo = build_experiment(variant, agent, **kw)
    algo.train()
    return algo
