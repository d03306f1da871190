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
            lo, hi = torch.as_tensor(lb, device=action.device), torch.as_tensor(ub, device=action.device)
            scaled = torch.minimum(torch.maximum(lo + (action + 1.0) * 0.5 * (hi - lo), lo), hi).contiguous()
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


# ----------------------------------------------------------------------------- batched collector
class BatchedPaths:
    """All paths of one collection round: stacked device tensors [T, N, .]; path n = column n.  Quacks like a list of rlkit
    path dicts for the statistics code (`__iter__` yields numpy path dicts lazily)."""

    def __init__(self, obs, act, rew, term, next_obs):
        self.obs, self.act, self.rew, self.term, self.next_obs = obs, act, rew, term, next_obs
        self.T, self.N = rew.shape[0], rew.shape[1]

    def __len__(self):
        return self.N

    def __iter__(self):
        rew, act = self.rew.cpu().numpy(), self.act.cpu().numpy()
        for n in range(self.N):
            yield dict(rewards=rew[:, n].reshape(-1, 1), actions=act[:, n])


class BatchedPathCollector:
    """N envs stepped together.  collect_new_paths(max_path_length, num_steps, discard) runs rounds of N fresh episodes
    (rlkit starts every rollout with env.reset()) until num_steps transitions are gathered; a round's length is
    min(max_path_length, ceil(remaining / N)).  Incomplete rounds are dropped when discard_incomplete_paths is set."""

    def __init__(self, env, policy, deterministic=False):
        self._env, self._policy, self._det = env, policy, deterministic
        self._epoch_paths = []
        self._num_steps_total = self._num_paths_total = 0

    def collect_new_paths(self, max_path_length, num_steps, discard_incomplete_paths):
        import torch
        env, N = self._env, self._env.num_envs
        rounds, collected = [], 0
        while collected < num_steps:
            T = int(min(max_path_length, -(-(num_steps - collected) // N)))
            if T != max_path_length and discard_incomplete_paths:
                break
            o = env.reset()
            O, A = o.shape[1], env.action_space.low.size
            obs = torch.empty(T + 1, N, O, device=o.device)
            act, rew = torch.empty(T, N, A, device=o.device), torch.empty(T, N, device=o.device)
            term = torch.empty(T, N, dtype=torch.uint8, device=o.device)
            obs[0].copy_(o)
            for t in range(T):
                a = self._policy.get_actions(obs[t]) if not self._det else self._policy.get_actions(obs[t], deterministic=True)
                act[t].copy_(a)
                o2, r, d, _ = env.step(act[t])
                obs[t + 1].copy_(o2); rew[t].copy_(r); term[t].copy_(d)
            rounds.append(BatchedPaths(obs[:-1], act, rew, term, obs[1:]))
            collected += T * N
        self._num_paths_total += sum(r.N for r in rounds)
        self._num_steps_total += collected
        self._epoch_paths.extend(rounds)
        return rounds

    def get_epoch_paths(self):
        return [p for r in self._epoch_paths for p in r]

    def end_epoch(self, epoch):
        self._epoch_paths = []

    def get_diagnostics(self):
        st = OrderedDict([("num steps total", self._num_steps_total), ("num paths total", self._num_paths_total)])
        lens = np.concatenate([np.full(r.N, r.T) for r in self._epoch_paths]) if self._epoch_paths else np.zeros(0)
        st.update(create_stats_ordered_dict("path length", lens))
        return st

    def get_snapshot(self):
        return dict(policy=self._policy)


def add_paths(replay, paths):
    """replay_buffer.add_paths for both collectors."""
    if paths and isinstance(paths[0], BatchedPaths):
        for r in paths:
            T, N = r.T, r.N
            replay.add_batch(r.obs.reshape(T * N, -1), r.act.reshape(T * N, -1), r.rew.reshape(-1), r.term.reshape(-1), r.next_obs.reshape(T * N, -1))
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
        with open(os.path.join(self.dir, "params.pkl"), "wb") as f:
            pickle.dump(snapshot, f)


# ----------------------------------------------------------------------------- the algorithm
class BatchRLAlgorithm:
    """CustomTorchBatchRLAlgorithm (util/rlkit_custom.py:163-312): same constructor arguments, same `_train` order."""

    def __init__(self, trainer, exploration_env, evaluation_env, exploration_data_collector, evaluation_data_collector, replay_buffer,
                 batch_size, expl_max_path_length, eval_max_path_length, num_epochs, num_eval_steps_per_epoch,
                 num_expl_steps_per_train_loop, num_trains_per_train_loop, num_train_loops_per_epoch=1,
                 min_num_steps_before_training=0, logger=None):
        self.trainer, self.expl_env, self.eval_env = trainer, exploration_env, evaluation_env
        self.expl_data_collector, self.eval_data_collector, self.replay_buffer = exploration_data_collector, evaluation_data_collector, replay_buffer
        self.batch_size, self.expl_max_path_length, self.eval_max_path_length = batch_size, expl_max_path_length, eval_max_path_length
        self.num_epochs, self.num_eval_steps_per_epoch = num_epochs, num_eval_steps_per_epoch
        self.num_expl_steps_per_train_loop, self.num_trains_per_train_loop = num_expl_steps_per_train_loop, num_trains_per_train_loop
        self.num_train_loops_per_epoch, self.min_num_steps_before_training = num_train_loops_per_epoch, min_num_steps_before_training
        self._start_epoch, self.logger, self._t0 = 0, logger, time.time()
        self.trainer.replay = replay_buffer
        assert trainer.B == batch_size, "trainer was allocated for another batch size"

    def to(self, device):
        return self

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
        """util/rlkit_custom.py:68-82: trainer/exploration/evaluation entries, `env` entries skipped; state dicts instead of modules."""
        snap = {}
        for k, v in self.trainer.get_snapshot().items():
            snap["trainer/" + k] = v.state_dict()
        snap["exploration/policy"] = self.trainer.policy.state_dict()
        snap["evaluation/policy"] = self.trainer.policy.state_dict()
        return snap

    def _end_epoch(self, epoch, times, e0):
        t = time.time()
        if self.logger is not None:
            self.logger.save_itr_params(epoch, self._get_snapshot())
        times["saving"] = time.time() - t
        self._log_stats(epoch, times, e0)
        self.expl_data_collector.end_epoch(epoch)
        self.eval_data_collector.end_epoch(epoch)
        self.replay_buffer.end_epoch(epoch)
        self.trainer.end_epoch(epoch)

    def _log_stats(self, epoch, times, e0):
        lg, t = self.logger, time.time()
        if lg is None:
            return
        lg.log("Epoch {} finished".format(epoch))
        lg.record_dict(self.replay_buffer.get_diagnostics(), prefix="replay_buffer/")
        lg.record_dict(self.trainer.get_diagnostics(), prefix="trainer/")
        lg.record_dict(self.expl_data_collector.get_diagnostics(), prefix="exploration/")
        lg.record_dict(get_generic_path_information(self.expl_data_collector.get_epoch_paths()), prefix="exploration/")
        lg.record_dict(self.eval_data_collector.get_diagnostics(), prefix="evaluation/")
        lg.record_dict(get_custom_generic_path_information(self.eval_data_collector.get_epoch_paths(), self.expl_max_path_length,
                                                           self.trainer.reward_scale), prefix="evaluation/")
        times["logging"] = time.time() - t
        for k in ("data storing", "evaluation sampling", "exploration sampling", "logging", "saving", "training"):
            lg.record_tabular(f"time/{k} (s)", times.get(k, 0.0))
        lg.record_tabular("time/epoch (s)", time.time() - e0)
        lg.record_tabular("time/total (s)", time.time() - self._t0)
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


def experiment(variant, agent="SAC", num_envs=1, device="cuda:0", log_dir=None, seed=None, tf32=True):
    """util/rlkit_utils.py:31-165 on the batched backend.  num_envs=1 reproduces the reference's single-env data flow; num_envs=N
    runs N exploration envs and N evaluation envs together (the algorithm_kwargs step counts are then totals over all envs)."""
    from .sac import EnvReplayBuffer, MakeDeterministic, ParamStore, SACTrainer, TanhGaussianPolicy
    if agent != "SAC":
        raise NotImplementedError("only the SAC agent is on the benchmark path (TD3 is named by the reference but out of scope)")
    seed = variant.get("seed", 0) if seed is None else seed
    expl_env = make_env(variant["expl_environment_kwargs"], num_envs=num_envs, device=device, seed=seed, env_id_base=0)
    eval_env = make_env(variant["eval_environment_kwargs"], num_envs=num_envs, device=device, seed=seed, env_id_base=1 << 20)
    obs_dim, action_dim = expl_env.observation_space.low.size, expl_env.action_space.low.size
    assert list(variant["policy_kwargs"]["hidden_sizes"]) == [256, 256] and list(variant["qf_kwargs"]["hidden_sizes"]) == [256, 256], \
        "the fused update is built for the benchmark's 256x256 networks"
    store = ParamStore(obs_dim, action_dim, device, seed=seed)
    policy = TanhGaussianPolicy(store)
    replay = EnvReplayBuffer(variant["replay_buffer_size"], expl_env, device=device, seed=seed)
    ak = variant["algorithm_kwargs"]
    trainer = SACTrainer(env=eval_env, store=store, policy=policy, replay_buffer=replay, batch_size=ak["batch_size"], seed=seed, tf32=tf32,
                         **variant["trainer_kwargs"])
    if num_envs == 1:
        expl_c, eval_c = MdpPathCollector(expl_env, policy), MdpPathCollector(eval_env, MakeDeterministic(policy))
    else:
        expl_c, eval_c = BatchedPathCollector(expl_env, policy), BatchedPathCollector(eval_env, policy, deterministic=True)
    logger = Logger(log_dir, variant) if log_dir else None
    algo = BatchRLAlgorithm(trainer=trainer, exploration_env=expl_env, evaluation_env=eval_env, exploration_data_collector=expl_c,
                            evaluation_data_collector=eval_c, replay_buffer=replay, logger=logger, **ak)
    algo.train()
    return algo
