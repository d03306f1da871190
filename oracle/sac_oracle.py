"""CPU ORACLE for the SAC half of the hot path (TEST INFRASTRUCTURE, not product code).

Plain PyTorch (CPU, fp32, autograd) restatement of rlkit's SACTrainer.train_from_torch + TanhGaussianPolicy + FlattenMlp +
torch.optim.Adam + soft_update_from_to as the reference reaches them (util/rlkit_custom.py:235-238, util/rlkit_utils.py:64-106).
rlkit (pinned b7f97b2463df1c5a1ecd2d293cfcc7a4971dd0ab, README.md:28; the committed runs record d63dab7 + patch) is NOT
vendored under /root/reference and not installed here, so this follows SURVEY.md A.4.  It IS pinned against the reference's
own artefacts: the epoch-0 `trainer/*` known answers logged in runs/*/progress.csv -- `Alpha` after the first
update = exp(-policy_lr) = 0.9990004897117615 (fp32), `Alpha Loss` = -0.0 -- checked in tests/test_sac_oracle.py against
tests/golden/sac_epoch0_known_answers.json (extracted from the committed runs by tests/golden/make_sac_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

LOG_SIG_MAX, LOG_SIG_MIN = 2.0, -20.0


def replay_indices(seed, step, batch, size):
    """The replay index rule of csrc/rsb_sac.cu restated on the host: Philox word0 of counter (row, step_lo, step_hi, 0xB0FFE7)
    under key `seed`, mapped to [0, size) by the high half of a 32x32-bit product (with replacement, like np.random.randint)."""
    from robosuite_benchmark_b200.philox import philox4x32
    out = np.empty(batch, np.int64)
    for b in range(batch):
        w = philox4x32([b, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, 0xB0FFE7], [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
        out[b] = (w[0] * int(size)) >> 32
    return out


def policy_noise(seed, env_ids, step, act_dim):
    """The collector's exploration noise restated on the host (csrc/rsb_collect.cu k_policy_act): for env id e and action dims 4 blk .. 4 blk + 3,
    Philox4x32-10(key = seed, counter = (e_lo, e_hi, 2, step * 8 + blk)) -> two Box-Muller pairs (arguments in double, result cast to fp32)."""
    from robosuite_benchmark_b200.philox import philox4x32
    out = np.zeros((len(env_ids), act_dim), np.float32)
    key = [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]
    for i, e in enumerate(env_ids):
        e = int(e)
        for blk in range((act_dim + 3) // 4):
            w = philox4x32([e & 0xFFFFFFFF, (e >> 32) & 0xFFFFFFFF, 2, (step * 8 + blk) & 0xFFFFFFFF], key)
            z = []
            for a, b in ((w[0], w[1]), (w[2], w[3])):
                u1, u2 = (a + 0.5) / 4294967296.0, (b + 0.5) / 4294967296.0
                rad, ang = np.sqrt(-2.0 * np.log(u1)), 2.0 * np.pi * u2
                z += [rad * np.cos(ang), rad * np.sin(ang)]
            for k in range(4):
                if 4 * blk + k < act_dim:
                    out[i, 4 * blk + k] = np.float32(z[k])
    return out


def policy_act(params, obs, seed=0, env_ids=None, step=0, deterministic=False):
    """rlkit TanhGaussianPolicy.get_action for a batch of observations in fp64 numpy (weights [in, out] as in the ParamStore): the checker of
    k_policy_act.  a = tanh(mean) (MakeDeterministic) or tanh(mean + exp(clamp(log_std, -20, 2)) * eps) with eps = policy_noise(...)."""
    g = lambda k: np.asarray(params[k], np.float64)
    x = np.asarray(obs, np.float64)
    h = np.maximum(x @ g("p_W0") + g("p_b0"), 0.0)
    h = np.maximum(h @ g("p_W1") + g("p_b1"), 0.0)
    out = h @ g("p_W2") + g("p_b2")
    A = out.shape[1] // 2
    mean, log_std = out[:, :A], np.clip(out[:, A:], LOG_SIG_MIN, LOG_SIG_MAX)
    if deterministic:
        return np.tanh(mean)
    env_ids = np.arange(len(x)) if env_ids is None else env_ids
    return np.tanh(mean + np.exp(log_std) * policy_noise(seed, env_ids, step, A).astype(np.float64))


def tf32_operand(x, mode):
    """fp32 -> the value a TF32 tensor-core product uses for this operand: sign, 8 exponent bits, 10 mantissa bits.
    "trunc": the low 13 mantissa bits are dropped (what `tcgen05.mma.kind::tf32` does with raw fp32 operands in shared memory);
    "rna": round to nearest, ties away from zero (`cvt.rna.tf32.f32`, what cuBLAS applies before `mma.sync`)."""
    i = x.detach().contiguous().view(torch.int32)
    if mode == "rna":
        i = i + 0x1000
    elif mode != "trunc":
        raise ValueError(mode)
    return (i & ~0x1FFF).view(torch.float32)


class _TF32MatMul(torch.autograd.Function):
    """a @ b as the product path computes it: the operands of a product that runs on the tensor cores are reduced to TF32, products and sums
    are exact (fp64 here; fp32 accumulation on the device).  `which` = (forward, input gradient, weight gradient) says which of the three
    products of this layer are tensor-core products: (True, True, True) for the 256-wide layers; (False, False, True) for the narrow last
    layers, whose forward and input gradient run in exact fp32 inside the fused CUDA-core kernels (csrc/rsb_sac_fused.cu)."""

    @staticmethod
    def forward(ctx, a, b, mode, which):
        ctx.save_for_backward(a, b)
        ctx.mode, ctx.which = mode, which
        r = (lambda x: tf32_operand(x, mode)) if which[0] else (lambda x: x.detach())
        return (r(a).double() @ r(b).double()).float()

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        m = ctx.mode
        ra = (lambda x: tf32_operand(x, m)) if ctx.which[1] else (lambda x: x.detach())
        rb = (lambda x: tf32_operand(x, m)) if ctx.which[2] else (lambda x: x.detach())
        ga = (ra(g).double() @ ra(b).double().transpose(-1, -2)).float()
        gb = (rb(a).double().transpose(-1, -2) @ rb(g).double()).float()
        while gb.dim() > b.dim():
            gb = gb.sum(0)
        while ga.dim() > a.dim():
            ga = ga.sum(0)
        return ga, gb, None, None


class SacOracle:
    """Parameters in the store's host layout (weights as [in, out]; twin Q stacked on a leading axis of 2)."""

    def __init__(self, params, targets, obs_dim, act_dim, discount=0.99, reward_scale=1.0, policy_lr=1e-3, qf_lr=1e-3,
                 soft_target_tau=1e-2, target_update_period=1, target_entropy=None, tf32=None):
        """tf32: None = plain fp32 products (rlkit on a CPU); "trunc" / "rna" = every product with TF32 operands (tf32_operand), the
        arithmetic of the tensor-core product path -- used to hold that path to a tight tolerance instead of the loose TF32 error bound."""
        self.O, self.A, self.tf32 = obs_dim, act_dim, tf32
        self.p = {k: torch.tensor(np.asarray(v, np.float32), requires_grad=True) for k, v in params.items()}
        self.t = {k: torch.tensor(np.asarray(v, np.float32)) for k, v in targets.items()}
        self.discount, self.reward_scale, self.tau, self.period = discount, reward_scale, soft_target_tau, target_update_period
        self.target_entropy = float(-act_dim if target_entropy is None else target_entropy)
        pol = [self.p[k] for k in ("p_W0", "p_b0", "p_W1", "p_b1", "p_W2", "p_b2")]
        qs = [self.p[k] for k in ("q_W0", "q_b0", "q_W1", "q_b1", "q_W2", "q_b2")]
        # rlkit keeps separate Adam instances for qf1 and qf2; Adam is element-wise so one instance over the stacked tensors is identical
        self.opt_pi, self.opt_q = torch.optim.Adam(pol, lr=policy_lr), torch.optim.Adam(qs, lr=qf_lr)
        self.opt_alpha = torch.optim.Adam([self.p["log_alpha"]], lr=policy_lr)
        self.n_steps, self.stats = 0, {}

    def mm(self, a, b, narrow=False):
        """narrow: a last layer (256 -> 2A / 256 -> 1)."""
        if self.tf32 is None:
            return a @ b
        return _TF32MatMul.apply(a, b, self.tf32, (False, False, True) if narrow else (True, True, True))

    def policy(self, obs, eps):
        p, A, mm = self.p, self.A, self.mm
        h = F.relu(mm(obs, p["p_W0"]) + p["p_b0"]); h = F.relu(mm(h, p["p_W1"]) + p["p_b1"]); out = mm(h, p["p_W2"], narrow=True) + p["p_b2"]
        mean, log_std = out[:, :A], out[:, A:].clamp(LOG_SIG_MIN, LOG_SIG_MAX)
        z = mean + log_std.exp() * eps
        a = torch.tanh(z)
        logp = (-0.5 * eps ** 2 - log_std - 0.5 * np.log(2 * np.pi) - torch.log(1 - a * a + 1e-6)).sum(1, keepdim=True)
        return a, logp, mean, log_std

    def qpair(self, w, obs, act):
        x = torch.cat([obs, act], 1)
        h = F.relu(self.mm(x.unsqueeze(0).expand(2, -1, -1), w["q_W0"]) + w["q_b0"][:, None, :])
        h = F.relu(self.mm(h, w["q_W1"]) + w["q_b1"][:, None, :])
        return self.mm(h, w["q_W2"], narrow=True) + w["q_b2"][:, None, :]          # [2, B, 1]

    def train(self, batch, eps):
        """batch: dict of numpy arrays (rlkit keys); eps: [2B, A] -- rows [0,B) drive pi(obs), rows [B,2B) drive pi(next_obs)."""
        g = lambda k: torch.tensor(np.asarray(batch[k], np.float32))
        obs, act, nxt = g("observations"), g("actions"), g("next_observations")
        rew, term = g("rewards").reshape(-1, 1), g("terminals").reshape(-1, 1)
        B = obs.shape[0]
        eps = torch.tensor(np.asarray(eps, np.float32))
        p = self.p
        a_new, logpi, mean, log_std = self.policy(obs, eps[:B])
        alpha_loss = -(p["log_alpha"] * (logpi + self.target_entropy).detach()).mean()
        alpha = p["log_alpha"].exp().detach()                      # PRE-update alpha enters the losses
        qn = self.qpair(p, obs, a_new)
        policy_loss = (alpha * logpi - torch.min(qn[0], qn[1])).mean()
        qpred = self.qpair(p, obs, act)
        with torch.no_grad():
            a2, logpi2, _, _ = self.policy(nxt, eps[B:])
            qt = self.qpair(self.t, nxt, a2)
            y = self.reward_scale * rew + (1.0 - term) * self.discount * (torch.min(qt[0], qt[1]) - alpha * logpi2)
        qf1_loss, qf2_loss = F.mse_loss(qpred[0], y), F.mse_loss(qpred[1], y)
        self.opt_alpha.zero_grad(); alpha_loss.backward(); self.opt_alpha.step()
        self.opt_pi.zero_grad(); self.opt_q.zero_grad()
        policy_loss.backward()                                     # also deposits into the Q grads: cleared before the Q backward
        grads = {k: p[k].grad.clone() for k in ("p_W0", "p_b0", "p_W1", "p_b1", "p_W2", "p_b2")}
        self.opt_pi.step()
        self.opt_q.zero_grad(); (qf1_loss + qf2_loss).backward()
        grads.update({k: p[k].grad.clone() for k in ("q_W0", "q_b0", "q_W1", "q_b1", "q_W2", "q_b2")})
        self.opt_q.step()
        if self.n_steps % self.period == 0:
            with torch.no_grad():
                for k in self.t:
                    self.t[k].mul_(1 - self.tau).add_(self.tau * p[k].detach())
        self.n_steps += 1
        self.stats = {"QF1 Loss": qf1_loss.item(), "QF2 Loss": qf2_loss.item(), "Policy Loss": policy_loss.item(),
                      "Alpha": p["log_alpha"].exp().item(), "Alpha Loss": alpha_loss.item(), "Log Pis Mean": logpi.mean().item(),
                      "Q Targets Mean": y.mean().item(), "Q1 Predictions Mean": qpred[0].mean().item()}
        return grads

    def params(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}, {k: v.numpy().copy() for k, v in self.t.items()}
