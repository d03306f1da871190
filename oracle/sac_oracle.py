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


class SacOracle:
    """Parameters in the store's host layout (weights as [in, out]; twin Q stacked on a leading axis of 2)."""

    def __init__(self, params, targets, obs_dim, act_dim, discount=0.99, reward_scale=1.0, policy_lr=1e-3, qf_lr=1e-3,
                 soft_target_tau=1e-2, target_update_period=1, target_entropy=None):
        self.O, self.A = obs_dim, act_dim
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

    def policy(self, obs, eps):
        p, A = self.p, self.A
        h = F.relu(obs @ p["p_W0"] + p["p_b0"]); h = F.relu(h @ p["p_W1"] + p["p_b1"]); out = h @ p["p_W2"] + p["p_b2"]
        mean, log_std = out[:, :A], out[:, A:].clamp(LOG_SIG_MIN, LOG_SIG_MAX)
        z = mean + log_std.exp() * eps
        a = torch.tanh(z)
        logp = (-0.5 * eps ** 2 - log_std - 0.5 * np.log(2 * np.pi) - torch.log(1 - a * a + 1e-6)).sum(1, keepdim=True)
        return a, logp, mean, log_std

    @staticmethod
    def qpair(w, obs, act):
        x = torch.cat([obs, act], 1)
        h = F.relu(torch.einsum("bi,nio->nbo", x, w["q_W0"]) + w["q_b0"][:, None, :])
        h = F.relu(torch.bmm(h, w["q_W1"]) + w["q_b1"][:, None, :])
        return torch.bmm(h, w["q_W2"]) + w["q_b2"][:, None, :]                     # [2, B, 1]

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
