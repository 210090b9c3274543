"""CPU restatement of the reference's TD3 gradient step (TEST INFRASTRUCTURE — only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product path never does).

Follows, line by line:
  TD3.train                      core/td3/td3.py:154-211   (target smoothing, twin-min target, MSE sum, delayed actor, polyak)
  create_mlp / Actor / critic    core/common/torch_layers.py:110-183, core/td3/policies.py:58,75-78, core/common/policies.py:966-987
  polyak_update                  core/common/utils.py:457-481
  torch.optim.Adam (defaults)    betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad  (torch/optim/adam.py, single-tensor path)
Further down: SACUpdateOracle (core/sac/sac.py:199-296) and BCQUpdateOracle (core/bcq/bcq.py:129-205, pinned by
tests/golden/bcq_update.npz) and MultiAgentDDPGOracle (core/maddpg/maddpg.py, core/iddpg/iddpg.py :131-185, pinned by
tests/golden/{maddpg,iddpg}_update.npz); their CUDA counterparts are cstr_bcq_update / cstr_ma_update (tests/test_gpu_bcq_ma.py).
Everything is float32 NumPy with hand-written backward passes.  Pinned against the unmodified reference running on CPU torch
(tests/golden/td3_update.npz, made by oracle/make_golden.py::gen_td3_update): weights agree to ~1e-6 after 6 gradient steps
(GEMM summation order is the only difference), tolerance written in tests/test_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
Params = List[np.ndarray]  # [W1 (H1,in), b1, W2 (H2,H1), b2, W3 (out,H2), b3] float32, torch nn.Linear layout


def mlp_forward(p: Params, x: np.ndarray, squash: bool):
    """nn.Sequential(Linear, ReLU, Linear, ReLU, Linear[, Tanh]) — returns (y, cache)."""
    W1, b1, W2, b2, W3, b3 = p
    h1 = np.maximum(x @ W1.T + b1, F32(0))
    h2 = np.maximum(h1 @ W2.T + b2, F32(0))
    y = h2 @ W3.T + b3
    if squash:
        y = np.tanh(y)
    return y.astype(F32), (x, h1, h2, y.astype(F32))


def mlp_backward(p: Params, cache, dy: np.ndarray, squash: bool, need_dx: bool = False):
    """Gradients of sum(dy * y) w.r.t. the six tensors (and the input)."""
    W1, b1, W2, b2, W3, b3 = p
    x, h1, h2, y = cache
    if squash:
        dy = dy * (F32(1) - y * y)
    dW3, db3 = dy.T @ h2, dy.sum(0)
    dz2 = (dy @ W3) * (h2 > 0)
    dW2, db2 = dz2.T @ h1, dz2.sum(0)
    dz1 = (dz2 @ W2) * (h1 > 0)
    dW1, db1 = dz1.T @ x, dz1.sum(0)
    dx = dz1 @ W1 if need_dx else None
    return [g.astype(F32) for g in (dW1, db1, dW2, db2, dW3, db3)], dx


class Adam:
    """torch.optim.Adam, single-tensor path: lerp / addcmul / sqrt / addcdiv in float32, scalar factors in Python floats."""

    def __init__(self, params: Params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params, self.lr, self.betas, self.eps = params, lr, betas, eps
        self.m = [np.zeros_like(p) for p in params]
        self.v = [np.zeros_like(p) for p in params]
        self.step_count = 0

    def step(self, grads: Sequence[np.ndarray]) -> None:
        self.step_count += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.step_count
        bc2_sqrt = math.sqrt(1 - b2 ** self.step_count)
        step_size = self.lr / bc1
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            m += (g - m) * F32(1 - b1)  # exp_avg.lerp_(grad, 1 - beta1)
            v *= F32(b2)
            v += F32(1 - b2) * g * g  # addcmul_(grad, grad, value=1 - beta2)
            denom = np.sqrt(v) / F32(bc2_sqrt) + F32(self.eps)
            p += F32(-step_size) * (m / denom)  # addcdiv_(exp_avg, denom, value=-step_size)


def polyak(params: Params, targets: Params, tau: float) -> None:
    for p, t in zip(params, targets):
        t *= F32(1 - tau)
        t += F32(tau) * p


class TD3UpdateOracle:
    def __init__(self, actor: Params, critics: Sequence[Params], actor_target: Optional[Params] = None, critic_targets=None, lr: float = 1e-3,
                 gamma: float = 0.99, tau: float = 0.005, policy_delay: int = 2, target_noise_clip: float = 0.5):
        cp = lambda ps: [np.array(a, F32, copy=True) for a in ps]  # noqa: E731
        self.actor, self.critics = cp(actor), [cp(c) for c in critics]
        self.actor_target = cp(actor_target if actor_target is not None else actor)
        self.critic_targets = [cp(c) for c in (critic_targets if critic_targets is not None else critics)]
        self.gamma, self.tau, self.policy_delay, self.noise_clip = gamma, tau, policy_delay, target_noise_clip
        self.actor_opt = Adam(self.actor, lr)
        self.critic_opt = Adam([t for c in self.critics for t in c], lr)  # one optimizer over both q-networks (policies.py:200-204)
        self.n_updates = 0
        self.critic_losses: List[float] = []
        self.actor_losses: List[float] = []

    def step(self, obs, actions, next_obs, dones, rewards, noise) -> Dict[str, np.ndarray]:
        """One iteration of the loop body td3.py:162-206.  ``noise`` is the N(0, target_policy_noise) draw of :168 (unclipped)."""
        self.n_updates += 1
        B = obs.shape[0]
        nz = np.clip(noise, F32(-self.noise_clip), F32(self.noise_clip))  # :169
        na, _ = mlp_forward(self.actor_target, next_obs, True)
        next_actions = np.clip(na + nz, F32(-1), F32(1))  # :170
        xin = np.concatenate([next_obs, next_actions], 1)
        q_next = np.minimum.reduce([mlp_forward(c, xin, False)[0] for c in self.critic_targets])  # :173-174 (one critic: DDPG)
        target = (rewards + (F32(1) - dones) * F32(self.gamma) * q_next).astype(F32)  # :175
        x = np.concatenate([obs, actions], 1)
        grads, loss, hidden = [], 0.0, []
        for c in self.critics:
            q, cache = mlp_forward(c, x, False)
            hidden.append((cache[1], cache[2]))
            diff = q - target
            loss += float(np.mean(diff * diff, dtype=F32))  # F.mse_loss, summed over critics (:181)
            g, _ = mlp_backward(c, cache, (F32(2) / F32(B)) * diff, False)
            grads += g
        self.critic_losses.append(loss)
        self.critic_opt.step(grads)
        out = {"target_q": target, "critic_grads": grads, "critic_hidden": hidden}
        if self.n_updates % self.policy_delay == 0:  # :189
            a, acache = mlp_forward(self.actor, obs, True)
            q1, ccache = mlp_forward(self.critics[0], np.concatenate([obs, a], 1), False)  # q1_forward (:191)
            self.actor_losses.append(float(-np.mean(q1, dtype=F32)))
            _, dx = mlp_backward(self.critics[0], ccache, np.full_like(q1, F32(-1) / F32(B)), False, need_dx=True)
            ag, _ = mlp_backward(self.actor, acache, dx[:, obs.shape[1]:].astype(F32), True)
            self.actor_opt.step(ag)
            for c, t in zip(self.critics, self.critic_targets):  # :199-200
                polyak(c, t, self.tau)
            polyak(self.actor, self.actor_target, self.tau)
            out["actor_grads"] = ag
            out["policy_hidden"] = [(acache[1], acache[2]), (ccache[1], ccache[2])]  # actor, critic0 at pi(s)
        return out


# ----------------------------------------------------------------------------------------------
# SAC gradient step (core/sac/sac.py:199-296, core/sac/policies.py:147-175, core/common/distributions.py:207-260)
# ----------------------------------------------------------------------------------------------
LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0
HALF_LOG_2PI = F32(0.5 * math.log(2.0 * math.pi))


def sac_actor_forward(p: Params, obs: np.ndarray, eps: np.ndarray, squash_eps: float = 1e-6):
    """Actor = latent MLP (2 ReLU layers) + [mu; log_std] head stacked as one (2A, H2) matrix (rows 0..A-1 = mu, A.. = log_std).
    Returns (action, log_prob, cache) for the reparameterised sample u = mean + std * eps, a = tanh(u)."""
    y, cache = mlp_forward(p, obs, False)
    A = y.shape[1] // 2
    mean, raw = y[:, :A], y[:, A:]
    log_std = np.clip(raw, F32(LOG_STD_MIN), F32(LOG_STD_MAX))
    std = np.exp(log_std)
    u = mean + std * eps
    a = np.tanh(u)
    # Normal(mean, std).log_prob(u) summed over dims (with (u - mean)/std computed as torch does), minus the tanh correction
    logp = (-((u - mean) ** 2) / (F32(2) * std * std) - log_std - HALF_LOG_2PI).sum(1) - np.log(F32(1) - a * a + F32(squash_eps)).sum(1)
    return a.astype(F32), logp.astype(F32), (cache, raw, std.astype(F32), a.astype(F32))


class SACUpdateOracle:
    def __init__(self, actor: Params, critics: Sequence[Params], critic_targets=None, lr: float = 3e-4, gamma: float = 0.99, tau: float = 0.005,
                 target_entropy: float = -2.0, log_ent_coef: float = 0.0, target_update_interval: int = 1):
        cp = lambda ps: [np.array(a, F32, copy=True) for a in ps]  # noqa: E731
        self.actor, self.critics = cp(actor), [cp(c) for c in critics]
        self.critic_targets = [cp(c) for c in (critic_targets if critic_targets is not None else critics)]
        self.gamma, self.tau, self.target_entropy, self.interval = gamma, tau, target_entropy, target_update_interval
        self.log_ent_coef = [np.array([log_ent_coef], F32)]
        self.actor_opt = Adam(self.actor, lr)
        self.critic_opt = Adam([t for c in self.critics for t in c], lr)
        self.ent_opt = Adam(self.log_ent_coef, lr)
        self.gradient_step = 0
        self.critic_losses, self.actor_losses, self.ent_coef_losses, self.ent_coefs = [], [], [], []

    def step(self, obs, actions, next_obs, dones, rewards, eps_pi, eps_next) -> Dict[str, np.ndarray]:
        """One iteration of sac.py:213-288.  eps_pi / eps_next are the standard-normal draws of the two rsample() calls."""
        B, A = obs.shape[0], actions.shape[1]
        a_pi, logp, (acache, raw, std, _) = sac_actor_forward(self.actor, obs, eps_pi)
        logp = logp.reshape(-1, 1)
        ent_coef = np.exp(self.log_ent_coef[0]).astype(F32)  # :230 (before the optimiser step)
        self.ent_coefs.append(float(ent_coef[0]))
        self.ent_coef_losses.append(float(-np.mean(self.log_ent_coef[0] * (logp + F32(self.target_entropy)), dtype=F32)))
        self.ent_opt.step([np.array([-np.mean(logp + F32(self.target_entropy), dtype=F32)], F32)])  # :231-243
        na, nlogp, _ = sac_actor_forward(self.actor, next_obs, eps_next)
        xin = np.concatenate([next_obs, na], 1)
        q_next = np.minimum(*[mlp_forward(c, xin, False)[0] for c in self.critic_targets]) - ent_coef * nlogp.reshape(-1, 1)  # :249-252
        target = (rewards + (F32(1) - dones) * F32(self.gamma) * q_next).astype(F32)
        x = np.concatenate([obs, actions], 1)
        grads, loss = [], 0.0
        for c in self.critics:
            q, cache = mlp_forward(c, x, False)
            diff = q - target
            loss += 0.5 * float(np.mean(diff * diff, dtype=F32))  # :261
            g, _ = mlp_backward(c, cache, (F32(1) / F32(B)) * diff, False)
            grads += g
        self.critic_losses.append(loss)
        self.critic_opt.step(grads)
        # actor loss (ent_coef * log_prob - min_i Q_i(s, a_pi)).mean() with the UPDATED critics (:271-281)
        xp = np.concatenate([obs, a_pi], 1)
        qs, caches = zip(*[mlp_forward(c, xp, False) for c in self.critics])
        q_min = np.minimum(*qs)
        self.actor_losses.append(float(np.mean(ent_coef * logp - q_min, dtype=F32)))
        pick0 = qs[0] <= qs[1]  # torch.min returns the first index on ties
        da = np.zeros_like(a_pi)
        for i, c in enumerate(self.critics):
            dq = np.where(pick0 if i == 0 else ~pick0, F32(-1) / F32(B), F32(0)).astype(F32)
            _, dx = mlp_backward(c, caches[i], dq, False, need_dx=True)
            da += dx[:, obs.shape[1]:]
        a = a_pi
        one_m = F32(1) - a * a
        dlogp = ent_coef / F32(B)  # d loss / d log_prob, per row
        du = da * one_m + dlogp * (F32(2) * a * one_m / (one_m + F32(1e-6)))  # through tanh: Q path + squash correction
        dmean = du  # the Gaussian log-density is constant in mean under reparameterisation
        dlog_std = (du * std * eps_pi - dlogp) * ((raw >= F32(LOG_STD_MIN)) & (raw <= F32(LOG_STD_MAX)))  # d(-log_std) = -1; clamp gate
        ag, _ = mlp_backward(self.actor, acache, np.concatenate([dmean, dlog_std], 1).astype(F32), False)
        self.actor_opt.step(ag)
        if self.gradient_step % self.interval == 0:
            for c, t in zip(self.critics, self.critic_targets):
                polyak(c, t, self.tau)
        self.gradient_step += 1
        return {"target_q": target, "critic_grads": grads, "actor_grads": ag, "log_prob": logp, "a_pi": a_pi}


class BCQUpdateOracle:
    """The loop body of ``BCQ.train`` (core/bcq/bcq.py:137-205) with the networks of core/bcq/policies.py:

      vae_enc  [W1,b1,W2,b2,Wh,bh]   encoder (obs+act -> H -> H) and the two heads stacked, Wh = [mean; log_std] (2L, H)   policies.py:47-55
      vae_dec  [W1..b3]              decoder (obs+L -> H -> H -> act, Tanh)                                                 :58-65
      pert     [W1..b3]              perturbation net (obs+act -> P -> P -> act, Tanh) * max_perturbation, clamp(-1, 1)     :147-160
      critics  2 x [W1..b3]          ContinuousCritic q-networks (obs+act -> h1 -> h2 -> 1)

    Three Adam optimisers (policies.py:357-382): the VAE's (every actor parameter that is not the perturbation net's), the
    perturbation net's, the critics'.  Targets: perturbation net and critics by polyak on actor steps; the target VAE is a plain copy of the
    VAE, refreshed after every VAE step (bcq.py:158-159, 198).  The random draws are arguments (standard normal, UNclamped):
    ``eps_vae`` (B, L) of :76 randn_like; ``z_next`` (n_candidates*B, L) of sample_action's randn (:122, clamped to +-0.5 there);
    ``z_actor`` (B, L) of the actor step's randn.  NOTE bcq.py:170-171 reshapes the (n_candidates*B, 1) q column — whose rows are ordered
    candidate-major by ``repeat`` — to (B, n_candidates) row-major, so the max runs over n_candidates CONSECUTIVE rows of the tiled batch,
    not over the candidates of one observation; restated as written."""

    def __init__(self, vae_enc: Params, vae_dec: Params, pert: Params, critics: Sequence[Params], lr: float = 1e-3, gamma: float = 0.99,
                 tau: float = 0.005, max_perturbation: float = 0.05, actor_delay: int = 2, n_candidates: int = 10):
        cp = lambda ps: [np.array(a, F32, copy=True) for a in ps]  # noqa: E731
        self.vae_enc, self.vae_dec, self.pert, self.critics = cp(vae_enc), cp(vae_dec), cp(pert), [cp(c) for c in critics]
        self.pert_target, self.critic_targets = cp(pert), [cp(c) for c in critics]
        self.L = self.vae_enc[4].shape[0] // 2
        self.gamma, self.tau, self.phi, self.actor_delay, self.n_candidates = gamma, tau, F32(max_perturbation), actor_delay, n_candidates
        self.vae_opt = Adam(self.vae_enc + self.vae_dec, lr)
        self.pert_opt = Adam(self.pert, lr)
        self.critic_opt = Adam([t for c in self.critics for t in c], lr)
        self.n_updates = 0
        self.vae_losses: List[float] = []
        self.critic_losses: List[float] = []
        self.actor_losses: List[float] = []

    def _decode(self, dec: Params, s, z_raw):
        z = np.clip(z_raw, F32(-0.5), F32(0.5))  # policies.py:111,122
        return mlp_forward(dec, np.concatenate([s, z], 1), True)

    def _perturb(self, pert: Params, s, a):
        xi, cache = mlp_forward(pert, np.concatenate([s, a], 1), True)
        pre = a + xi * self.phi
        return np.clip(pre, F32(-1), F32(1)).astype(F32), (cache, pre)

    def step(self, obs, actions, next_obs, dones, rewards, eps_vae, z_next, z_actor=None) -> Dict[str, np.ndarray]:
        self.n_updates += 1
        B, L = obs.shape[0], self.L
        # ---- VAE (bcq.py:142-155): recon MSE + 0.5 * KL ----
        enc_out, ecache = mlp_forward(self.vae_enc, np.concatenate([obs, actions], 1), False)
        mean, raw_ls = enc_out[:, :L], enc_out[:, L:]
        log_std = np.clip(raw_ls, F32(-4), F32(15))
        std = np.exp(log_std).astype(F32)
        z = (mean + std * eps_vae).astype(F32)
        recon, dcache = mlp_forward(self.vae_dec, np.concatenate([obs, z], 1), True)
        diff = recon - actions
        recon_loss = np.mean(diff * diff, dtype=F32)
        kl = F32(-0.5) * np.mean(F32(1) + np.log(std * std) - mean * mean - std * std, dtype=F32)
        self.vae_losses.append(float(recon_loss + F32(0.5) * kl))
        dgrads, dx = mlp_backward(self.vae_dec, dcache, (F32(2) / F32(diff.size)) * diff, True, need_dx=True)
        dz = dx[:, obs.shape[1]:]
        n = F32(B * L)
        dmean = dz + F32(0.5) * mean / n
        dstd = dz * eps_vae + F32(0.5) * (std - F32(1) / std) / n
        draw = dstd * std * ((raw_ls >= F32(-4)) & (raw_ls <= F32(15)))  # clamp passes the gradient inside (and at) the bounds
        egrads, _ = mlp_backward(self.vae_enc, ecache, np.concatenate([dmean, draw], 1).astype(F32), False)
        self.vae_opt.step(egrads + dgrads)
        out = {"vae_grads": egrads + dgrads}
        # ---- target (bcq.py:157-172): candidates from the (just refreshed) target VAE + target perturbation net ----
        K = self.n_candidates
        s_rep = np.tile(next_obs, (K, 1))
        cand, _ = self._decode(self.vae_dec, s_rep, z_next)
        cand_p, _ = self._perturb(self.pert_target, s_rep, cand)
        xin = np.concatenate([s_rep, cand_p], 1)
        q = np.minimum.reduce([mlp_forward(c, xin, False)[0] for c in self.critic_targets])
        q = q.reshape(B, K).max(1)[:, None]  # as written at :170-171 (see the class docstring)
        target = (rewards + (F32(1) - dones) * F32(self.gamma) * q).astype(F32)
        # ---- critics (bcq.py:174-186) ----
        x = np.concatenate([obs, actions], 1)
        grads, loss = [], 0.0
        for c in self.critics:
            qc, cache = mlp_forward(c, x, False)
            d = qc - target
            loss += float(np.mean(d * d, dtype=F32))
            g, _ = mlp_backward(c, cache, (F32(2) / F32(B)) * d, False)
            grads += g
        self.critic_losses.append(loss)
        self.critic_opt.step(grads)
        out.update(target_q=target, critic_grads=grads)
        # ---- delayed perturbation-actor step (bcq.py:188-203) ----
        if self.n_updates % self.actor_delay == 0:
            a0, _ = self._decode(self.vae_dec, obs, z_actor)
            a, (pcache, pre) = self._perturb(self.pert, obs, a0)
            q1, ccache = mlp_forward(self.critics[0], np.concatenate([obs, a], 1), False)
            self.actor_losses.append(float(-np.mean(q1, dtype=F32)))
            _, dx = mlp_backward(self.critics[0], ccache, np.full_like(q1, F32(-1) / F32(B)), False, need_dx=True)
            da = dx[:, obs.shape[1]:] * ((pre >= F32(-1)) & (pre <= F32(1)))
            pg, _ = mlp_backward(self.pert, pcache, (da * self.phi).astype(F32), True)  # only the perturbation optimiser steps (:196)
            self.pert_opt.step(pg)
            for c, t in zip(self.critics, self.critic_targets):
                polyak(c, t, self.tau)
            polyak(self.pert, self.pert_target, self.tau)  # the VAE half of polyak_update(actor) is overwritten by the copy at :198
            out["pert_grads"] = pg
        return out


class MultiAgentDDPGOracle:
    """The loop body of ``MADDPG.train`` (core/maddpg/maddpg.py:131-185; ``centralised=True``) and of ``IDDPG.train``
    (core/iddpg/iddpg.py:131-185, the same text; ``centralised=False``).  Per agent i: an actor on the agent's observation slice
    (policies.py:64-72) and ``n_critics`` q-networks — over ALL observations and ALL actions for MADDPG (maddpg/policies.py:236-241), over the
    agent's own slices for IDDPG (iddpg/policies.py:135) — each with targets; one Adam per agent for the actor and one for its critics.

    Restated as written, including: the target actions of every agent are formed once, before the agent loop (:132-144); the polyak update of
    ALL actors and critics sits inside the agent loop (:181-182), so it runs once per agent on an actor step and agent i+1's critic target
    already sees it; the actor step evaluates EVERY actor on agent i's observation slice (:169-171 ``mu_list[id](agent_observations)``).
    Learning rates: ``train()`` calls ``_update_learning_rate([actor_opt_i, critic_opt_i])`` per agent (:121-123) and that method
    (core/common/base_class.py:1112-1136) pairs the k-th list entry with ``lr_schedule_list[k]`` — so with two agents EVERY actor runs at
    ``learning_rate_list[0]`` and EVERY critic at ``learning_rate_list[1]``; ``reference_lrs`` returns that assignment."""

    @staticmethod
    def reference_lrs(learning_rate_list: Sequence[float]):
        n = len(learning_rate_list)
        return [learning_rate_list[0]] * n, [learning_rate_list[1]] * n

    def __init__(self, actors: Sequence[Params], critics: Sequence[Sequence[Params]], obs_splits, act_splits, centralised: bool,
                 actor_lrs: Sequence[float], critic_lrs: Sequence[float], gamma: float = 0.99, tau: float = 0.005, policy_delay: int = 2,
                 target_noise_clip: float = 0.5):
        cp = lambda ps: [np.array(a, F32, copy=True) for a in ps]  # noqa: E731
        self.actors, self.critics = [cp(a) for a in actors], [[cp(q) for q in qs] for qs in critics]
        self.actor_targets, self.critic_targets = [cp(a) for a in actors], [[cp(q) for q in qs] for qs in critics]
        self.obs_splits, self.act_splits, self.centralised = [list(s) for s in obs_splits], [list(s) for s in act_splits], centralised
        self.gamma, self.tau, self.policy_delay, self.noise_clip = gamma, tau, policy_delay, target_noise_clip
        self.actor_opts = [Adam(a, lr) for a, lr in zip(self.actors, actor_lrs)]
        self.critic_opts = [Adam([t for q in qs for t in q], lr) for qs, lr in zip(self.critics, critic_lrs)]
        self.n = len(self.actors)
        self.n_updates = 0
        self.critic_losses: List[List[float]] = [[] for _ in range(self.n)]
        self.actor_losses: List[List[float]] = [[] for _ in range(self.n)]

    def _critic_input(self, i, obs, acts):
        if self.centralised:
            return np.concatenate([obs, acts], 1), obs.shape[1] + self.act_splits[i][0]
        return np.concatenate([obs[:, self.obs_splits[i]], acts[:, self.act_splits[i]]], 1), len(self.obs_splits[i])

    def step(self, obs, actions, next_obs, dones, rewards, noises: Sequence[np.ndarray]) -> None:
        self.n_updates += 1
        B = obs.shape[0]
        nxt = []
        for i in range(self.n):  # :132-144
            nz = np.clip(noises[i], F32(-self.noise_clip), F32(self.noise_clip))
            a, _ = mlp_forward(self.actor_targets[i], next_obs[:, self.obs_splits[i]], True)
            nxt.append(np.clip(a + nz, F32(-1), F32(1)))
        next_actions = np.concatenate(nxt, 1).astype(F32)
        for i in range(self.n):
            xin, _ = self._critic_input(i, next_obs, next_actions)
            q_next = np.minimum.reduce([mlp_forward(q, xin, False)[0] for q in self.critic_targets[i]])  # :147-151
            target = (rewards + (F32(1) - dones) * F32(self.gamma) * q_next).astype(F32)
            x, _ = self._critic_input(i, obs, actions)
            grads, loss = [], 0.0
            for q in self.critics[i]:
                qv, cache = mlp_forward(q, x, False)
                d = qv - target
                loss += float(np.mean(d * d, dtype=F32))
                g, _ = mlp_backward(q, cache, (F32(2) / F32(B)) * d, False)
                grads += g
            self.critic_losses[i].append(loss)
            self.critic_opts[i].step(grads)
            if self.n_updates % self.policy_delay == 0:  # :166
                oi = obs[:, self.obs_splits[i]]
                outs = [mlp_forward(a, oi, True) for a in self.actors]  # every actor on agent i's slice (:169-171)
                joint = np.concatenate([o[0] for o in outs], 1).astype(F32)
                xa, col = self._critic_input(i, obs, joint)
                q1, ccache = mlp_forward(self.critics[i][0], xa, False)
                self.actor_losses[i].append(float(-np.mean(q1, dtype=F32)))
                _, dx = mlp_backward(self.critics[i][0], ccache, np.full_like(q1, F32(-1) / F32(B)), False, need_dx=True)
                width = len(self.act_splits[i])
                ag, _ = mlp_backward(self.actors[i], outs[i][1], dx[:, col:col + width].astype(F32), True)
                self.actor_opts[i].step(ag)  # only agent i's optimiser steps (:177-179)
                for qs, ts in zip(self.critics, self.critic_targets):  # :181-182, all agents
                    for q, t in zip(qs, ts):
                        polyak(q, t, self.tau)
                for a, t in zip(self.actors, self.actor_targets):
                    polyak(a, t, self.tau)
