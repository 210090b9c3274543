"""CPU restatement of the reference's TD3 gradient step (TEST INFRASTRUCTURE — only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product path never does).

Follows, line by line:
  TD3.train                      core/td3/td3.py:154-211   (target smoothing, twin-min target, MSE sum, delayed actor, polyak)
  create_mlp / Actor / critic    core/common/torch_layers.py:110-183, core/td3/policies.py:58,75-78, core/common/policies.py:966-987
  polyak_update                  core/common/utils.py:457-481
  torch.optim.Adam (defaults)    betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad  (torch/optim/adam.py, single-tensor path)
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
        q_next = np.minimum(*[mlp_forward(c, xin, False)[0] for c in self.critic_targets])  # :173-174
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
