"""SAC gradient step on the device against the reference's SAC.train (fixture tests/golden/sac_update.npz, recorded from
the unmodified reference on CPU torch) and the NumPy restatement oracle/td3_oracle.py::SACUpdateOracle.
Parity bar (float32 both sides, GEMM / reduction order differs): weights atol 1e-5 vs the fixture after 5 steps, 2e-5 vs the oracle at the
default [256,256] architecture; gradients 2e-5 of the tensor's max; losses / entropy coefficient rel 1e-5."""
import numpy as np
import pytest
import torch

import td3_oracle as T
import td3_util as U

pytestmark = pytest.mark.gpu


def _assert_nets(got, want, atol):
    for name in U.SAC_NETS:
        for k, (a, b) in enumerate(zip(got[name], want[name])):
            np.testing.assert_allclose(a, b, rtol=0, atol=atol, err_msg=f"{name}[{k}]")


def test_five_steps_vs_reference_fixture(pkg, golden):
    g = golden("sac_update.npz")
    gamma, tau, target_entropy, lr, interval = [float(x) for x in g["hyper"]]
    eng = pkg.FusedSACUpdate([64, 48], 64, gamma=gamma, tau=tau, learning_rate=lr, target_entropy=target_entropy,
                             ent_coef_init=float(np.exp(g["init_log_ent_coef"][0])), target_update_interval=int(interval))
    eng.load_nets(U.sac_nets_from(g, "init"))
    for k in range(g["eps_pi"].shape[0]):
        eng.update((g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k]),
                   eps_pi=g["eps_pi"][k], eps_next=g["eps_next"][k])
    _assert_nets(eng.nets(), U.sac_nets_from(g, "final"), 1e-5)
    np.testing.assert_allclose(eng.log_ent_coef.cpu().numpy(), g["final_log_ent_coef"], rtol=0, atol=1e-7)
    critic_loss, actor_loss, ent_loss, ent_coef = eng.pop_losses()
    assert critic_loss == pytest.approx(float(g["critic_loss_mean"]), rel=1e-5) and actor_loss == pytest.approx(float(g["actor_loss_mean"]), rel=1e-5)
    assert ent_coef == pytest.approx(float(g["ent_coef_mean"]), rel=1e-6) and ent_loss == pytest.approx(float(g["ent_coef_loss_mean"]), rel=1e-4)


def _batches(rng, K, B):
    return [(rng.uniform(-1, 1, (B, 4)).astype(np.float32), rng.uniform(-1, 1, (B, 2)).astype(np.float32), rng.uniform(-1, 1, (B, 4)).astype(np.float32),
             (rng.random((B, 1)) < 0.1).astype(np.float32), rng.normal(-1, 1, (B, 1)).astype(np.float32), rng.normal(size=(B, 2)).astype(np.float32),
             rng.normal(size=(B, 2)).astype(np.float32)) for _ in range(K)]


@pytest.mark.parametrize("arch,B,K", [([256, 256], 256, 4), ([36, 20], 100, 3), ([400, 300], 1000, 2), ([128, 64], 7, 3)])
def test_vs_oracle(pkg, arch, B, K):
    rng = np.random.default_rng(B + 3)
    nets = U.random_sac_nets(rng, *arch)
    nets["actor"][5][2:4] += np.float32(-1.0)  # typical log_std level; some rows also hit the clamp below
    o = T.SACUpdateOracle(nets["actor"], [nets["critic0"], nets["critic1"]], log_ent_coef=float(np.log(0.7)))
    eng = pkg.FusedSACUpdate(arch, B, ent_coef_init=0.7)
    eng.load_nets(nets)
    for batch in _batches(rng, K, B):
        out = o.step(*batch)
        eng.update(batch[:5], eps_pi=batch[5], eps_next=batch[6])
        gv = eng.views("grads")
        for name, want_list in (("critic0", out["critic_grads"][:6]), ("critic1", out["critic_grads"][6:]), ("actor", out["actor_grads"])):
            for k, want in enumerate(want_list):
                np.testing.assert_allclose(gv[name][k].cpu().numpy(), want, rtol=0, atol=2e-5 * max(np.abs(want).max(), 1e-3), err_msg=f"{name} grad {k}")
    want = {"actor": o.actor, "critic0": o.critics[0], "critic1": o.critics[1], "critic0_target": o.critic_targets[0], "critic1_target": o.critic_targets[1]}
    _assert_nets(eng.nets(), want, 2e-5)
    np.testing.assert_allclose(eng.log_ent_coef.cpu().numpy(), o.log_ent_coef[0], rtol=0, atol=1e-6)
    critic_loss, actor_loss, ent_loss, ent_coef = eng.pop_losses()
    assert critic_loss == pytest.approx(np.mean(o.critic_losses), rel=1e-5) and actor_loss == pytest.approx(np.mean(o.actor_losses), rel=1e-4, abs=1e-6)
    assert ent_coef == pytest.approx(np.mean(o.ent_coefs), rel=1e-6)


def test_phases_split_equals_one_call(pkg):
    """GRAD / APPLY phases with a no-op all-reduce in between (the data-parallel call pattern) = the single call, bit for bit."""
    rng = np.random.default_rng(12)
    nets = U.random_sac_nets(rng, 256, 256)
    batches = _batches(rng, 4, 300)

    def run(split):
        eng = pkg.FusedSACUpdate([256, 256], 300, ent_coef_init=0.5, seed=4)
        eng.load_nets(nets)
        seen = []
        for b in batches:
            eng.update(b[:5], allreduce=(lambda flat: seen.append(flat.numel())) if split else None)
        return eng.params.clone(), eng.targets.clone(), eng.adam_m.clone(), eng.pop_losses(), seen

    a, b = run(False), run(True)
    assert all(torch.equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] == b[3]
    eng = pkg.FusedSACUpdate([256, 256], 300)
    n_critics_ent = eng._ent_offset + 4 - eng.critic_range[0]
    assert b[4] == [n_critics_ent, eng.actor_range[1] - eng.actor_range[0]] * 4  # critics + log_ent_coef slot, then the actor


def test_log_std_clamp_and_philox_noise(pkg):
    """Rows whose raw log_std is outside [-20, 2] take the clamped value and pass no gradient to the log_std head; the Philox
    draws are standard normal (checked through log_prob statistics) and reproducible."""
    rng = np.random.default_rng(0)
    nets = U.random_sac_nets(rng, 64, 48)
    nets["actor"][5][2] = np.float32(-25.0)  # log_std bias far below the clamp for action 0 -> every row clamped at -20 (the high side
    # would saturate tanh: 1 - a^2 ~ 1e-7 next to the 1e-6 epsilon makes the squash term depend on the last ulp of tanh)
    o = T.SACUpdateOracle(nets["actor"], [nets["critic0"], nets["critic1"]])
    eng = pkg.FusedSACUpdate([64, 48], 128)
    eng.load_nets(nets)
    batch = _batches(rng, 1, 128)[0]
    out = o.step(*batch)
    eng.update(batch[:5], eps_pi=batch[5], eps_next=batch[6])
    g = eng.views("grads")["actor"]
    assert float(g[5][2].abs().item()) == 0.0 and float(g[4][2].abs().max().item()) == 0.0  # clamped head row: zero gradient
    np.testing.assert_allclose(g[4].cpu().numpy(), out["actor_grads"][4], rtol=0, atol=2e-5 * np.abs(out["actor_grads"][4]).max())
    a = pkg.FusedSACUpdate([64, 48], 4096, seed=3)
    b = pkg.FusedSACUpdate([64, 48], 4096, seed=3)
    c = pkg.FusedSACUpdate([64, 48], 4096, seed=4)
    big = _batches(rng, 1, 4096)[0]
    nets2 = U.random_sac_nets(rng, 64, 48)
    for e in (a, b, c):
        e.load_nets(nets2)
        e.update(big[:5])
    assert torch.equal(a.params, b.params) and not torch.equal(a.params, c.params)


def test_tensor_gemm_mode_tracks_fp32(pkg):
    """gemm="tensor" (tcgen05, 3-plane bf16 split) for the SAC step: one update from identical weights — losses agree to 1e-4, the
    gradients to 1e-4 of their norm (ReLU flips at |activation| < 2e-5 excepted, see tests/test_gpu_td3.py), the weights within Adam's
    first-step bound of 2*lr."""
    rng = np.random.default_rng(11)
    nets = U.random_sac_nets(rng, 256, 256)
    batch = _batches(rng, 1, 512)[0]
    engs = {}
    for gemm in ("fp32", "tensor"):
        e = pkg.FusedSACUpdate([256, 256], 512, gemm=gemm)
        e.load_nets(nets)
        e.update(batch[:5], eps_pi=batch[5], eps_next=batch[6])
        engs[gemm] = e
    a, b = engs["fp32"], engs["tensor"]
    la, lb = a.pop_losses(), b.pop_losses()
    for x, y in zip(la, lb):
        assert y == pytest.approx(x, rel=1e-4, abs=1e-6)
    ga, gb = a.grads.cpu().numpy(), b.grads.cpu().numpy()
    assert np.linalg.norm(ga - gb) <= 2e-3 * np.linalg.norm(ga)
    assert float((a.params - b.params).abs().max().item()) <= 2.1 * 3e-4


def test_cuda_graph_replay_equals_launch_by_launch(pkg):
    rng = np.random.default_rng(8)
    nets = U.random_sac_nets(rng, 256, 256)
    n_envs = 512
    buf = pkg.GpuReplayBuffer(32 * n_envs, n_envs=n_envs, index_mode="philox", seed=5)
    buf.records.uniform_(-1, 1)
    buf.records[..., 11:13] = 0
    buf.pos, buf.full = 0, True

    def run(graph):
        buf._draw = 0
        eng = pkg.FusedSACUpdate([256, 256], 256, seed=2)
        eng.load_nets(nets)
        for steps in (4, 3):
            eng.train(steps, buf, 256, graph=graph)
        return eng

    a, b = run(False), run(True)
    assert a.n_updates == b.n_updates == 7 and b._graph is not None and buf._draw == 7
    np.testing.assert_allclose(b.params.cpu().numpy(), a.params.cpu().numpy(), rtol=0, atol=3e-7)
    np.testing.assert_allclose(b.targets.cpu().numpy(), a.targets.cpu().numpy(), rtol=0, atol=3e-7)
    for x, y in zip(a.pop_losses(), b.pop_losses()):
        assert y == pytest.approx(x, rel=1e-5)


def _torch_sac_twin(seed, H=64, lr=3e-4):
    """A stand-in for the reference's SACPolicy (same attribute names) with torch.optim.Adam optimisers, and one gradient step of
    sac.py:213-288 written with torch autograd — an independent check of the fused step and of the optimiser-state hand-over."""
    import torch.nn as nn
    from types import SimpleNamespace as NS

    torch.manual_seed(seed)

    def mlp(i, o=None):
        return nn.Sequential(*([nn.Linear(i, H), nn.ReLU(), nn.Linear(H, H), nn.ReLU()] + ([nn.Linear(H, o)] if o else []))).cuda()

    actor = NS(latent_pi=mlp(4), mu=nn.Linear(H, 2).cuda(), log_std=nn.Linear(H, 2).cuda())
    critic, critic_t = NS(q_networks=[mlp(6, 1), mlp(6, 1)]), NS(q_networks=[mlp(6, 1), mlp(6, 1)])
    for a, b in zip(critic.q_networks, critic_t.q_networks):
        b.load_state_dict(a.state_dict())
    tw = NS(policy=NS(actor=actor, critic=critic, critic_target=critic_t), H=H)
    tw.actor_params = list(actor.latent_pi.parameters()) + list(actor.mu.parameters()) + list(actor.log_std.parameters())
    tw.critic_params = [p for q in critic.q_networks for p in q.parameters()]
    tw.log_ent = torch.log(torch.ones(1, device="cuda") * 0.8).requires_grad_(True)
    tw.opt_a, tw.opt_c, tw.opt_e = (torch.optim.Adam(ps, lr=lr) for ps in (tw.actor_params, tw.critic_params, [tw.log_ent]))

    def pi(obs, eps):
        lat = actor.latent_pi(obs)
        mu, std = actor.mu(lat), actor.log_std(lat).clamp(-20, 2).exp()
        g = mu + std * eps
        a = torch.tanh(g)
        return a, torch.distributions.Normal(mu, std).log_prob(g).sum(1) - torch.log(1 - a ** 2 + 1e-6).sum(1)

    def step(batch, gamma=0.99, tau=0.005, target_entropy=-2.0):
        obs, act, nobs, dones, rew, e1, e2 = (torch.as_tensor(x, device="cuda") for x in batch)
        a_pi, logp = pi(obs, e1)
        ent_coef = tw.log_ent.detach().exp()
        ent_loss = -(tw.log_ent * (logp.reshape(-1, 1) + target_entropy).detach()).mean()
        tw.opt_e.zero_grad(); ent_loss.backward(); tw.opt_e.step()
        with torch.no_grad():
            na, nlogp = pi(nobs, e2)
            nq = torch.cat([q(torch.cat([nobs, na], 1)) for q in critic_t.q_networks], 1).min(1, keepdim=True)[0] - ent_coef * nlogp.reshape(-1, 1)
            target = rew + (1 - dones) * gamma * nq
        closs = 0.5 * sum(torch.nn.functional.mse_loss(q(torch.cat([obs, act], 1)), target) for q in critic.q_networks)
        tw.opt_c.zero_grad(); closs.backward(); tw.opt_c.step()
        qpi = torch.cat([q(torch.cat([obs, a_pi], 1)) for q in critic.q_networks], 1).min(1, keepdim=True)[0]
        aloss = (ent_coef * logp.reshape(-1, 1) - qpi).mean()
        tw.opt_a.zero_grad(); aloss.backward(); tw.opt_a.step()
        with torch.no_grad():
            for q, qt in zip(critic.q_networks, critic_t.q_networks):
                for p, t in zip(q.parameters(), qt.parameters()):
                    t.mul_(1 - tau).add_(p, alpha=tau)

    tw.step = step
    return tw


def test_adopt_policy_and_optimizer_state_hand_over(pkg):
    """Two torch steps, then the engine adopts the modules AND the three Adam states and does two more: same weights as four torch
    steps; exporting writes the moments and the step count back into the torch optimisers."""
    rng = np.random.default_rng(21)
    batches = _batches(rng, 4, 192)
    a, b = _torch_sac_twin(5), _torch_sac_twin(5)
    for batch in batches:
        a.step(batch)
    for batch in batches[:2]:
        b.step(batch)
    eng = pkg.FusedSACUpdate([b.H, b.H], 192, ent_coef_init=1.0)
    eng.adopt_policy(b.policy)
    eng.log_ent_coef.copy_(b.log_ent.detach())
    b.log_ent.data = eng.log_ent_coef
    eng.import_optimizer_state(b.opt_a, b.opt_c, b.opt_e)
    assert eng.critic_step == 2 and torch.equal(eng.views("adam_m")["actor"][4][2:4], b.opt_a.state[b.policy.actor.log_std.weight]["exp_avg"])
    eng.n_updates = 2
    for batch in batches[2:]:
        eng.update(batch[:5], eps_pi=batch[5], eps_next=batch[6])
    for pa, pb in zip(a.actor_params + a.critic_params + [a.log_ent], b.actor_params + b.critic_params + [b.log_ent]):
        np.testing.assert_allclose(pb.detach().cpu().numpy(), pa.detach().cpu().numpy(), rtol=0, atol=2e-5)  # b's modules are views of the engine's block
    for qa, qb in zip(a.policy.critic_target.q_networks, b.policy.critic_target.q_networks):
        for pa, pb in zip(qa.parameters(), qb.parameters()):
            np.testing.assert_allclose(pb.detach().cpu().numpy(), pa.detach().cpu().numpy(), rtol=0, atol=2e-5)
    eng.export_optimizer_state(b.opt_a, b.opt_c, b.opt_e)
    for oa, ob in ((a.opt_a, b.opt_a), (a.opt_c, b.opt_c), (a.opt_e, b.opt_e)):
        for pa, pb in zip(oa.param_groups[0]["params"], ob.param_groups[0]["params"]):
            sa, sb = oa.state[pa], ob.state[pb]
            assert float(sb["step"]) == float(sa["step"]) == 4
            scale = max(float(sa["exp_avg"].abs().max()), 1e-6)
            np.testing.assert_allclose(sb["exp_avg"].cpu().numpy(), sa["exp_avg"].cpu().numpy(), rtol=0, atol=2e-4 * scale)
    b.step(batches[0])  # and torch can carry on from the exported state (moments are views of the engine's blocks)
    assert float(b.opt_c.state[b.critic_params[0]]["step"]) == 5


def test_target_sync_follows_the_loop_index_of_train(pkg):
    """sac.py:284 tests `gradient_step % target_update_interval` with the LOOP index, which restarts at 0 in every train() call: with
    interval 2 and one gradient step per call (the usual train_freq=1, gradient_steps=1) the reference polyaks on EVERY call."""
    rng = np.random.default_rng(3)
    nets = U.random_sac_nets(rng, 64, 48)
    n_envs = 256
    buf = pkg.GpuReplayBuffer(16 * n_envs, n_envs=n_envs, index_mode="philox", seed=5)
    buf.records.uniform_(-1, 1)
    buf.records[..., 11:13] = 0
    buf.pos, buf.full = 0, True
    eng = pkg.FusedSACUpdate([64, 48], 128, seed=2, target_update_interval=2)
    eng.load_nets(nets)
    orc = T.SACUpdateOracle(nets["actor"], [nets["critic0"], nets["critic1"]], [nets["critic0_target"], nets["critic1_target"]],
                            target_update_interval=2)
    for call in range(3):  # three train() calls of (1, then 3, then 1) gradient steps
        steps = (1, 3, 1)[call]
        orc.gradient_step = 0  # the reference's loop variable
        for g in range(steps):
            b = buf.sample(128)
            e1, e2 = rng.normal(size=(128, 2)).astype(np.float32), rng.normal(size=(128, 2)).astype(np.float32)
            before = eng.targets.clone()
            eng.update(b, eps_pi=e1, eps_next=e2, gradient_step=g)
            orc.step(*[t.cpu().numpy() for t in b], e1, e2)
            assert (not torch.equal(before, eng.targets)) == (g % 2 == 0)
    got = eng.nets()
    for name, want in (("critic0_target", orc.critic_targets[0]), ("critic1_target", orc.critic_targets[1]), ("actor", orc.actor)):
        for a, w in zip(got[name], want):
            np.testing.assert_allclose(a, w, rtol=0, atol=2e-5)
