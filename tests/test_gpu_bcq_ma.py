"""-m gpu: the BCQ and MADDPG / IDDPG gradient steps on the device (cstr_bcq_update, cstr_ma_update) against the reference's
own BCQ.train / MADDPG.train / IDDPG.train (fixtures tests/golden/{bcq,maddpg,iddpg}_update.npz, recorded from the unmodified reference on CPU
torch) and against the NumPy restatements oracle/td3_oracle.py::BCQUpdateOracle / MultiAgentDDPGOracle at the default widths.
Parity bar: float32 on both sides, GEMM / batch-reduction order differs — weights atol 1e-5 vs the fixtures, 2e-5 vs the oracle at default
widths (as TD3), gradients 2e-5 of the tensor's max, losses rel 2e-5."""
import numpy as np
import pytest
import torch

import td3_oracle as T
import td3_util as U

pytestmark = pytest.mark.gpu


def _assert_nets(got, want, names, atol):
    for name in names:
        for k, (a, b) in enumerate(zip(got[name], want[name])):
            np.testing.assert_allclose(a, b, rtol=0, atol=atol, err_msg=f"{name}[{k}]")


# ---- BCQ -----------------------------------------------------------------------------------------------------------------------------
def _bcq_engine(pkg, nets, latent, hv, hp, critic_arch, batch, **kw):
    eng = pkg.FusedBCQUpdate(latent, hv, hp, critic_arch, batch, **kw)
    eng.load_nets(nets)
    return eng


def test_bcq_five_steps_vs_reference_fixture(pkg, golden):
    g = golden("bcq_update.npz")
    gamma, tau, phi, lr, delay = [float(x) for x in g["hyper"]]
    init = U.bcq_nets_from(g, "init")
    L, hv, hp = init["vae_enc"][4].shape[0] // 2, init["vae_enc"][0].shape[0], init["pert"][0].shape[0]
    arch = [init["critic0"][0].shape[0], init["critic0"][2].shape[0]]
    B = g["batch_obs"].shape[1]
    eng = _bcq_engine(pkg, init, L, hv, hp, arch, B, gamma=gamma, tau=tau, learning_rate=lr, max_perturbation=phi, actor_delay=int(delay))
    for k in range(g["eps_vae"].shape[0]):
        eng.update((g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k]),
                   eps_vae=g["eps_vae"][k], z_next=g["z_next"][k], z_actor=g["z_actor"][k])
    _assert_nets(eng.nets(), U.bcq_nets_from(g, "final"), U.BCQ_NETS, 1e-5)
    vae_loss, critic_loss, actor_loss = eng.pop_losses()
    assert vae_loss == pytest.approx(float(g["vae_loss_mean"]), rel=2e-5)
    assert critic_loss == pytest.approx(float(g["critic_loss_mean"]), rel=2e-5)
    assert actor_loss == pytest.approx(float(g["actor_loss_mean"]), rel=1e-4, abs=1e-6)
    assert eng.critic_step == 5 and eng.actor_step == 2


def _random_mlp(rng, i, h1, h2, o):
    out = []
    for fi, fo in ((i, h1), (h1, h2), (h2, o)):
        b = 1.0 / np.sqrt(fi)
        out += [rng.uniform(-b, b, (fo, fi)).astype(np.float32), rng.uniform(-b, b, fo).astype(np.float32)]
    return out


def _bcq_random(rng, L, hv, hp, h1, h2):
    return {"vae_enc": _random_mlp(rng, 6, hv, hv, 2 * L), "vae_dec": _random_mlp(rng, 4 + L, hv, hv, 2), "pert": _random_mlp(rng, 6, hp, hp, 2),
            "critic0": _random_mlp(rng, 6, h1, h2, 1), "critic1": _random_mlp(rng, 6, h1, h2, 1)}


@pytest.mark.parametrize("L,hv,hp,arch,B,K", [(32, 64, 64, [400, 300], 256, 4),   # BCQPolicy's defaults (policies.py:320-322)
                                             (12, 700, 400, [400, 300], 100, 2),  # the experiment script's sizes (HalfCheetah_BCQ.py:55-58)
                                             (8, 36, 20, [36, 20], 37, 3)])
def test_bcq_vs_oracle(pkg, L, hv, hp, arch, B, K):
    rng = np.random.default_rng(L + B)
    nets = _bcq_random(rng, L, hv, hp, *arch)
    o = T.BCQUpdateOracle(nets["vae_enc"], nets["vae_dec"], nets["pert"], [nets["critic0"], nets["critic1"]])
    eng = _bcq_engine(pkg, nets, L, hv, hp, arch, B)
    for k in range(K):
        batch = (rng.uniform(-1, 1, (B, 4)).astype(np.float32), rng.uniform(-1, 1, (B, 2)).astype(np.float32), rng.uniform(-1, 1, (B, 4)).astype(np.float32),
                 (rng.uniform(size=(B, 1)) < 0.1).astype(np.float32), rng.normal(size=(B, 1)).astype(np.float32))
        eps, zn, za = (rng.normal(size=(B, L)).astype(np.float32), rng.normal(size=(10 * B, L)).astype(np.float32), rng.normal(size=(B, L)).astype(np.float32))
        out = o.step(*batch, eps, zn, za)
        eng.update(batch, eps_vae=eps, z_next=zn, z_actor=za)
        if k == 0:  # gradients of the first step (identical weights on both sides)
            gv = eng.views("grads")
            for name, want in (("vae_enc", out["vae_grads"][:6]), ("vae_dec", out["vae_grads"][6:]), ("critic0", out["critic_grads"][:6]),
                               ("critic1", out["critic_grads"][6:])):
                for j, w in enumerate(want):
                    np.testing.assert_allclose(gv[name][j].cpu().numpy(), w, rtol=0, atol=2e-5 * max(np.abs(w).max(), 1e-3), err_msg=f"{name} grad {j}")
    want = {"vae_enc": o.vae_enc, "vae_dec": o.vae_dec, "pert": o.pert, "critic0": o.critics[0], "critic1": o.critics[1], "vae_enc_target": o.vae_enc,
            "vae_dec_target": o.vae_dec, "pert_target": o.pert_target, "critic0_target": o.critic_targets[0], "critic1_target": o.critic_targets[1]}
    # Adam's early steps are lr * g/|g|-like: an element whose gradient is at rounding level can move by up to lr either way, so the bar
    # on weights after K steps is on the bulk (as in the TD3 tensor-core test), with a hard cap of K * lr
    got = eng.nets()
    for name in U.BCQ_NETS:
        for a, b in zip(got[name], want[name]):
            d = np.abs(a - b)
            assert d.max() <= K * 1.1e-3 and (d > 2e-5).mean() < 2e-3, (name, d.max(), (d > 2e-5).mean())
    vae_loss, critic_loss, actor_loss = eng.pop_losses()
    assert vae_loss == pytest.approx(np.mean(o.vae_losses), rel=5e-5)
    assert critic_loss == pytest.approx(np.mean(o.critic_losses), rel=5e-5)
    if o.actor_losses:
        assert actor_loss == pytest.approx(np.mean(o.actor_losses), rel=2e-4, abs=1e-6)


def test_bcq_philox_draws_graph_replay_and_determinism(pkg):
    """Default draws come from Philox inside the kernels (deterministic for a seed); train(graph=True) replays cycles of actor_delay updates
    from one CUDA graph and lands on the same weights as launch by launch."""
    rng = np.random.default_rng(3)
    nets = _bcq_random(rng, 32, 64, 64, 400, 300)
    n_envs = 512
    buf = pkg.GpuReplayBuffer(32 * n_envs, n_envs=n_envs, index_mode="philox", seed=5)
    buf.records.uniform_(-1, 1)
    buf.records[..., 11:13] = 0
    buf.pos, buf.full = 0, True

    def run(graph, seed=9):
        buf._draw = 0
        eng = _bcq_engine(pkg, nets, 32, 64, 64, [400, 300], 256, seed=seed)
        for steps in (6, 5):
            eng.train(steps, buf, 256, graph=graph)
        return eng

    a, b, c, d = run(False), run(True), run(False), run(False, seed=10)
    assert a.n_updates == b.n_updates == 11 and a.actor_step == b.actor_step == 5 and b._graph is not None
    assert torch.equal(a.params, c.params) and not torch.equal(a.params, d.params)
    np.testing.assert_allclose(b.params.cpu().numpy(), a.params.cpu().numpy(), rtol=0, atol=3e-7)
    np.testing.assert_allclose(b.targets.cpu().numpy(), a.targets.cpu().numpy(), rtol=0, atol=3e-7)
    for x, y in zip(a.pop_losses(), b.pop_losses()):
        assert y == pytest.approx(x, rel=1e-4)
    assert bool(torch.isfinite(a.params).all())


# ---- MADDPG / IDDPG ---------------------------------------------------------------------------------------------------------------------
def _ma_engine(pkg, nets, arch, batch, centralised, **kw):
    eng = pkg.FusedMultiAgentUpdate(arch, batch, centralised, **kw)
    eng.load_nets(nets)
    return eng


@pytest.mark.parametrize("algo,centralised", [("maddpg", True), ("iddpg", False)])
def test_multi_agent_four_steps_vs_reference_fixture(pkg, golden, algo, centralised):
    g = golden(f"{algo}_update.npz")
    gamma, tau, delay, clip, lr0, lr1 = [float(x) for x in g["hyper"]]
    init = U.ma_nets_from(g, "init")
    arch = [init["actor0"][0].shape[0], init["actor0"][2].shape[0]]
    B = g["batch_obs"].shape[1]
    actor_lrs, critic_lrs = pkg.FusedMultiAgentUpdate.reference_lrs([lr0, lr1])
    eng = _ma_engine(pkg, init, arch, B, centralised, gamma=gamma, tau=tau, policy_delay=int(delay), target_noise_clip=clip, actor_lrs=actor_lrs,
                     critic_lrs=critic_lrs)
    for k in range(g["noise"].shape[0]):
        eng.update((g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k]), noise=list(g["noise"][k]))
    _assert_nets(eng.nets(), U.ma_nets_from(g, "final"), U.MA_NETS, 1e-5)
    losses = eng.pop_losses()
    for i in range(2):
        assert losses[i][0] == pytest.approx(float(g[f"critic_loss_mean_{i}"]), rel=2e-5)
        assert losses[i][1] == pytest.approx(float(g[f"actor_loss_mean_{i}"]), rel=1e-4, abs=1e-6)
    # the naive per-agent reading of learning_rate_list misses the reference (the fixture uses unequal rates)
    naive = _ma_engine(pkg, init, arch, B, centralised, gamma=gamma, tau=tau, policy_delay=int(delay), target_noise_clip=clip, actor_lrs=[lr0, lr1],
                       critic_lrs=[lr0, lr1])
    for k in range(g["noise"].shape[0]):
        naive.update((g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k]), noise=list(g["noise"][k]))
    assert np.abs(naive.nets()["critic0_0"][0] - U.ma_nets_from(g, "final")["critic0_0"][0]).max() > 1e-4


def _ma_random(rng, h1, h2, centralised):
    ci = 6 if centralised else 3
    nets = {f"actor{i}": _random_mlp(rng, 2, h1, h2, 1) for i in range(2)}
    nets.update({f"critic{i}_{k}": _random_mlp(rng, ci, h1, h2, 1) for i in range(2) for k in range(2)})
    return nets


@pytest.mark.parametrize("centralised", [True, False])
@pytest.mark.parametrize("arch,B,K", [([400, 300], 256, 4), ([36, 20], 37, 3)])
def test_multi_agent_vs_oracle(pkg, centralised, arch, B, K):
    rng = np.random.default_rng(B + int(centralised))
    nets = _ma_random(rng, *arch, centralised)
    lrs = ([1e-3, 1e-3], [5e-4, 5e-4])
    o = T.MultiAgentDDPGOracle([nets["actor0"], nets["actor1"]], [[nets["critic0_0"], nets["critic0_1"]], [nets["critic1_0"], nets["critic1_1"]]],
                               [[0, 1], [2, 3]], [[0], [1]], centralised, *lrs)
    eng = _ma_engine(pkg, nets, arch, B, centralised, actor_lrs=lrs[0], critic_lrs=lrs[1])
    for _ in range(K):
        batch = (rng.uniform(-1, 1, (B, 4)).astype(np.float32), rng.uniform(-1, 1, (B, 2)).astype(np.float32), rng.uniform(-1, 1, (B, 4)).astype(np.float32),
                 (rng.uniform(size=(B, 1)) < 0.1).astype(np.float32), rng.normal(size=(B, 1)).astype(np.float32))
        nz = [rng.normal(0, 0.2, (B, 1)).astype(np.float32) for _ in range(2)]
        o.step(*batch, nz)
        eng.update(batch, noise=nz)
    got = eng.nets()
    for i in range(2):
        pairs = [(f"actor{i}", o.actors[i]), (f"actor{i}_target", o.actor_targets[i])]
        pairs += [(f"critic{i}_{k}", o.critics[i][k]) for k in range(2)] + [(f"critic{i}_{k}_target", o.critic_targets[i][k]) for k in range(2)]
        for name, want in pairs:
            for a, b in zip(got[name], want):
                d = np.abs(a - b)
                assert d.max() <= K * 1.1e-3 and (d > 2e-5).mean() < 2e-3, (name, d.max(), (d > 2e-5).mean())
    losses = eng.pop_losses()
    for i in range(2):
        assert losses[i][0] == pytest.approx(np.mean(o.critic_losses[i]), rel=5e-5)
        assert losses[i][1] == pytest.approx(np.mean(o.actor_losses[i]), rel=2e-4, abs=1e-6)


@pytest.mark.parametrize("centralised", [True, False])
def test_multi_agent_philox_noise_and_graph_replay(pkg, centralised):
    rng = np.random.default_rng(11)
    nets = _ma_random(rng, 400, 300, centralised)
    n_envs = 512
    buf = pkg.GpuReplayBuffer(32 * n_envs, n_envs=n_envs, index_mode="philox", seed=5)
    buf.records.uniform_(-1, 1)
    buf.records[..., 11:13] = 0
    buf.pos, buf.full = 0, True

    def run(graph):
        buf._draw = 0
        eng = _ma_engine(pkg, nets, [400, 300], 256, centralised, seed=4, actor_lrs=[1e-3, 1e-3], critic_lrs=[5e-4, 5e-4])
        for steps in (6, 5):
            eng.train(steps, buf, 256, graph=graph)
        return eng

    a, b, c = run(False), run(True), run(False)
    assert a.n_updates == b.n_updates == 11 and a.actor_step == b.actor_step == 5 and b._graph is not None
    assert torch.equal(a.params, c.params)
    np.testing.assert_allclose(b.params.cpu().numpy(), a.params.cpu().numpy(), rtol=0, atol=5e-7)
    np.testing.assert_allclose(b.targets.cpu().numpy(), a.targets.cpu().numpy(), rtol=0, atol=5e-7)


def test_config_errors(pkg):
    with pytest.raises(ValueError):
        pkg.FusedBCQUpdate(latent_dim=6)  # not a multiple of 4
    with pytest.raises(ValueError):
        pkg.FusedMultiAgentUpdate([400, 300], n_critics=3)
    with pytest.raises(pkg.CstrLibraryError):
        pkg.FusedBCQUpdate(device="cpu")
