"""-m gpu: fused rollout kernel (actor -> noise -> bounds -> step -> record) against the oracle."""
import numpy as np
import pytest
import torch

import build_oracle as B
import cstr_oracle as O

pytestmark = pytest.mark.gpu


def _actor(golden):
    g = golden("td3_actor.npz")
    return g, [(g["W1"], g["b1"]), (g["W2"], g["b2"]), (g["W3"], g["b3"])]


def _env_action_from_buffer_action(b):
    f = np.float32
    return (f(-1.0) + (f(0.5) * (b + f(1.0)) * f(2.0))).astype(f)


# fp32 path: fp32 accumulation order differs from torch/NumPy -> 3e-6.  tc path: h1 and W2 are rounded to bf16
# (8-bit mantissa) before the tcgen05 MMA, fp32 accumulate -> |Δaction| <= 5e-3 (measured 7e-4), the env step
# on the action it stored stays bit-exact.
@pytest.mark.parametrize("actor_mode,atol", [("fp32", 3e-6), ("tc", 5e-3)])
def test_rollout_matches_oracle_composition(pkg, golden, actor_mode, atol):
    g, w = _actor(golden)
    n, K, seed = 256, 3, 11
    env = pkg.GpuCSTRVecEnv(n, seed=seed, monitor=False)
    env.reset()
    env.set_state(g["obs"], np.zeros(n, np.int32))  # the fixture's 256 observations
    buf = pkg.GpuReplayBuffer(8 * n, device="cuda", n_envs=n)
    actor = pkg.ActorWeights(g["W1"], g["b1"], g["W2"], g["b2"], g["W3"], g["b3"])
    roll = pkg.FusedRollout(env, buf, actor, sigma=0.1, actor_mode=actor_mode)
    rng = np.random.default_rng(0)
    noise = (0.1 * rng.standard_normal((K, n, 2))).astype(np.float32)
    noise[0, :4] = g["noise"][:4]  # saturating rows
    noise[0, 4:] = g["noise"][4:]
    roll.collect(K, noise=torch.as_tensor(noise, device="cuda"))
    rec = buf.records.cpu().numpy()
    assert buf.pos == K and not buf.full
    state = g["obs"].copy()
    sc = np.zeros(n, np.int32)
    for k in range(K):
        # (1) the record's obs is the state the actor saw
        assert np.array_equal(rec[k, :, 0:4], state)
        # (2) actor + action maps: reference torch forward (fixture) for k=0, fp64 oracle otherwise
        mu = g["mu"] if k == 0 else O.actor_forward(state, w).astype(np.float32)
        a_env, a_buf = O.sample_action_maps(mu, noise[k])
        np.testing.assert_allclose(rec[k, :, 8:10], a_buf, rtol=0, atol=atol)
        if actor_mode == "tc":
            print("tc actor |Δaction| max", float(np.abs(rec[k, :, 8:10] - a_buf).max()))
        if k == 0:
            np.testing.assert_allclose(rec[k, :, 8:10], g["buffer_action"], rtol=0, atol=atol)
        assert np.abs(rec[k, :, 8:10]).max() <= 1.0
        # (3) given the action it stored, the env step is the strict kernel bit for bit
        s, r, tr, sc, _ = B.step_f32(state, _env_action_from_buffer_action(rec[k, :, 8:10]), sc, exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)
        assert np.array_equal(rec[k, :, 4:8], s) and np.array_equal(rec[k, :, 10], r)
        assert np.array_equal(rec[k, :, 11], tr.astype(np.float32)) and np.array_equal(rec[k, :, 12], tr.astype(np.float32))
        state = s
    assert np.array_equal(env.state.cpu().numpy(), state) and np.array_equal(env.step_count.cpu().numpy(), sc)


def test_rollout_warmup_autoreset_and_ring(pkg):
    n, seed = 640, 5  # 5 CTAs of 128
    env = pkg.GpuCSTRVecEnv(n, seed=seed, monitor=False)
    env.reset()
    env.step_count.fill_(398)
    buf = pkg.GpuReplayBuffer(4 * n, device="cuda", n_envs=n)  # 4-row ring
    roll = pkg.FusedRollout(env, buf, None, sigma=0.1)
    rs = torch.zeros(1, dtype=torch.float64, device="cuda")
    st0 = env.state.cpu().numpy().copy()
    roll.collect(3, warmup=True, reward_sum=rs)
    rec = buf.records.cpu().numpy()
    assert buf.pos == 3 and roll.t == 3
    assert np.array_equal(rec[0, :, 0:4], st0)
    assert np.abs(rec[:3, :, 8:10]).max() <= 1.0 and abs(rec[:3, :, 8:10].mean()) < 0.05  # U(-1,1) warm-up actions
    assert (rec[1, :, 11] == 1).all() and (rec[0, :, 11] == 0).all() and (rec[1, :, 12] == 1).all()  # step 400: done+timeout
    # row 2 starts from the post-reset state (episode 1 of the Philox reset stream), not from the terminal obs
    st_reset, _, _, _ = B.reset_f32(n, 0, seed, 0, episode=np.ones(n, np.int32))
    assert np.array_equal(rec[2, :, 0:4], st_reset) and not np.array_equal(rec[2, :, 0:4], rec[1, :, 4:8])
    assert abs(rs.item() - rec[:3, :, 10].astype(np.float64).sum()) < 1e-6 * abs(rs.item())
    roll.collect(2, warmup=True)  # wraps
    assert buf.pos == 1 and buf.full
    rec2 = buf.records.cpu().numpy()
    assert np.array_equal(rec2[3, :, 0:4], rec[2, :, 4:8]) and np.array_equal(rec2[0, :, 0:4], rec2[3, :, 4:8])
    s = buf.sample(128)
    assert s.observations.shape == (128, 4)


def test_rollout_philox_noise_statistics(pkg, golden):
    g, w = _actor(golden)
    n = 128 * 64
    env = pkg.GpuCSTRVecEnv(n, seed=1, monitor=False)
    env.reset()
    st = env.state.cpu().numpy().copy()
    buf = pkg.GpuReplayBuffer(2 * n, device="cuda", n_envs=n)
    actor = pkg.ActorWeights(g["W1"], g["b1"], g["W2"], g["b2"], g["W3"], g["b3"])
    pkg.FusedRollout(env, buf, actor, sigma=0.1).collect(1)
    a = buf.records[0, :, 8:10].cpu().numpy()
    mu = O.actor_forward(st, w)
    _, base = O.sample_action_maps(mu.astype(np.float32), np.zeros_like(a))
    d = (a - base)[np.abs(a) < 0.999]
    assert abs(d.mean()) < 5e-3 and abs(d.std() - 0.1) < 5e-3  # N(0, 0.1^2) exploration noise


def test_tc_rollout_multi_tile_many_steps(pkg, golden):
    """tcgen05 path over several CTAs, a ragged last tile and enough steps to wrap every mbarrier phase many times;
    compared with the fp32-actor kernel run on identical inputs (same Philox noise stream)."""
    g, w = _actor(golden)
    n, K = 128 * 5 + 37, 23
    outs = {}
    for mode in ("fp32", "tc"):
        env = pkg.GpuCSTRVecEnv(n, seed=4, monitor=False)
        env.reset()
        env.step_count.fill_(390)  # crosses the truncation row inside the launch
        buf = pkg.GpuReplayBuffer(32 * n, device="cuda", n_envs=n)
        actor = pkg.ActorWeights(g["W1"], g["b1"], g["W2"], g["b2"], g["W3"], g["b3"])
        rs = torch.zeros(1, dtype=torch.float64, device="cuda")
        pkg.FusedRollout(env, buf, actor, sigma=0.1, actor_mode=mode).collect(K, reward_sum=rs)
        torch.cuda.synchronize()
        outs[mode] = (buf.records.cpu().numpy()[:K], env.state.cpu().numpy(), env.step_count.cpu().numpy(), env.episode.cpu().numpy(), rs.item())
    a, b = outs["fp32"], outs["tc"]
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])  # counters / episodes identical
    assert np.array_equal(a[0][:, :, 11], b[0][:, :, 11])  # dones
    assert (a[0][9, :, 11] == 1).all()  # 390 + 10 = 400
    # first step: same observation, actions within the bf16 tolerance
    assert np.array_equal(a[0][0, :, 0:4], b[0][0, :, 0:4])
    assert np.abs(a[0][0, :, 8:10] - b[0][0, :, 8:10]).max() < 2e-2
    # trajectories stay close (contractive dynamics) and the step after the reset starts from identical states
    assert np.array_equal(a[0][10, :, 0:4], b[0][10, :, 0:4])
    assert np.abs(a[0][:, :, 0:4] - b[0][:, :, 0:4]).max() < 5e-2
    assert abs(a[4] - b[4]) < 1e-2 * abs(a[4])


def test_td3_example_learns_end_to_end():
    """examples/td3_fused_rollout.py: fused tcgen05 rollout -> GPU replay (Philox sample) -> fused TD3 update (CUDA graph) -> weights back
    to the rollout kernel, nothing on the host.  Rewards depend on the episode phase, so the example reports whole 400-step episodes: the
    episode return must improve from the random-policy level (about -290) by more than 120 within 15 episodes (measured: -281 -> -34;
    the best constant action reaches -75)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "td3_fused_rollout.py"), "--n-envs", "16384", "--iters", "1500"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["updates"] == 1500 * 16 and res["transitions"] == 1500 * 4 * 16384 and res["update"] == "fused"
    assert 400 * res["mean_reward_last"] > 400 * res["mean_reward_first"] + 120, res


def _run_example(name, *args, timeout=900):
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", name), *args], capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1]), out.stdout


@pytest.mark.parametrize("flag,algo", [((), "MADDPG"), (("--iddpg",), "IDDPG")])
def test_multi_agent_example_runs_entirely_on_the_device(flag, algo):
    """examples/maddpg_two_agents.py: cstr_rollout_fused_multi (two per-agent actors + step + record, 4 steps per launch) and cstr_ma_update
    (graph-replayed once the ring is full): one rollout launch per iteration whatever the number of env copies, no per-step VecEnv call, no
    replay add from the host."""
    res, _ = _run_example("maddpg_two_agents.py", "--n-envs", "8192", "--iters", "110", "--batch", "256", *flag)
    assert res["algo"] == algo and res["transitions"] == 110 * 4 * 8192 and res["rollout_launches"] == 110
    assert res["updates"] == 4 * (110 - 1)  # the first two iterations (8 ring rows) collect uniform warm-up actions; updates start after the second
    assert res["env_step_launches"] <= 1 and res["peer_error"] == 0  # (the reset)
    assert -2000.0 < res["episode_return_last"] < 0.0  # 400 steps of rewards in [-7.2, 0]; the reference's MADDPG does not improve on this task either


def test_bcq_example_trains_on_the_device_resident_dataset():
    """examples/bcq_offline.py at a reduced dataset size: dataset written by the fused rollout kernel, every update cstr_bcq_update; the VAE and
    critic losses must fall, and the eager-torch twin is timed beside it."""
    res, _ = _run_example("bcq_offline.py", "--transitions", "1024000", "--updates", "400", "--batch", "256", "--torch")
    assert res["transitions"] == 1_024_000 and res["updates"] == 400 and res["sizes"]["latent"] == 32
    assert res["vae_loss_first_last"][1] < res["vae_loss_first_last"][0]
    assert res["fused_ms_per_update"] < res["torch_eager_ms_per_update"]
    assert all(np.isfinite(v) for v in res["critic_loss_first_last"])


def test_sac_example_learns_end_to_end():
    """examples/sac_fused_rollout.py: squashed-Gaussian actor on tcgen05 in the fused rollout -> Philox sample -> fused SAC update."""
    res, _ = _run_example("sac_fused_rollout.py", "--n-envs", "16384", "--iters", "1200")
    assert res["transitions"] == 1200 * res["steps_per_iter"] * 16384 and res["peer_error"] == 0
    assert 400 * res["mean_reward_last"] > 400 * res["mean_reward_first"] + 100, res


@pytest.mark.parametrize("actor_mode,atol", [("fp32", 3e-6), ("tc", 5e-3)])
def test_sac_gaussian_actor_rollout(pkg, actor_mode, atol):
    """SAC-shaped actor (4-256-256, mu/log_std heads, tanh-squashed Gaussian) in the fused rollout, eps injected."""
    torch.manual_seed(3)
    trunk = torch.nn.Sequential(torch.nn.Linear(4, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU())
    mu, log_std = torch.nn.Linear(256, 2), torch.nn.Linear(256, 2)
    with torch.no_grad():
        log_std.bias.add_(-1.0)
    actor = pkg.ActorWeights.from_sac_actor(trunk, mu, log_std)
    assert actor.kind == "gaussian" and tuple(actor.W3.shape) == (4, 256)
    n, K = 128 * 3 + 5, 9  # 9 steps x 4 K-chunks: every mbarrier phase wraps several times with an even chunk count
    env = pkg.GpuCSTRVecEnv(n, seed=2, monitor=False)
    env.reset()
    buf = pkg.GpuReplayBuffer(16 * n, device="cuda", n_envs=n)
    eps = np.random.default_rng(0).standard_normal((K, n, 2)).astype(np.float32)
    state = env.state.cpu().numpy().copy()
    pkg.FusedRollout(env, buf, actor, sigma=0.3, actor_mode=actor_mode).collect(K, noise=torch.as_tensor(eps, device="cuda"))
    rec = buf.records.cpu().numpy()
    lin = [m for m in trunk if isinstance(m, torch.nn.Linear)]
    w = [(lin[0].weight.detach().numpy(), lin[0].bias.detach().numpy()), (lin[1].weight.detach().numpy(), lin[1].bias.detach().numpy()),
         (np.concatenate([mu.weight.detach().numpy(), log_std.weight.detach().numpy()]), np.concatenate([mu.bias.detach().numpy(), log_std.bias.detach().numpy()]))]
    sc = np.zeros(n, np.int32)
    for k in range(K):
        assert np.array_equal(rec[k, :, 0:4], state)
        a = O.sac_actor_forward(state, w, eps[k]).astype(np.float32)
        _, a_buf = O.sample_action_maps(a, np.zeros_like(a))  # SAC adds no action noise (sigma is ignored)
        np.testing.assert_allclose(rec[k, :, 8:10], a_buf, rtol=0, atol=atol)
        # reference torch actor on the same eps (distributions.py): tanh(mean + std * eps)
        with torch.no_grad():
            lat = trunk(torch.as_tensor(state))
            ref = torch.tanh(mu(lat) + torch.exp(torch.clamp(log_std(lat), -20, 2)) * torch.as_tensor(eps[k])).numpy()
        np.testing.assert_allclose(rec[k, :, 8:10], ref, rtol=0, atol=max(atol, 5e-6))
        state, r, tr, sc, _ = B.step_f32(state, _env_action_from_buffer_action(rec[k, :, 8:10]), sc, exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)
        assert np.array_equal(rec[k, :, 4:8], state) and np.array_equal(rec[k, :, 10], r)
    # Philox eps: the spread of the stored action around the deterministic one matches std = exp(log_std)
    env2 = pkg.GpuCSTRVecEnv(128 * 64, seed=9, monitor=False)
    env2.reset()
    buf2 = pkg.GpuReplayBuffer(2 * env2.num_envs, device="cuda", n_envs=env2.num_envs)
    pkg.FusedRollout(env2, buf2, actor, actor_mode=actor_mode).collect(1)
    a2 = buf2.records[0, :, 8:10].cpu().numpy()
    assert np.abs(a2).max() <= 1.0 and 0.05 < a2.std() < 1.0


@pytest.mark.parametrize("actor_mode", ["fp32", "tc"])
def test_device_episode_stats(pkg, golden, actor_mode):
    """Monitor semantics on the device: every finished episode's (return, length) comes back exactly once and the
    return equals the sum of that reactor's recorded rewards since its last reset."""
    g, _ = _actor(golden)
    n, K = 128 * 2 + 9, 12
    env = pkg.GpuCSTRVecEnv(n, seed=6, monitor=False)
    env.reset()
    start = torch.arange(n, device="cuda", dtype=torch.int32) % 7 + 392  # reactors truncate at different steps of the launch
    env.step_count.copy_(start)
    buf = pkg.GpuReplayBuffer(16 * n, device="cuda", n_envs=n)
    actor = pkg.ActorWeights(g["W1"], g["b1"], g["W2"], g["b2"], g["W3"], g["b3"])
    stats = pkg.EpisodeStats(n)
    roll = pkg.FusedRollout(env, buf, actor, sigma=0.1, actor_mode=actor_mode)
    roll.collect(K, stats=stats)
    fin = stats.pop()
    rec = buf.records.cpu().numpy()[:K]
    rewards, dones = rec[:, :, 10], rec[:, :, 11] > 0
    assert len(fin) == n == int(dones.sum())  # each reactor finishes exactly one episode in this window
    first_done = dones.argmax(axis=0)
    expect_ret = np.array([rewards[: first_done[i] + 1, i].astype(np.float64).sum() for i in range(n)])
    assert sorted(np.round(fin[:, 0], 3)) == pytest.approx(sorted(np.round(expect_ret, 3)), abs=2e-3)
    assert (fin[:, 1] == 400).all()
    # the running returns now hold the partial sums of the new episodes
    tail = np.array([rewards[first_done[i] + 1:, i].astype(np.float64).sum() for i in range(n)])
    np.testing.assert_allclose(stats.ep_return.cpu().numpy(), tail, rtol=1e-6, atol=1e-6)
    assert len(stats.pop()) == 0
    roll.collect(3, stats=stats)
    assert len(stats.pop()) == 0  # nobody finishes in the next 3 steps


def test_actor_weights_refresh_from_tensors(pkg, golden):
    """ActorWeights.refresh_from_tensors (the hand-over from a FusedTD3Update / FusedSACUpdate flat block to the rollout kernel): the
    fp32 tensors are copied in place and the bf16 UMMA image of W2 is rebuilt, so the next collect() acts with the new weights."""
    g, _ = _actor(golden)
    n = 256
    actor = pkg.ActorWeights(g["W1"], g["b1"], g["W2"], g["b2"], g["W3"], g["b3"])
    env = pkg.GpuCSTRVecEnv(n, seed=3, monitor=False)
    env.reset()
    buf = pkg.GpuReplayBuffer(8 * n, device="cuda", n_envs=n)
    roll = pkg.FusedRollout(env, buf, actor, sigma=0.0, actor_mode="tc")
    roll.collect(1)
    a_old = buf.records[0, :, 8:10].clone()
    packed_old = actor.packed_bf16.clone()
    new = [torch.as_tensor(g[k]).cuda() for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
    new[2] = -new[2]  # flip the hidden layer: a different policy
    ptrs = [t.data_ptr() for t in (actor.W1, actor.W2, actor.packed_bf16)]
    actor.refresh_from_tensors(new)
    assert ptrs == [t.data_ptr() for t in (actor.W1, actor.W2, actor.packed_bf16)]  # in place: the kernel's pointers stay valid
    assert torch.equal(actor.W2, new[2]) and not torch.equal(actor.packed_bf16, packed_old)
    env.reset()
    buf.reset()
    roll.collect(1)
    assert not torch.allclose(buf.records[0, :, 8:10], a_old, atol=1e-3)
