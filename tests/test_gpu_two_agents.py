"""-m gpu: the two-agent fused rollout kernel (cstr_rollout_fused_multi, BASELINE config #5) against the oracle composition.
Kept in a file of its own that sorts last: it was added after the round's last complete GPU run of the suite."""
import numpy as np
import pytest
import torch

import build_oracle as B
import cstr_oracle as O

pytestmark = pytest.mark.gpu


def test_multi_agent_rollout_matches_oracle_composition(pkg):
    """cstr_rollout_fused_multi (BASELINE config #5: the two reactors as two agents): per step, agent i's actor 2 -> H1 -> H2 -> 1 on its observation
    slice, predict()'s unscale and nothing else (the reference's multi-agent _sample_action never applies noise or rescaling: quirk Q5,
    multiagent_policy_algorithm.py:369,390-392), the strict env step, the replay record.  Stored action vs the oracle actor: 3e-6 (float32
    summation order); record given the stored action: bit-exact."""
    n, K, seed, H1, H2 = 300, 3, 13, 400, 300  # a ragged last tile
    rng = np.random.default_rng(2)
    torch.manual_seed(0)
    agents = [torch.nn.Sequential(torch.nn.Linear(2, H1), torch.nn.ReLU(), torch.nn.Linear(H1, H2), torch.nn.ReLU(), torch.nn.Linear(H2, 1), torch.nn.Tanh()).cuda()
              for _ in range(2)]
    env = pkg.GpuCSTRVecEnv(n, seed=seed, monitor=False)
    env.reset()
    state = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    env.set_state(state, np.full(n, 397, np.int32))  # crosses the truncation row
    buf = pkg.GpuReplayBuffer(8 * n, device="cuda", n_envs=n)
    actor = pkg.AgentActorWeights(agents, device="cuda")
    roll = pkg.FusedRollout(env, buf, actor, sigma=0.3)  # sigma is ignored in multi-agent mode (Q5)
    stats = pkg.EpisodeStats(n)
    roll.collect(K, stats=stats)
    rec = buf.records.cpu().numpy()
    f = np.float32
    sc = np.full(n, 397, np.int32)
    for k in range(K):
        assert np.array_equal(rec[k, :, 0:4], state)
        mus = []
        for i, m in enumerate(agents):
            lin = [l for l in m if isinstance(l, torch.nn.Linear)]
            w = [(l.weight.detach().cpu().numpy(), l.bias.detach().cpu().numpy()) for l in lin]
            mus.append(O.actor_forward(state[:, 2 * i:2 * i + 2], w).astype(f))
        mu = np.concatenate(mus, 1)
        u = (f(-1.0) + (f(0.5) * (mu + f(1.0)) * f(2.0))).astype(f)
        np.testing.assert_allclose(rec[k, :, 8:10], u, rtol=0, atol=3e-6)
        s, r, tr, sc, _ = B.step_f32(state, rec[k, :, 8:10], sc, exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)  # the env receives the stored action itself
        assert np.array_equal(rec[k, :, 4:8], s) and np.array_equal(rec[k, :, 10], r) and np.array_equal(rec[k, :, 11], tr.astype(f))
        state = s
        if tr.any():
            st_reset, _, _, _ = B.reset_f32(n, 0, seed, 0, episode=np.ones(n, np.int32))
            state = np.where(tr[:, None], st_reset, s)
            sc = np.where(tr, 0, sc).astype(np.int32)
    assert np.array_equal(env.state.cpu().numpy(), state)
    assert stats.pop().shape == (n, 2)
    # a changed weight reaches the kernel through refresh_from_modules only
    with torch.no_grad():
        agents[1][4].bias.add_(0.5)
    actor.refresh_from_modules(agents)
    roll.collect(1)
    assert float((buf.records[K, :, 9] - buf.records[K - 1, :, 9]).abs().mean().item()) > 0.05
    # warm-up: raw uniform actions, stored and applied as drawn
    roll.collect(2, warmup=True)
    rw = buf.records[K + 1:K + 3].cpu().numpy()
    assert np.abs(rw[..., 8:10]).max() <= 1.0 and abs(float(rw[..., 8:10].mean())) < 0.08
    s2, _, _, _, _ = B.step_f32(rw[0, :, 0:4], rw[0, :, 8:10], np.zeros(n, np.int32), exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)
    assert np.array_equal(rw[0, :, 4:8], s2)
