"""-m gpu: ring replay buffer kernels vs the reference fixture / oracle — bit-exact."""
import pickle

import numpy as np
import pytest
import torch

import cstr_oracle as O

pytestmark = pytest.mark.gpu
STORES = ("observations", "next_observations", "actions", "rewards", "dones", "timeouts")


def _fill(buf, g, tag, n_add):
    for i in range(n_add):
        infos = [{"TimeLimit.truncated": bool(x)} for x in g[f"{tag}_add_timeout"][i]]
        buf.add(g[f"{tag}_add_obs"][i], g[f"{tag}_add_next_obs"][i], g[f"{tag}_add_action"][i], g[f"{tag}_add_reward"][i],
                g[f"{tag}_add_done"][i], infos)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_add_and_sample_match_reference_fixture(pkg, golden, tag):
    g = golden("replay.npz")
    size, n_envs, n_add, batch = (int(v) for v in g[f"{tag}_cfg"])
    buf = pkg.GpuReplayBuffer(size, device="cuda", n_envs=n_envs)
    assert buf.buffer_size == max(size // n_envs, 1)
    _fill(buf, g, tag, n_add)
    for name in STORES:
        got = getattr(buf, name).cpu().numpy()
        assert got.shape == g[f"{tag}_store_{name}"].shape and np.array_equal(got, g[f"{tag}_store_{name}"]), name
    assert buf.pos == int(g[f"{tag}_pos"]) and buf.full == bool(g[f"{tag}_full"]) and buf.size() == (buf.buffer_size if buf.full else buf.pos)
    np.random.seed(11)  # same global-RNG draws as the reference (buffers.py:114,309)
    s = buf.sample(batch)
    for field, key in (("observations", "s_obs"), ("actions", "s_act"), ("next_observations", "s_next_obs"), ("dones", "s_dones"),
                       ("rewards", "s_rewards")):
        t = getattr(s, field)
        assert t.dtype == torch.float32 and t.is_cuda
        assert np.array_equal(t.cpu().numpy(), g[f"{tag}_{key}"]), field
    assert s.dones.shape == (batch, 1) and s.rewards.shape == (batch, 1)


def test_device_tensor_add_and_views(pkg):
    n = 1000
    rng = np.random.default_rng(0)
    buf = pkg.GpuReplayBuffer(4 * n, device="cuda", n_envs=n)
    ora = O.ReplayOracle(4 * n, n)
    for _ in range(6):  # wraps the 4-row ring
        o, no = rng.random((n, 4), np.float32), rng.random((n, 4), np.float32)
        a, r = rng.random((n, 2), np.float32), rng.random(n, np.float32)
        d = rng.random(n) < 0.3
        to = d & (rng.random(n) < 0.5)
        buf.add(*(torch.as_tensor(x, device="cuda") for x in (o, no, a, r, d)), infos=None, timeouts=torch.as_tensor(to, device="cuda"))
        ora.add(o, no, a, r, d, to)
    for name in STORES:
        assert np.array_equal(getattr(buf, name).cpu().numpy(), getattr(ora, name))
    assert (buf.pos, buf.full) == (ora.pos, ora.full)
    buf.dones[1] = True  # the reference mutates stores in place (off_policy_algorithm.py:290)
    assert bool((buf.records[1, :, 11] == 1).all().item())
    # explicit index pairs incl. repeated and boundary rows
    bi = np.array([0, 3, 3, 1, 2, 0]); ei = np.array([0, n - 1, n - 1, 5, 17, n - 1])
    s = buf.gather(bi, ei)
    ora.dones[1] = 1.0
    obs, act, nobs, dones, rew = ora.gather(bi, ei)
    assert np.array_equal(s.observations.cpu().numpy(), obs) and np.array_equal(s.actions.cpu().numpy(), act)
    assert np.array_equal(s.next_observations.cpu().numpy(), nobs) and np.array_equal(s.dones.cpu().numpy(), dones)
    assert np.array_equal(s.rewards.cpu().numpy(), rew)


def test_philox_sampling(pkg, G=None):
    import gpu_util as G

    n, rows = 500, 8
    buf = pkg.GpuReplayBuffer(rows * n, device="cuda", n_envs=n, index_mode="philox", seed=9)
    buf.records.copy_(torch.rand_like(buf.records))
    with pytest.raises(ValueError):
        buf.sample(4)  # empty
    buf.pos = 5
    lib = G.L.load()
    B_ = 20000
    outs = buf._alloc_out(B_)
    bi = torch.empty(B_, dtype=torch.int64, device="cuda"); ei = torch.empty(B_, dtype=torch.int64, device="cuda")
    rc = lib.cstr_replay_sample_philox(9, 0, n, 5, B_, buf.records.data_ptr(), *(t.data_ptr() for t in outs), bi.data_ptr(), ei.data_ptr(), None, None)
    assert rc == 0
    torch.cuda.synchronize()
    b, e = bi.cpu().numpy(), ei.cpu().numpy()
    assert b.min() == 0 and b.max() == 4 and e.min() >= 0 and e.max() <= n - 1
    assert abs(b.mean() - 2.0) < 0.05 and abs(e.mean() - (n - 1) / 2) < 5  # uniform
    rec = buf.records.cpu().numpy()
    assert np.array_equal(outs[0].cpu().numpy(), rec[b, e, 0:4]) and np.array_equal(outs[4].cpu().numpy()[:, 0], rec[b, e, 10])
    assert np.array_equal(outs[3].cpu().numpy()[:, 0], rec[b, e, 11] * (1 - rec[b, e, 12]))
    s1 = buf.sample(64); s2 = buf.sample(64)
    assert not torch.equal(s1.observations, s2.observations)  # the draw counter advances
    # mirror of the index recipe: counter=(i, draw), stream 4, 64-bit multiply-shift
    i = np.arange(16, dtype=np.uint64)
    ctr = np.stack([i.astype(np.uint32), np.zeros(16, np.uint32), np.zeros(16, np.uint32), np.full(16, 4 << 8, np.uint32)], 1)
    r = O.philox4x32(ctr, np.array([9, 0], np.uint32)).astype(np.uint64)
    w0 = (r[:, 0] << np.uint64(32)) | r[:, 1]
    assert np.array_equal(b[:16], np.array([(int(w) * 5) >> 64 for w in w0]))


def test_pickle_roundtrip_and_reference_layout(pkg, golden):
    g = golden("replay.npz")
    size, n_envs, n_add, batch = (int(v) for v in g["a_cfg"])
    buf = pkg.GpuReplayBuffer(size, device="cuda", n_envs=n_envs)
    _fill(buf, g, "a", n_add)
    blob = pickle.dumps(buf)
    state = pickle.loads(blob).__dict__
    buf2 = pickle.loads(blob)
    assert torch.equal(buf2.records, buf.records) and (buf2.pos, buf2.full) == (buf.pos, buf.full)
    arrays = buf.to_numpy_arrays()
    for name in STORES:  # reference attribute layout: contiguous numpy arrays of the reference's shapes
        assert arrays[name].flags["C_CONTIGUOUS"] and np.array_equal(arrays[name], g[f"a_store_{name}"])

    class FakeRef:  # stands in for an unpickled reference ReplayBuffer (BCQ dataset path, Q8)
        pass

    ref = FakeRef()
    ref.buffer_size, ref.n_envs, ref.pos, ref.full = buf.buffer_size, n_envs, buf.pos, buf.full
    ref.observation_space, ref.action_space = buf.observation_space, buf.action_space
    for name in STORES:
        setattr(ref, name, g[f"a_store_{name}"])
    buf3 = pkg.GpuReplayBuffer.from_reference(ref, device="cuda")
    assert torch.equal(buf3.records, buf.records)


def test_buffer_ctor_contract(pkg):
    with pytest.raises(ValueError):
        pkg.GpuReplayBuffer(10, device="cuda", n_envs=64)  # Q11: size counts transitions
    with pytest.raises(ValueError):
        pkg.GpuReplayBuffer(100, device="cuda", optimize_memory_usage=True)
    with pytest.raises(pkg.CstrLibraryError):
        pkg.GpuReplayBuffer(100, device="cpu")
    b = pkg.GpuReplayBuffer(100, device="auto", n_envs=3)
    assert b.buffer_size == 33 and b.obs_shape == (4,) and b.action_dim == 2 and b.device.type == "cuda"
    b.device = "cuda"  # load_replay_buffer assigns it (off_policy_algorithm.py:254)


def test_bcq_dataset_10m_transitions(pkg):
    """BASELINE config #4 shape: a 10,000,000-transition CSTR dataset (25,000 reactors x 400 steps of the tape kernel
    with random actions) resident on the GPU in the n_envs=1 layout OfflineAlgorithm expects (640 MB of records);
    size-independent properties + spot parity of sampled rows against the oracle step."""
    import build_oracle as B

    n, T = 25_000, 400
    env = pkg.GpuCSTRVecEnv(n, seed=21, monitor=False)
    env.reset()
    obs0 = env.state.clone()
    acts = torch.rand((T, n, 2), device="cuda") * 2 - 1
    res = env.tape(T, acts, want_obs=True)
    buf = pkg.GpuReplayBuffer(n * T, device="cuda", n_envs=1, index_mode="philox", seed=5)
    assert buf.buffer_size == 10_000_000 and buf.records.numel() * 4 == 640_000_000
    obs = torch.cat([obs0[None], res["obs"][:-1]], 0)  # observation before each step
    rec = buf.records.view(T, n, 16)
    rec[..., 0:4], rec[..., 4:8], rec[..., 8:10] = obs, res["obs"], acts
    rec[..., 10], rec[..., 11], rec[..., 12] = res["rewards"], res["dones"].float(), res["dones"].float()
    buf.pos, buf.full = 0, True
    s = buf.sample(65_536)
    assert s.observations.shape == (65_536, 4) and s.dones.shape == (65_536, 1)
    assert float(s.dones.sum().item()) == 0.0  # every done in this dataset is a time-limit truncation: dones*(1-timeouts) == 0
    assert float(s.observations.abs().max().item()) <= 1.0 and float(s.rewards.max().item()) <= 0.0
    # sampled (obs, action) -> (next_obs, reward) must be the strict step, bit for bit (rows that are not the truncation row)
    o, a = s.observations.cpu().numpy(), s.actions.cpu().numpy()
    nx, r, _, _, _ = B.step_f32(o, a, np.zeros(len(o), np.int32), exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)
    same = (nx == s.next_observations.cpu().numpy()).all(axis=1)
    assert same.mean() > 0.995  # the ~1/400 truncation rows store the post-reset observation in this construction
    assert np.array_equal(r[same], s.rewards.cpu().numpy()[same, 0])
    small = buf.sample(256)  # the reference's batch size
    assert small.actions.shape == (256, 2)


class ReferenceLikeReplayBuffer:
    """Stands in for the reference's ``core.common.buffers.ReplayBuffer`` as the base of a bound class (module-level: importable)."""


def test_bound_class_pickles_and_stays_an_instance_of_the_reference_base(pkg, golden, tmp_path):
    """``model.save_replay_buffer`` pickles the bound buffer (save_util.py:339-373) and ``load_replay_buffer`` asserts
    ``isinstance(buffer, ReplayBuffer)`` (off_policy_algorithm.py:239): the dynamic class must survive the round trip."""
    g = golden("replay.npz")
    size, n_envs, n_add, batch = (int(v) for v in g["a_cfg"])
    Bound = pkg.bind_replay_buffer_class(ReferenceLikeReplayBuffer)
    assert pkg.bind_replay_buffer_class(ReferenceLikeReplayBuffer) is Bound
    buf = Bound(size, device="cuda", n_envs=n_envs)
    _fill(buf, g, "a", n_add)
    path = tmp_path / "replay.pkl"
    with open(path, "wb") as f:
        pickle.dump(buf, f)
    with open(path, "rb") as f:
        back = pickle.load(f)
    assert isinstance(back, ReferenceLikeReplayBuffer) and isinstance(back, pkg.GpuReplayBuffer) and type(back) is Bound
    assert torch.equal(back.records, buf.records) and (back.pos, back.full) == (buf.pos, buf.full)
    back.device = "cuda"
    np.random.seed(3)
    s1 = buf.sample(32)
    np.random.seed(3)
    s2 = back.sample(32)
    assert torch.equal(s1.observations, s2.observations) and torch.equal(s1.rewards, s2.rewards)
