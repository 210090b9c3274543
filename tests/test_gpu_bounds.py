"""-m gpu: out-of-bounds guards.  compute-sanitizer is closed on this pool, so every kernel family is run on
ragged sizes with its buffers embedded in larger, canary-filled allocations; bytes outside the logical extent
must be untouched and the in-range result must equal the same call on exact-size buffers."""
from ctypes import byref

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CANARY = -12345.5


def _padded(shape, dtype, pad_elems=64):
    """A canary-filled allocation with `pad_elems` guard elements (of the trailing dim size) on both sides."""
    n = int(np.prod(shape))
    guard = 256  # elements; keeps 16-byte alignment for every dtype used here
    full = torch.full((n + 2 * guard,), CANARY if dtype.is_floating_point else 77, dtype=dtype, device="cuda")
    return full, full[guard:guard + n].view(*shape), guard


def _guards_intact(full, guard, dtype):
    ref = CANARY if dtype.is_floating_point else 77
    return bool((full[:guard] == ref).all().item()) and bool((full[-guard:] == ref).all().item())


@pytest.mark.parametrize("n", [1, 31, 130, 4097])
def test_step_tape_reset_ragged_sizes(pkg, n):
    import gpu_util as G

    lib = G.L.load()
    p = G.params(seed=5)
    T = 3
    bufs = {}
    for name, shape, dt in (("state", (n, 4), torch.float32), ("sc", (n,), torch.int32), ("ep", (n,), torch.int32),
                            ("act", (T, n, 2), torch.float32), ("rew", (T, n), torch.float32), ("done", (T, n), torch.uint8),
                            ("obs", (T, n, 4), torch.float32), ("term", (n, 4), torch.float32), ("r1", (n,), torch.float32),
                            ("d1", (n,), torch.uint8), ("t1", (n,), torch.uint8)):
        bufs[name] = _padded(shape, dt)
    v = {k: b[1] for k, b in bufs.items()}
    v["sc"].zero_(); v["ep"].zero_()
    v["act"].copy_(torch.rand((T, n, 2), device="cuda") * 2 - 1)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.cstr_reset(byref(p), n, None, v["state"].data_ptr(), 0, v["sc"].data_ptr(), v["ep"].data_ptr(), None, st) == 0
    ref_env = pkg.GpuCSTRVecEnv(n, seed=5, monitor=False)
    assert np.array_equal(ref_env.reset(), v["state"].cpu().numpy())
    for mode in (0, 1):
        assert lib.cstr_tape_f32(byref(p), n, T, mode, v["act"].data_ptr(), 0, v["state"].data_ptr(), v["sc"].data_ptr(), v["ep"].data_ptr(), None,
                                 v["rew"].data_ptr(), v["done"].data_ptr(), v["obs"].data_ptr(), None, st) == 0
    assert lib.cstr_vec_step_f32(byref(p), n, 0, 1, v["act"].data_ptr(), v["state"].data_ptr(), v["sc"].data_ptr(), v["ep"].data_ptr(), None,
                                 v["term"].data_ptr(), v["r1"].data_ptr(), v["d1"].data_ptr(), v["t1"].data_ptr(), None, None, None, st) == 0
    torch.cuda.synchronize()
    for name, (full, view, guard) in bufs.items():
        assert _guards_intact(full, guard, full.dtype), name
    assert bool(torch.isfinite(v["rew"]).all().item()) and int(v["sc"].max().item()) == 2 * T + 1


@pytest.mark.parametrize("n", [1, 130, 641])
def test_replay_and_rollout_ragged_sizes(pkg, golden, n):
    g = golden("td3_actor.npz")
    rows, K = 4, 3
    full_rec, rec, guard = _padded((rows, n, 16), torch.float32)
    rec.zero_()
    env = pkg.GpuCSTRVecEnv(n, seed=8, monitor=False)
    env.reset()
    buf = pkg.GpuReplayBuffer(rows * n, device="cuda", n_envs=n, index_mode="philox")
    buf.records = rec  # the buffer's storage now sits between canaries
    actor = pkg.ActorWeights(g["W1"], g["b1"], g["W2"], g["b2"], g["W3"], g["b3"])
    outs = []
    for mode in ("fp32", "tc"):
        e = pkg.GpuCSTRVecEnv(n, seed=8, monitor=False)
        e.reset()
        buf.reset()
        pkg.FusedRollout(e, buf, actor, sigma=0.0, actor_mode=mode).collect(K)
        torch.cuda.synchronize()
        assert _guards_intact(full_rec, guard, torch.float32), mode
        outs.append(rec.clone())
    assert float((outs[0][:K, :, 8:10] - outs[1][:K, :, 8:10]).abs().max().item()) < 5e-3  # fp32 vs tcgen05 actor
    assert torch.equal(outs[0][0, :, 0:4], outs[1][0, :, 0:4])
    o = torch.rand((n, 4), device="cuda")
    buf.add(o, o, torch.rand((n, 2), device="cuda"), torch.rand(n, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda"), None)
    torch.cuda.synchronize()
    assert _guards_intact(full_rec, guard, torch.float32) and buf.pos == 0 and buf.full  # K=3 rows + 1 = wrap of the 4-row ring
    B_ = 257
    fo, vo, go = _padded((B_, 4), torch.float32)
    s = buf.sample(B_)
    assert s.observations.shape == (B_, 4) and bool(torch.isfinite(s.rewards).all().item())
