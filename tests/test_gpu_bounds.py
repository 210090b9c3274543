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


@pytest.mark.parametrize("gemm", ["fp32", "tensor"])
@pytest.mark.parametrize("B,arch", [(1, [36, 20]), (130, [400, 300]), (641, [132, 260])])
def test_td3_update_ragged_sizes_and_argument_errors(pkg, B, arch, gemm):
    """cstr_td3_update on ragged batches with every state block and the workspace inside canary-guarded allocations."""
    import gpu_util as G

    lib = G.L.load()
    eng = pkg.FusedTD3Update(arch, B, gemm=gemm, policy_delay=1)
    P = eng.param_count
    blocks = {name: _padded((P,), torch.float32) for name in ("params", "targets", "grads", "adam_m", "adam_v")}
    ws = _padded((eng._workspace.numel(),), torch.float32)
    data = {name: _padded(shape, torch.float32) for name, shape in (("obs", (B, 4)), ("act", (B, 2)), ("nobs", (B, 4)), ("done", (B, 1)), ("rew", (B, 1)))}
    torch.manual_seed(B)
    for name in ("params", "targets"):
        blocks[name][1].copy_(torch.randn(P, device="cuda") * 0.1)
    for name in ("grads", "adam_m", "adam_v"):
        blocks[name][1].zero_()
    for name, (_, view, _) in data.items():
        view.copy_(torch.rand(view.shape, device="cuda") * 2 - 1)
    data["done"][1].zero_()
    eng.params, eng.targets, eng.grads, eng.adam_m, eng.adam_v = (blocks[k][1] for k in ("params", "targets", "grads", "adam_m", "adam_v"))
    eng._workspace = ws[1]
    before = eng.params.clone()
    for _ in range(2):
        eng.update(tuple(data[k][1] for k in ("obs", "act", "nobs", "done", "rew")))
    torch.cuda.synchronize()
    for name, (full, _, guard) in {**blocks, **data, "workspace": ws}.items():
        assert _guards_intact(full, guard, torch.float32), name
    assert bool(torch.isfinite(eng.params).all().item()) and not torch.equal(before, eng.params)
    # argument errors: negative code + message, nothing launched
    cfg = eng._config(B)
    st = G.L.Td3State(params=eng.params.data_ptr(), targets=eng.targets.data_ptr(), grads=eng.grads.data_ptr(), adam_m=eng.adam_m.data_ptr(),
                      adam_v=eng.adam_v.data_ptr(), workspace=ws[1].data_ptr(), workspace_bytes=16, losses=None)
    ptrs = [data[k][1].data_ptr() for k in ("obs", "act", "nobs", "done", "rew")]
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.cstr_td3_update(byref(cfg), byref(st), *ptrs, None, 1, 1, 1, 15, stream) < 0 and b"workspace" in lib.cstr_last_error()
    st.workspace_bytes = ws[1].numel() * 4
    assert lib.cstr_td3_update(byref(cfg), byref(st), ptrs[0] + 4, *ptrs[1:], None, 1, 1, 1, 15, stream) < 0 and b"alignment" in lib.cstr_last_error()
    assert lib.cstr_td3_update(byref(cfg), byref(st), None, *ptrs[1:], None, 1, 1, 1, 15, stream) < 0
    assert lib.cstr_td3_update(byref(cfg), byref(st), *ptrs, None, 0, 1, 1, 15, stream) < 0  # counters are 1-based
    cfg.h1 = 37
    assert lib.cstr_td3_update(byref(cfg), byref(st), *ptrs, None, 1, 1, 1, 15, stream) < 0 and lib.cstr_td3_param_count(37, 20) == -1


@pytest.mark.parametrize("B,arch", [(1, [36, 20]), (130, [256, 256]), (641, [132, 260])])
def test_sac_update_ragged_sizes_and_argument_errors(pkg, B, arch):
    import gpu_util as G

    lib = G.L.load()
    eng = pkg.FusedSACUpdate(arch, B)
    P = eng.param_count
    blocks = {name: _padded((P,), torch.float32) for name in ("params", "targets", "grads", "adam_m", "adam_v")}
    ws = _padded((eng._workspace.numel(),), torch.float32)
    data = {name: _padded(shape, torch.float32) for name, shape in (("obs", (B, 4)), ("act", (B, 2)), ("nobs", (B, 4)), ("done", (B, 1)), ("rew", (B, 1)))}
    torch.manual_seed(B)
    blocks["params"][1].copy_(torch.randn(P, device="cuda") * 0.1)
    blocks["targets"][1].copy_(blocks["params"][1])
    for name in ("grads", "adam_m", "adam_v"):
        blocks[name][1].zero_()
    for name, (_, view, _) in data.items():
        view.copy_(torch.rand(view.shape, device="cuda") * 2 - 1)
    data["done"][1].zero_()
    eng.params, eng.targets, eng.grads, eng.adam_m, eng.adam_v = (blocks[k][1] for k in ("params", "targets", "grads", "adam_m", "adam_v"))
    eng._workspace = ws[1]
    for _ in range(2):
        eng.update(tuple(data[k][1] for k in ("obs", "act", "nobs", "done", "rew")))
    torch.cuda.synchronize()
    for name, (full, _, guard) in {**blocks, **data, "workspace": ws}.items():
        assert _guards_intact(full, guard, torch.float32), name
    assert bool(torch.isfinite(eng.params).all().item())
    cfg, st = eng._sac_config(B), eng._state(counters=False)
    ptrs = [data[k][1].data_ptr() for k in ("obs", "act", "nobs", "done", "rew")]
    stream = torch.cuda.current_stream().cuda_stream
    st.workspace_bytes = 16
    assert lib.cstr_sac_update(byref(cfg), byref(st), *ptrs, None, None, 1, 1, 15, stream) < 0 and b"workspace" in lib.cstr_last_error()
    st.workspace_bytes = ws[1].numel() * 4
    assert lib.cstr_sac_update(byref(cfg), byref(st), *ptrs, None, None, 0, 1, 15, stream) < 0
    assert lib.cstr_sac_update(byref(cfg), byref(st), *ptrs, None, None, 1, 1, 0, stream) < 0 and b"phases" in lib.cstr_last_error()
    assert lib.cstr_sac_update(byref(cfg), byref(st), *ptrs, None, None, 1, 1, 16, stream) < 0
    cfg.target_update_interval = 0
    assert lib.cstr_sac_update(byref(cfg), byref(st), *ptrs, None, None, 1, 1, 15, stream) < 0 and lib.cstr_sac_param_count(37, 20) == -1
