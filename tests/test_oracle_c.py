"""C restatement (oracle/cstr_oracle.c) vs the NumPy restatement.  CPU only."""
import numpy as np
import pytest

import build_oracle as B
import cstr_oracle as O
from conftest import ulp32


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(2)
    N = 50_000
    st = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
    st[:32] = np.sign(st[:32])
    ac = rng.uniform(-1.3, 1.3, (N, 2)).astype(np.float32)
    ac[40] = [np.nan, 0.0]
    ac[41] = [np.inf, -np.inf]
    sc = (np.arange(N) % 402).astype(np.int32)
    return st, ac, sc


def test_shared_exp_accuracy():
    x = np.linspace(-100, 100, 1_000_001).astype(np.float32)
    y = B.expf_shared(x)
    ref = np.exp(x.astype(np.float64))
    normal = (ref < 3e38) & (ref > 1.2e-38)
    assert ulp32(ref[normal], y[normal]).max() < 1.0  # faithful rounding over the whole clipped range
    assert np.isinf(y[ref > 3.5e38]).all()
    xs = np.linspace(-37.0, -24.0, 500_001).astype(np.float32)  # Arrhenius arguments for T in [273.15, 400] K
    assert ulp32(np.exp(xs.astype(np.float64)), B.expf_shared(xs)).max() < 1.0


def test_c_equals_numpy_in_shared_mode(data):
    st, ac, sc = data
    o = O.step_f32(st, ac, sc, square=O.square_mul, exp=B.expf_shared)
    s, r, tr, sc2, bad = B.step_f32(st, ac, sc, exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)
    assert np.array_equal(s, o.obs) and np.array_equal(r, o.reward)
    assert np.array_equal(tr, o.truncated) and np.array_equal(sc2, o.step_count) and np.array_equal(bad, o.nan_row)


def test_c_libm_mode_within_exp_tolerance(data):
    st, ac, sc = data
    o = O.step_f32(st, ac, sc)
    s, r, tr, _, _ = B.step_f32(st, ac, sc, exp_mode=B.EXP_LIBM, sq_mode=B.SQ_POWF)
    np.testing.assert_allclose(s, o.obs, rtol=0, atol=2e-6)
    np.testing.assert_allclose(r, o.reward, rtol=0, atol=1e-5)
    assert np.array_equal(tr, o.truncated)
    assert np.array_equal(B.powf2(np.abs(st[:, 0])), O._powf2(np.abs(st[:, 0])))


def test_f64_step(data):
    st, ac, sc = data
    st64, ac64 = st[:5000].astype(np.float64), ac[:5000].astype(np.float64)
    o = O.step_f64(st64, ac64, sc[:5000])
    s, r, tr, _, bad = B.step_f64(st64, ac64, sc[:5000])
    ok = ~o.nan_row
    assert np.array_equal(bad, o.nan_row) and np.array_equal(tr, o.truncated)
    np.testing.assert_allclose(s[ok], o.obs[ok], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(r[ok], o.reward[ok], rtol=1e-12, atol=1e-13)


def test_philox_known_answers_and_mirror():
    # Random123 known-answer vectors for philox4x32-10
    z = B.philox(np.zeros((1, 4), np.uint32), np.zeros(2, np.uint32))[0]
    assert [int(v) for v in z] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = B.philox(np.full((1, 4), 0xFFFFFFFF, np.uint32), np.full(2, 0xFFFFFFFF, np.uint32))[0]
    assert [int(v) for v in f] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    p = B.philox(np.array([[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], np.uint32),
                 np.array([0xA4093822, 0x299F31D0], np.uint32))[0]
    assert [int(v) for v in p] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, (512, 4), dtype=np.uint64).astype(np.uint32)
    key = np.array([123, 456], np.uint32)
    assert np.array_equal(O.philox4x32(ctr, key), B.philox(ctr, key))


def test_reset_recipe_and_tape():
    n, seed, env0 = 300, 0xDEADBEEF12345, 5_000_000_000
    ep = (np.arange(n) % 7).astype(np.int32)
    u = B.reset_uniforms(seed, env0, ep)
    assert u.min() >= 0.0 and u.max() < 1.0
    st, sc, ep2, _ = B.reset_f32(n, env0, seed, 0, episode=ep)
    assert np.array_equal(st, O.obs_from_raw_f64(O.initial_state_from_uniforms(u)))
    assert np.array_equal(ep2, ep + 1) and not sc.any()
    # tape == repeated single steps + resets, incl. the step-400 truncation row
    rng = np.random.default_rng(4)
    T = 403
    acts = rng.uniform(-1, 1, (T, n, 2)).astype(np.float32)
    res = B.tape_f32(st, sc, ep2, acts, env0, seed, 0, exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL, want_obs=True,
                     want_term=True)
    s, c, e = st.copy(), sc.copy(), ep2.copy()
    for t in range(T):
        s, r, tr, c, _ = B.step_f32(s, acts[t], c, exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)
        assert np.array_equal(r, res["rewards"][t]) and np.array_equal(tr, res["dones"][t])
        assert np.array_equal(s, res["term"][t])
        if tr.any():
            assert t == 399 and tr.all()
            s, c, e, _ = B.reset_f32(n, env0, seed, 0, episode=e, mask=tr, state=s, step_count=c)
        assert np.array_equal(s, res["obs"][t])
    assert np.array_equal(s, res["state"]) and np.array_equal(c, res["step_count"]) and np.array_equal(e, res["episode"])
