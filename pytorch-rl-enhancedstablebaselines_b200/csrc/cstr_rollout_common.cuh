// Pieces shared by the two fused-rollout kernels (fp32 CUDA-core actor, bf16 tcgen05 actor):
// the reference's float32 action plumbing, the Philox exploration noise and the replay record store.
#pragma once
#include "cstr_device.cuh"

namespace cstr {

// float32 action plumbing for a Box(-1,1) action space, same association as the reference:
//   predict():        u = low + (0.5*(mu+1))*(high-low)                      policies.py:375,402-413
//   _sample_action(): s = clip(2*((u-low)/(high-low)) - 1 + noise, -1, 1)    off_policy_algorithm.py:398-402
//                     a = low + (0.5*(s+1))*(high-low)                       :405
// (x+1)-1 is not the identity in float32, so the three maps are applied, not skipped.
__device__ __forceinline__ void action_maps(float mu, float noise, float &env_action, float &buffer_action) {
    const float u = __fadd_rn(-1.0f, __fmul_rn(__fmul_rn(0.5f, __fadd_rn(mu, 1.0f)), 2.0f));
    float s = __fadd_rn(__fmul_rn(2.0f, __fmul_rn(__fadd_rn(u, 1.0f), 0.5f)), -1.0f);
    s = clampf(__fadd_rn(s, noise), -1.0f, 1.0f);
    buffer_action = s;
    env_action = __fadd_rn(-1.0f, __fmul_rn(__fmul_rn(0.5f, __fadd_rn(s, 1.0f)), 2.0f));
}

// N(0,1) pair by Box-Muller from one Philox call (stream "noise", counter = global step)
__device__ __forceinline__ float2 philox_normal2(uint64_t seed, uint64_t env, uint32_t g) {
    const uint4 r = philox_env(seed, env, g, STREAM_NOISE, 0);
    const float u1 = fmaf(u24(r.x), 1.0f, 5.9604644775390625e-08f);  // (0,1]
    const float u2 = u24(r.y);
    const float rad = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

// Actor head -> squashed action in [-1,1] and the additive exploration noise that follows it.
//   TANH (TD3):      a = tanh(o[0..1]);  noise = sigma*N(0,1) (or the caller's tensor)        td3/policies.py:75-78, noise.py:29-48
//   GAUSSIAN (SAC):  a = tanh(mu + exp(clamp(log_std,-20,2))*eps), no additive noise          sac/policies.py:151-168
template <int KIND>
__device__ __forceinline__ void actor_head(const float o[4], float2 nz, float &a0, float &a1, float2 &add_noise) {
    if (KIND == CSTR_ACTOR_TANH) {
        a0 = tanhf(o[0]);
        a1 = tanhf(o[1]);
        add_noise = nz;
    } else {
        const float s0 = expf(clampf(o[2], -20.0f, 2.0f)), s1 = expf(clampf(o[3], -20.0f, 2.0f));
        a0 = tanhf(fmaf(s0, nz.x, o[0]));
        a1 = tanhf(fmaf(s1, nz.y, o[1]));
        add_noise = make_float2(0.0f, 0.0f);
    }
}

// Monitor semantics on device: running return per reactor; finished episodes appended to a compact list.
__device__ __forceinline__ void episode_account(const cstr_episode_stats &st, bool enabled, double &ep_ret, float reward, bool done, int length) {
    if (!enabled) return;
    ep_ret += (double)reward;
    if (done) {
        const uint32_t idx = atomicAdd(st.count, 1u);
        if (idx < st.capacity) {
            st.finished[2 * idx] = (float)ep_ret;
            st.finished[2 * idx + 1] = (float)length;
        }
        ep_ret = 0.0;
    }
}

__device__ __forceinline__ void store_record(float4 *__restrict__ rec, float4 obs, float4 next_obs, float2 act, float reward, bool done) {
    rec[0] = obs;
    rec[1] = next_obs;
    rec[2] = make_float4(act.x, act.y, reward, done ? 1.0f : 0.0f);
    rec[3] = make_float4(done ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);  // timeout == truncated (terminated is always False)
}

}  // namespace cstr
