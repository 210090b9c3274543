// K3/K4 — GPU-resident ring replay buffer (sm_100a): coalesced vectorised add, random-index sample.
//
// Replaces ReplayBuffer.add / _get_samples / to_torch (reference core/common/buffers.py:247-325,128-140).
// HBM-bound.  One 64-byte record per transition (layout in include/cstr_b200.h):
//   add    : 42 B read + 64 B written per transition, every access a full-sector coalesced vector op
//   sample : 16 B of indices + one 64 B record read (two whole sectors) + 48 B written per sample
#include "cstr_abi.cuh"
#include "cstr_device.cuh"
#include "cstr_norm.cuh"

namespace cstr {

__global__ void __launch_bounds__(256)
replay_add_kernel(int64_t n, const float4 *__restrict__ obs, const float4 *__restrict__ next_obs, const float2 *__restrict__ action,
                  const float *__restrict__ reward, const uint8_t *__restrict__ done, const uint8_t *__restrict__ timeout,
                  float4 *__restrict__ row /* records + pos*n*4 float4 */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = obs[i], no = next_obs[i];
    const float2 a = action[i];
    const float r = reward[i];
    const float d = done[i] ? 1.0f : 0.0f;
    const float to = (timeout && timeout[i]) ? 1.0f : 0.0f;
    float4 *rec = row + 4 * i;
    rec[0] = o;
    rec[1] = no;
    rec[2] = make_float4(a.x, a.y, r, d);
    rec[3] = make_float4(to, 0.0f, 0.0f, 0.0f);
}

struct NormArgs {  // by-value copy of cstr_norm_params (stats == nullptr: off)
    const double *stats;
    double eps, clip_obs, clip_reward;
    int norm_obs, norm_reward;
};

__device__ __forceinline__ float4 norm4(float4 o, const NormArgs &na) {
    const double *s = na.stats;
    return make_float4(normalize_obs_value(o.x, s[0], s[4], na.eps, na.clip_obs), normalize_obs_value(o.y, s[1], s[5], na.eps, na.clip_obs),
                       normalize_obs_value(o.z, s[2], s[6], na.eps, na.clip_obs), normalize_obs_value(o.w, s[3], s[7], na.eps, na.clip_obs));
}

// 16-byte read-only load that asks L2 to fetch only the 64 bytes around it on a miss (one record = 64 bytes; the default prefetch
// size pulls a whole 128-byte line from DRAM for every random record)
__device__ __forceinline__ float4 ldg64(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L2::64B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void gather_one(const float4 *__restrict__ records, int64_t flat, int64_t i, float4 *out_obs, float2 *out_act,
                                           float4 *out_next_obs, float *out_dones, float *out_rewards, const NormArgs &na) {
    const float4 *rec = records + 4 * flat;
    float4 o = ldg64(rec), no = ldg64(rec + 1);
    const float4 ar = ldg64(rec + 2);
    const float to = ldg64(rec + 3).x;
    float rew = ar.z;
    if (na.stats) {  // VecNormalize inside the gather (buffers.py:314-323)
        if (na.norm_obs) { o = norm4(o, na); no = norm4(no, na); }
        if (na.norm_reward) rew = normalize_reward_value(rew, na.stats[10], na.eps, na.clip_reward);
    }
    out_obs[i] = o;
    out_next_obs[i] = no;
    out_act[i] = make_float2(ar.x, ar.y);
    out_rewards[i] = rew;
    out_dones[i] = __fmul_rn(ar.w, __fsub_rn(1.0f, to));  // dones * (1 - timeouts), buffers.py:322
}

__global__ void __launch_bounds__(256)
replay_sample_kernel(int64_t n_envs, int64_t batch, const int64_t *__restrict__ batch_inds, const int64_t *__restrict__ env_inds,
                     const float4 *__restrict__ records, float4 *out_obs, float2 *out_act, float4 *out_next_obs, float *out_dones,
                     float *out_rewards, NormArgs na) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    gather_one(records, batch_inds[i] * n_envs + env_inds[i], i, out_obs, out_act, out_next_obs, out_dones, out_rewards, na);
}

__global__ void __launch_bounds__(256)
replay_sample_philox_kernel(uint64_t seed, uint64_t draw_arg, const int64_t *__restrict__ draw_dev, int64_t n_envs, int64_t upper, int64_t batch, const float4 *__restrict__ records,
                            float4 *out_obs, float2 *out_act, float4 *out_next_obs, float *out_dones, float *out_rewards,
                            int64_t *out_batch_inds, int64_t *out_env_inds, NormArgs na) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const uint64_t draw = draw_dev ? (uint64_t)*draw_dev : draw_arg;  // device-resident draw counter: capturable in a CUDA graph
    const uint4 r = philox_env(seed, (uint64_t)i, (uint32_t)draw, STREAM_SAMPLE, (uint32_t)(draw >> 32) & 0xffu);
    // multiply-shift range reduction on 64-bit words: floor(u64 * range / 2^64)
    const uint64_t w0 = ((uint64_t)r.x << 32) | r.y, w1 = ((uint64_t)r.z << 32) | r.w;
    const int64_t b = (int64_t)__umul64hi(w0, (uint64_t)upper);
    const int64_t e = (int64_t)__umul64hi(w1, (uint64_t)n_envs);
    if (out_batch_inds) out_batch_inds[i] = b;
    if (out_env_inds) out_env_inds[i] = e;
    gather_one(records, b * n_envs + e, i, out_obs, out_act, out_next_obs, out_dones, out_rewards, na);
}

}  // namespace cstr

using namespace cstr;

extern "C" {

int cstr_replay_add(int64_t n_envs, int64_t pos, const float *obs, const float *next_obs, const float *action, const float *reward,
                    const uint8_t *done, const uint8_t *timeout, float *records, void *stream) {
    if (n_envs < 0 || pos < 0 || !obs || !next_obs || !action || !reward || !done || !records)
        return fail_arg(CSTR_EINVAL, "replay_add: null pointer or negative size");
    if (!aligned(obs, 16) || !aligned(next_obs, 16) || !aligned(action, 8) || !aligned(records, 16))
        return fail_arg(CSTR_EALIGN, "replay_add: obs/next_obs/records 16 B, action 8 B alignment");
    if (n_envs == 0) return 0;
    const int block = 256, grid = (int)((n_envs + block - 1) / block);
    replay_add_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(n_envs, (const float4 *)obs, (const float4 *)next_obs, (const float2 *)action,
                                                                 reward, done, timeout, (float4 *)records + pos * n_envs * 4);
    return check_launch("replay_add_kernel");
}

static NormArgs norm_args(const cstr_norm_params *norm) {
    NormArgs na = {nullptr, 0.0, 0.0, 0.0, 0, 0};
    if (norm && norm->stats) na = {norm->stats, norm->epsilon, norm->clip_obs, norm->clip_reward, norm->norm_obs, norm->norm_reward};
    return na;
}

static int check_sample_out(const float *records, float *out_obs, float *out_act, float *out_next_obs, float *out_dones, float *out_rewards) {
    if (!records || !out_obs || !out_act || !out_next_obs || !out_dones || !out_rewards)
        return fail_arg(CSTR_EINVAL, "replay_sample: null pointer");
    if (!aligned(records, 16) || !aligned(out_obs, 16) || !aligned(out_next_obs, 16) || !aligned(out_act, 8))
        return fail_arg(CSTR_EALIGN, "replay_sample: 16 B / 8 B alignment");
    return 0;
}

int cstr_replay_sample(int64_t n_envs, int64_t batch, const int64_t *batch_inds, const int64_t *env_inds, const float *records,
                       float *out_obs, float *out_act, float *out_next_obs, float *out_dones, float *out_rewards,
                       const cstr_norm_params *norm, void *stream) {
    if (n_envs <= 0 || batch < 0 || !batch_inds || !env_inds) return fail_arg(CSTR_EINVAL, "replay_sample: bad sizes or null indices");
    if (int rc = check_sample_out(records, out_obs, out_act, out_next_obs, out_dones, out_rewards)) return rc;
    if (batch == 0) return 0;
    const int block = 128, grid = (int)((batch + block - 1) / block);
    replay_sample_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(n_envs, batch, batch_inds, env_inds, (const float4 *)records, (float4 *)out_obs,
                                                                    (float2 *)out_act, (float4 *)out_next_obs, out_dones, out_rewards, norm_args(norm));
    return check_launch("replay_sample_kernel");
}

static int sample_philox(uint64_t seed, uint64_t draw, const int64_t *draw_dev, int64_t n_envs, int64_t upper, int64_t batch, const float *records,
                         float *out_obs, float *out_act, float *out_next_obs, float *out_dones, float *out_rewards, int64_t *out_batch_inds,
                         int64_t *out_env_inds, const cstr_norm_params *norm, void *stream) {
    if (n_envs <= 0 || upper <= 0 || batch < 0) return fail_arg(CSTR_EINVAL, "replay_sample_philox: bad sizes (empty buffer?)");
    if (int rc = check_sample_out(records, out_obs, out_act, out_next_obs, out_dones, out_rewards)) return rc;
    if (batch == 0) return 0;
    const int block = 128, grid = (int)((batch + block - 1) / block);
    replay_sample_philox_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(seed, draw, draw_dev, n_envs, upper, batch, (const float4 *)records,
                                                                           (float4 *)out_obs, (float2 *)out_act, (float4 *)out_next_obs,
                                                                           out_dones, out_rewards, out_batch_inds, out_env_inds, norm_args(norm));
    return check_launch("replay_sample_philox_kernel");
}

int cstr_replay_sample_philox(uint64_t seed, uint64_t draw, int64_t n_envs, int64_t upper, int64_t batch, const float *records,
                              float *out_obs, float *out_act, float *out_next_obs, float *out_dones, float *out_rewards,
                              int64_t *out_batch_inds, int64_t *out_env_inds, const cstr_norm_params *norm, void *stream) {
    return sample_philox(seed, draw, nullptr, n_envs, upper, batch, records, out_obs, out_act, out_next_obs, out_dones, out_rewards, out_batch_inds,
                         out_env_inds, norm, stream);
}

int cstr_replay_sample_philox_dev(uint64_t seed, const int64_t *draw_dev, int64_t n_envs, int64_t upper, int64_t batch, const float *records,
                                  float *out_obs, float *out_act, float *out_next_obs, float *out_dones, float *out_rewards,
                                  int64_t *out_batch_inds, int64_t *out_env_inds, const cstr_norm_params *norm, void *stream) {
    if (!draw_dev) return fail_arg(CSTR_EINVAL, "replay_sample_philox_dev: null draw counter");
    return sample_philox(seed, 0, draw_dev, n_envs, upper, batch, records, out_obs, out_act, out_next_obs, out_dones, out_rewards, out_batch_inds,
                         out_env_inds, norm, stream);
}

}  // extern "C"
