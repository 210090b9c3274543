// K1 — the CSTR environment step kernels and their C-ABI entry points (sm_100a).
//
//   vec_step_*   one control interval for N reactors with DummyVecEnv.step_wait semantics
//                (reference core/common/vec_env/dummy_vec_env.py:56-73 over twoseriescstr.py:394-454)
//   tape_*       T control intervals per launch; the float4 state, step counter and episode counter
//                stay in registers for the whole launch, HBM is touched only for the action tape
//                (8 B/step) and the requested outputs (reward 4 B, done 1 B, obs 16 B per step)
//   reset        TwoSeriesCSTREnv.reset / generate_initial_state (twoseriescstr.py:167-269)
//
// Roofline: FP32 issue (FMA + ALU pipes) for the tape kernels, HBM for the single-step kernel at
// large N (70 B/step).  See DESIGN.md §Kernels for the per-step instruction and byte budgets.
#include "cstr_abi.cuh"
#include "cstr_device.cuh"

namespace cstr {

char *last_error_buf() {
    static thread_local char buf[256] = {0};
    return buf;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            cached = v;
        else
            return 148;
    }
    return cached;
}

enum { MODE_STRICT = CSTR_MATH_STRICT, MODE_FAST = CSTR_MATH_FAST };

template <int MODE>
__device__ __forceinline__ StepResult step_f32(float4 &s, float2 a, int &sc, float target, int max_steps) {
    if (MODE == MODE_STRICT) return step_strict_f32(s, a, sc, target, max_steps);
    return step_fast_f32(s, a, sc, target, max_steps);
}

// ---------------------------------------------------------------------------------------------------
// reset
// ---------------------------------------------------------------------------------------------------
__global__ void reset_kernel(cstr_env_params p, int64_t n, const uint8_t *__restrict__ mask, void *state, int is_f64,
                             int32_t *__restrict__ step_count, int32_t *__restrict__ episode, double *static_base) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (mask && !mask[i])) return;
    double o[4];
    const int ep = episode[i];
    reset_draw(p.seed, (uint64_t)(p.env_offset + i), (uint32_t)ep, p.init_mode, static_base ? static_base + 4 * i : nullptr, o);
    if (is_f64) {
        double2 *s = reinterpret_cast<double2 *>(state) + 2 * i;
        s[0] = make_double2(o[0], o[1]);
        s[1] = make_double2(o[2], o[3]);
    } else {
        reinterpret_cast<float4 *>(state)[i] = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
    }
    step_count[i] = 0;
    episode[i] = ep + 1;
}

// ---------------------------------------------------------------------------------------------------
// one VecEnv step
// ---------------------------------------------------------------------------------------------------
// Monitor-style episode accounting (core/common/monitor.py:85-111) on device
__device__ __forceinline__ void episode_stats(int64_t i, double reward, bool done, int length, double *ep_return,
                                              double *ep_final_return, int32_t *ep_final_length) {
    if (!ep_return) return;
    double er = ep_return[i] + reward;
    if (done) {
        if (ep_final_return) ep_final_return[i] = er;
        if (ep_final_length) ep_final_length[i] = length;
        er = 0.0;
    }
    ep_return[i] = er;
}

template <int MODE>
__global__ void __launch_bounds__(256)
vec_step_f32_kernel(cstr_env_params p, int64_t n, int auto_reset, const float2 *__restrict__ actions, float4 *__restrict__ state,
                    int32_t *__restrict__ step_count, int32_t *__restrict__ episode, double *static_base,
                    float4 *__restrict__ terminal_obs, float *__restrict__ reward, uint8_t *__restrict__ done,
                    uint8_t *__restrict__ timeout, double *__restrict__ ep_return, double *__restrict__ ep_final_return,
                    int32_t *__restrict__ ep_final_length) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 s = state[i];
    const float2 a = actions[i];
    int sc = step_count[i];
    const StepResult r = step_f32<MODE>(s, a, sc, p.target_c2, p.max_steps);
    if (terminal_obs) terminal_obs[i] = s;
    reward[i] = r.reward;
    done[i] = (uint8_t)r.truncated;
    if (timeout) timeout[i] = (uint8_t)r.truncated;  // terminated is always False -> timeout == truncated
    episode_stats(i, (double)r.reward, r.truncated, sc, ep_return, ep_final_return, ep_final_length);
    if (r.truncated && auto_reset) {
        int ep = episode[i];
        s = reset_f32_env(p, i, ep, static_base);
        episode[i] = ep;
        sc = 0;
    }
    state[i] = s;
    step_count[i] = sc;
}

__global__ void __launch_bounds__(256)
vec_step_f64_kernel(cstr_env_params p, int64_t n, int auto_reset, const double2 *__restrict__ actions, double2 *__restrict__ state,
                    int32_t *__restrict__ step_count, int32_t *__restrict__ episode, double *static_base,
                    double2 *__restrict__ terminal_obs, double *__restrict__ reward, uint8_t *__restrict__ done,
                    uint8_t *__restrict__ timeout, double *__restrict__ ep_return, double *__restrict__ ep_final_return,
                    int32_t *__restrict__ ep_final_length) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 s01 = state[2 * i], s23 = state[2 * i + 1];
    double s[4] = {s01.x, s01.y, s23.x, s23.y};
    const double2 a = actions[i];
    int sc = step_count[i];
    const StepResult64 r = step_f64(s, a.x, a.y, sc, (double)p.target_c2, p.max_steps);
    if (terminal_obs) {
        terminal_obs[2 * i] = make_double2(s[0], s[1]);
        terminal_obs[2 * i + 1] = make_double2(s[2], s[3]);
    }
    reward[i] = r.reward;
    done[i] = (uint8_t)r.truncated;
    if (timeout) timeout[i] = (uint8_t)r.truncated;
    episode_stats(i, r.reward, r.truncated, sc, ep_return, ep_final_return, ep_final_length);
    if (r.truncated && auto_reset) {
        const int ep = episode[i];
        reset_draw(p.seed, (uint64_t)(p.env_offset + i), (uint32_t)ep, p.init_mode, static_base ? static_base + 4 * i : nullptr, s);
        episode[i] = ep + 1;
        sc = 0;
    }
    state[2 * i] = make_double2(s[0], s[1]);
    state[2 * i + 1] = make_double2(s[2], s[3]);
    step_count[i] = sc;
}

// ---------------------------------------------------------------------------------------------------
// T steps per launch, state in registers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum_to(double v, double *dst) {
    // warp shuffle reduce, then one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(dst, v);
}

template <int MODE, bool PHILOX>
__global__ void __launch_bounds__(256)
tape_f32_kernel(cstr_env_params p, int64_t n, int64_t T, const float2 *__restrict__ actions, uint32_t t_base,
                float4 *__restrict__ state, int32_t *__restrict__ step_count, int32_t *__restrict__ episode, double *static_base,
                float *__restrict__ rewards, uint8_t *__restrict__ dones, float4 *__restrict__ obs_tape, double *reward_sum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    double acc = 0.0;
    if (live) {
        float4 s = state[i];
        int sc = step_count[i], ep = episode[i];
        const uint64_t env = (uint64_t)(p.env_offset + i);
        uint4 cache = make_uint4(0, 0, 0, 0);
        float2 a_next = make_float2(0.f, 0.f);
        if (!PHILOX) a_next = actions[i];
        for (int64_t t = 0; t < T; ++t) {
            float2 a;
            if (PHILOX) {
                const uint32_t g = t_base + (uint32_t)t;
                a = philox_action(p.seed, env, g, cache, t == 0 || (g & 1u) == 0);
            } else {
                a = a_next;
                if (t + 1 < T) a_next = actions[(t + 1) * n + i];  // prefetch: the load overlaps this step's math
            }
            const StepResult r = step_f32<MODE>(s, a, sc, p.target_c2, p.max_steps);
            if (rewards) rewards[t * n + i] = r.reward;
            if (dones) dones[t * n + i] = (uint8_t)r.truncated;
            if (reward_sum) acc += (double)r.reward;
            if (r.truncated) {
                s = reset_f32_env(p, i, ep, static_base);
                sc = 0;
            }
            if (obs_tape) obs_tape[t * n + i] = s;
        }
        state[i] = s;
        step_count[i] = sc;
        episode[i] = ep;
    }
    if (reward_sum) block_sum_to(acc, reward_sum);
}

template <bool PHILOX>
__global__ void __launch_bounds__(256)
tape_f64_kernel(cstr_env_params p, int64_t n, int64_t T, const double2 *__restrict__ actions, uint32_t t_base,
                double2 *__restrict__ state, int32_t *__restrict__ step_count, int32_t *__restrict__ episode, double *static_base,
                double *__restrict__ rewards, uint8_t *__restrict__ dones, double2 *__restrict__ obs_tape, double *reward_sum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    double acc = 0.0;
    if (live) {
        const double2 s01 = state[2 * i], s23 = state[2 * i + 1];
        double s[4] = {s01.x, s01.y, s23.x, s23.y};
        int sc = step_count[i], ep = episode[i];
        const uint64_t env = (uint64_t)(p.env_offset + i);
        uint4 cache = make_uint4(0, 0, 0, 0);
        for (int64_t t = 0; t < T; ++t) {
            double a0, a1;
            if (PHILOX) {
                const uint32_t g = t_base + (uint32_t)t;
                const float2 a = philox_action(p.seed, env, g, cache, t == 0 || (g & 1u) == 0);
                a0 = (double)a.x;
                a1 = (double)a.y;
            } else {
                const double2 a = actions[t * n + i];
                a0 = a.x;
                a1 = a.y;
            }
            const StepResult64 r = step_f64(s, a0, a1, sc, (double)p.target_c2, p.max_steps);
            if (rewards) rewards[t * n + i] = r.reward;
            if (dones) dones[t * n + i] = (uint8_t)r.truncated;
            if (reward_sum) acc += r.reward;
            if (r.truncated) {
                reset_draw(p.seed, env, (uint32_t)ep, p.init_mode, static_base ? static_base + 4 * i : nullptr, s);
                ep += 1;
                sc = 0;
            }
            if (obs_tape) {
                obs_tape[2 * (t * n + i)] = make_double2(s[0], s[1]);
                obs_tape[2 * (t * n + i) + 1] = make_double2(s[2], s[3]);
            }
        }
        state[2 * i] = make_double2(s[0], s[1]);
        state[2 * i + 1] = make_double2(s[2], s[3]);
        step_count[i] = sc;
        episode[i] = ep;
    }
    if (reward_sum) block_sum_to(acc, reward_sum);
}

// host scratch for cstr_tape_f32_host (grow-only, per process)
struct HostTapeScratch {
    void *buf = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (buf) cudaFree(buf);
        buf = nullptr;
        cap = 0;
        const int rc = check_cuda(cudaMalloc(&buf, bytes), "cudaMalloc(tape scratch)");
        if (rc == 0) cap = bytes;
        return rc;
    }
};
static HostTapeScratch g_scratch;

}  // namespace cstr

using namespace cstr;

extern "C" {

int cstr_b200_abi_version(void) { return CSTR_B200_ABI_VERSION; }

const char *cstr_last_error(void) { return last_error_buf(); }

int cstr_device_info(int *sm_count_out, int *sm_clock_khz, int *cc_major, int *cc_minor) {
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    int v = 0;
    if (sm_count_out) {
        if ((rc = check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev), "attr"))) return rc;
        *sm_count_out = v;
    }
    if (sm_clock_khz) {
        if ((rc = check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev), "attr"))) return rc;
        *sm_clock_khz = v;
    }
    if (cc_major) {
        if ((rc = check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev), "attr"))) return rc;
        *cc_major = v;
    }
    if (cc_minor) {
        if ((rc = check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev), "attr"))) return rc;
        *cc_minor = v;
    }
    return 0;
}

static int check_env_args(const cstr_env_params *p, int64_t n, const void *state, const void *step_count, const void *episode,
                          const void *static_base, size_t state_align) {
    if (!p || n < 0 || !state || !step_count || !episode) return fail_arg(CSTR_EINVAL, "null pointer or negative n");
    if (p->init_mode != CSTR_INIT_RANDOM && p->init_mode != CSTR_INIT_STATIC) return fail_arg(CSTR_EINVAL, "unknown init_mode");
    if (p->init_mode == CSTR_INIT_STATIC && !static_base) return fail_arg(CSTR_EINVAL, "static init_mode needs static_base");
    if (!aligned(state, state_align)) return fail_arg(CSTR_EALIGN, "state must be 16-byte aligned");
    return 0;
}

int cstr_reset(const cstr_env_params *p, int64_t n, const uint8_t *mask, void *state, int is_f64, int32_t *step_count,
               int32_t *episode, double *static_base, void *stream) {
    if (int rc = check_env_args(p, n, state, step_count, episode, static_base, 16)) return rc;
    if (n == 0) return 0;
    const int block = 128, grid = (int)((n + block - 1) / block);
    reset_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*p, n, mask, state, is_f64, step_count, episode, static_base);
    return check_launch("reset_kernel");
}

int cstr_vec_step_f32(const cstr_env_params *p, int64_t n, int math_mode, int auto_reset, const float *actions, float *state,
                      int32_t *step_count, int32_t *episode, double *static_base, float *terminal_obs, float *reward,
                      uint8_t *done, uint8_t *timeout, double *ep_return, double *ep_final_return, int32_t *ep_final_length,
                      void *stream) {
    if (int rc = check_env_args(p, n, state, step_count, episode, static_base, 16)) return rc;
    if (!actions || !reward || !done) return fail_arg(CSTR_EINVAL, "null actions/reward/done");
    if (!aligned(actions, 8) || (terminal_obs && !aligned(terminal_obs, 16))) return fail_arg(CSTR_EALIGN, "actions 8 B / terminal_obs 16 B alignment");
    if (math_mode != CSTR_MATH_STRICT && math_mode != CSTR_MATH_FAST) return fail_arg(CSTR_EINVAL, "unknown math_mode");
    if (n == 0) return 0;
    int grid, block;
    env_launch_geometry(n, grid, block);
    cudaStream_t st = (cudaStream_t)stream;
    if (math_mode == CSTR_MATH_STRICT)
        vec_step_f32_kernel<MODE_STRICT><<<grid, block, 0, st>>>(*p, n, auto_reset, (const float2 *)actions, (float4 *)state, step_count,
                                                                 episode, static_base, (float4 *)terminal_obs, reward, done, timeout,
                                                                 ep_return, ep_final_return, ep_final_length);
    else
        vec_step_f32_kernel<MODE_FAST><<<grid, block, 0, st>>>(*p, n, auto_reset, (const float2 *)actions, (float4 *)state, step_count,
                                                               episode, static_base, (float4 *)terminal_obs, reward, done, timeout,
                                                               ep_return, ep_final_return, ep_final_length);
    return check_launch("vec_step_f32_kernel");
}

int cstr_vec_step_f64(const cstr_env_params *p, int64_t n, int auto_reset, const double *actions, double *state, int32_t *step_count,
                      int32_t *episode, double *static_base, double *terminal_obs, double *reward, uint8_t *done, uint8_t *timeout,
                      double *ep_return, double *ep_final_return, int32_t *ep_final_length, void *stream) {
    if (int rc = check_env_args(p, n, state, step_count, episode, static_base, 16)) return rc;
    if (!actions || !reward || !done) return fail_arg(CSTR_EINVAL, "null actions/reward/done");
    if (!aligned(actions, 16) || (terminal_obs && !aligned(terminal_obs, 16))) return fail_arg(CSTR_EALIGN, "16 B alignment");
    if (n == 0) return 0;
    int grid, block;
    env_launch_geometry(n, grid, block);
    vec_step_f64_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*p, n, auto_reset, (const double2 *)actions, (double2 *)state, step_count,
                                                                   episode, static_base, (double2 *)terminal_obs, reward, done, timeout,
                                                                   ep_return, ep_final_return, ep_final_length);
    return check_launch("vec_step_f64_kernel");
}

int cstr_tape_f32(const cstr_env_params *p, int64_t n, int64_t T, int math_mode, const float *actions, uint32_t t_base, float *state,
                  int32_t *step_count, int32_t *episode, double *static_base, float *rewards, uint8_t *dones, float *obs_tape,
                  double *reward_sum, void *stream) {
    if (int rc = check_env_args(p, n, state, step_count, episode, static_base, 16)) return rc;
    if (T < 0) return fail_arg(CSTR_EINVAL, "negative T");
    if ((actions && !aligned(actions, 8)) || (obs_tape && !aligned(obs_tape, 16))) return fail_arg(CSTR_EALIGN, "actions 8 B / obs_tape 16 B alignment");
    if (math_mode != CSTR_MATH_STRICT && math_mode != CSTR_MATH_FAST) return fail_arg(CSTR_EINVAL, "unknown math_mode");
    if (n == 0 || T == 0) return 0;
    int grid, block;
    env_launch_geometry(n, grid, block);
    cudaStream_t st = (cudaStream_t)stream;
#define CSTR_LAUNCH_TAPE(MODE, PH)                                                                                              \
    tape_f32_kernel<MODE, PH><<<grid, block, 0, st>>>(*p, n, T, (const float2 *)actions, t_base, (float4 *)state, step_count, episode, \
                                                      static_base, rewards, dones, (float4 *)obs_tape, reward_sum)
    if (math_mode == CSTR_MATH_STRICT) {
        if (actions) CSTR_LAUNCH_TAPE(MODE_STRICT, false); else CSTR_LAUNCH_TAPE(MODE_STRICT, true);
    } else {
        if (actions) CSTR_LAUNCH_TAPE(MODE_FAST, false); else CSTR_LAUNCH_TAPE(MODE_FAST, true);
    }
#undef CSTR_LAUNCH_TAPE
    return check_launch("tape_f32_kernel");
}

int cstr_tape_f64(const cstr_env_params *p, int64_t n, int64_t T, const double *actions, uint32_t t_base, double *state,
                  int32_t *step_count, int32_t *episode, double *static_base, double *rewards, uint8_t *dones, double *obs_tape,
                  double *reward_sum, void *stream) {
    if (int rc = check_env_args(p, n, state, step_count, episode, static_base, 16)) return rc;
    if (T < 0) return fail_arg(CSTR_EINVAL, "negative T");
    if ((actions && !aligned(actions, 16)) || (obs_tape && !aligned(obs_tape, 16))) return fail_arg(CSTR_EALIGN, "16 B alignment");
    if (n == 0 || T == 0) return 0;
    int grid, block;
    env_launch_geometry(n, grid, block);
    cudaStream_t st = (cudaStream_t)stream;
    if (actions)
        tape_f64_kernel<false><<<grid, block, 0, st>>>(*p, n, T, (const double2 *)actions, t_base, (double2 *)state, step_count, episode,
                                                       static_base, rewards, dones, (double2 *)obs_tape, reward_sum);
    else
        tape_f64_kernel<true><<<grid, block, 0, st>>>(*p, n, T, nullptr, t_base, (double2 *)state, step_count, episode, static_base,
                                                      rewards, dones, (double2 *)obs_tape, reward_sum);
    return check_launch("tape_f64_kernel");
}

int cstr_tape_f32_host(const cstr_env_params *p, int64_t n, int64_t T, int math_mode, const float *h_actions, float *h_state,
                       int32_t *h_step_count, int32_t *h_episode, float *h_rewards, uint8_t *h_dones, void *stream) {
    if (!p || n < 0 || T < 0 || !h_actions || !h_state || !h_step_count || !h_episode)
        return fail_arg(CSTR_EINVAL, "null pointer or negative size");
    if (p->init_mode != CSTR_INIT_RANDOM) return fail_arg(CSTR_EINVAL, "host tape supports init_mode=random only");
    if (n == 0 || T == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // scratch layout: state | step_count | episode | actions | rewards | dones  (256 B aligned sections)
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_state = 0, o_sc = up(o_state + (size_t)n * 16), o_ep = up(o_sc + (size_t)n * 4), o_act = up(o_ep + (size_t)n * 4);
    const size_t o_rew = up(o_act + (size_t)T * n * 8), o_done = up(o_rew + (size_t)T * n * 4), total = up(o_done + (size_t)T * n);
    if (int rc = g_scratch.ensure(total)) return rc;
    char *d = (char *)g_scratch.buf;
    int rc;
    if ((rc = check_cuda(cudaMemcpyAsync(d + o_state, h_state, (size_t)n * 16, cudaMemcpyHostToDevice, st), "H2D state"))) return rc;
    if ((rc = check_cuda(cudaMemcpyAsync(d + o_sc, h_step_count, (size_t)n * 4, cudaMemcpyHostToDevice, st), "H2D step_count"))) return rc;
    if ((rc = check_cuda(cudaMemcpyAsync(d + o_ep, h_episode, (size_t)n * 4, cudaMemcpyHostToDevice, st), "H2D episode"))) return rc;
    if ((rc = check_cuda(cudaMemcpyAsync(d + o_act, h_actions, (size_t)T * n * 8, cudaMemcpyHostToDevice, st), "H2D actions"))) return rc;
    rc = cstr_tape_f32(p, n, T, math_mode, (const float *)(d + o_act), 0u, (float *)(d + o_state), (int32_t *)(d + o_sc),
                       (int32_t *)(d + o_ep), nullptr, h_rewards ? (float *)(d + o_rew) : nullptr,
                       h_dones ? (uint8_t *)(d + o_done) : nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    if ((rc = check_cuda(cudaMemcpyAsync(h_state, d + o_state, (size_t)n * 16, cudaMemcpyDeviceToHost, st), "D2H state"))) return rc;
    if ((rc = check_cuda(cudaMemcpyAsync(h_step_count, d + o_sc, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H step_count"))) return rc;
    if ((rc = check_cuda(cudaMemcpyAsync(h_episode, d + o_ep, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H episode"))) return rc;
    if (h_rewards && (rc = check_cuda(cudaMemcpyAsync(h_rewards, d + o_rew, (size_t)T * n * 4, cudaMemcpyDeviceToHost, st), "D2H rewards"))) return rc;
    if (h_dones && (rc = check_cuda(cudaMemcpyAsync(h_dones, d + o_done, (size_t)T * n, cudaMemcpyDeviceToHost, st), "D2H dones"))) return rc;
    return check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
}

}  // extern "C"
