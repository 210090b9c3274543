// Element-wise / row kernels specific to the BCQ update (core/bcq/bcq.py:129-213, core/bcq/policies.py:21-166) and to the multi-agent
// DDPG updates (core/maddpg/maddpg.py:117-191, core/iddpg/iddpg.py:117-191).  The matrix work runs on the shared kernels (cstr_td3.cu,
// cstr_mlp.cuh); these are the pieces between them.  Restated from oracle/td3_oracle.py::BCQUpdateOracle / MultiAgentDDPGOracle.
#pragma once

namespace cstr {

enum : uint32_t { STREAM_BCQ_EPS = 6u, STREAM_BCQ_NEXT = 7u, STREAM_BCQ_ACTOR = 8u, STREAM_MA = 9u };
constexpr float BCQ_LOG_STD_MIN = -4.f, BCQ_LOG_STD_MAX = 15.f, BCQ_Z_CLIP = 0.5f;

// VAE latent (policies.py:70-79): [mean | raw] = encoder head (B, 2L); std = exp(clamp(raw, -4, 15)); z = mean + std * eps;
// writes std, the eps used, and the decoder input xdec = [obs | z]  (B, 4 + L).  One thread per (row, 4 latent dims).
__global__ void __launch_bounds__(256)
bcq_latent_kernel(int B, int L, const float *__restrict__ enc_y, const float *__restrict__ eps_in, uint64_t seed, uint32_t update_index,
                  const float *__restrict__ dev_scalars, const float4 *__restrict__ obs, float *__restrict__ std_out, float *__restrict__ eps_out,
                  float *__restrict__ xdec) {
    pdl_enter();
    const int q = L >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * q) return;
    const int b = (int)(i / q), l = (int)(i % q) * 4;
    const float *y = enc_y + (int64_t)b * 2 * L;
    const float4 mean = *reinterpret_cast<const float4 *>(y + l), raw = *reinterpret_cast<const float4 *>(y + L + l);
    float4 e;
    if (eps_in) e = *reinterpret_cast<const float4 *>(eps_in + (int64_t)b * L + l);
    else e = philox_normal4(seed, (uint64_t)b, dev_scalars ? __float_as_uint(dev_scalars[4]) : update_index, STREAM_BCQ_EPS, (uint32_t)(l >> 2));
    const float m[4] = {mean.x, mean.y, mean.z, mean.w}, r[4] = {raw.x, raw.y, raw.z, raw.w}, ev[4] = {e.x, e.y, e.z, e.w};
    float sd[4], z[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        sd[c] = expf(fminf(fmaxf(r[c], BCQ_LOG_STD_MIN), BCQ_LOG_STD_MAX));
        z[c] = __fadd_rn(m[c], __fmul_rn(sd[c], ev[c]));
    }
    *reinterpret_cast<float4 *>(std_out + (int64_t)b * L + l) = make_float4(sd[0], sd[1], sd[2], sd[3]);
    *reinterpret_cast<float4 *>(eps_out + (int64_t)b * L + l) = e;
    float *x = xdec + (int64_t)b * (4 + L);
    *reinterpret_cast<float4 *>(x + 4 + l) = make_float4(z[0], z[1], z[2], z[3]);
    if (l == 0) *reinterpret_cast<float4 *>(x) = obs[b];
}

// VAE loss (bcq.py:146-151) and the gradient entering the decoder's tanh head:
//   loss = mse(recon, a) + 0.5 * KL,  KL = -0.5 * mean(1 + log(std^2) - mean^2 - std^2);   d_pre = (2 / (2B)) (recon - a) (1 - recon^2)
__global__ void __launch_bounds__(256)
bcq_vae_loss_kernel(int B, int L, const float2 *__restrict__ recon, const float2 *__restrict__ act, const float *__restrict__ enc_y,
                    const float *__restrict__ std_in, float2 *__restrict__ d_pre, float *__restrict__ loss_partial) {
    pdl_enter();
    __shared__ float sl[256];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float contrib = 0.f;
    if (b < B) {
        const float2 r = recon[b], a = act[b];
        const float dx = r.x - a.x, dy = r.y - a.y, g = 2.f / (2.f * (float)B);
        d_pre[b] = make_float2(g * dx * (1.f - r.x * r.x), g * dy * (1.f - r.y * r.y));
        float kl = 0.f;
        const float *y = enc_y + (int64_t)b * 2 * L, *sd = std_in + (int64_t)b * L;
        for (int l = 0; l < L; ++l) {
            const float s = sd[l], m = y[l];
            kl += 1.f + logf(s * s) - m * m - s * s;
        }
        contrib = (dx * dx + dy * dy) / (2.f * (float)B) + 0.5f * (-0.5f) * kl / ((float)B * (float)L);
    }
    sl[threadIdx.x] = contrib;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sl[threadIdx.x] += sl[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_partial[blockIdx.x] = sl[0];
}

// gradient entering the encoder head, [dmean | draw] (B, 2L), from dz (through the decoder's input layer) and the KL term
__global__ void __launch_bounds__(256)
bcq_enc_grad_kernel(int B, int L, const float *__restrict__ dz, const float *__restrict__ enc_y, const float *__restrict__ std_in,
                    const float *__restrict__ eps, float *__restrict__ dy) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * L) return;
    const int b = (int)(i / L), l = (int)(i % L);
    const float n = (float)B * (float)L;
    const float mean = enc_y[(int64_t)b * 2 * L + l], raw = enc_y[(int64_t)b * 2 * L + L + l], sd = std_in[i], g = dz[i];
    const float dmean = g + 0.5f * mean / n;
    const float dstd = g * eps[i] + 0.5f * (sd - 1.f / sd) / n;
    const float draw = (raw >= BCQ_LOG_STD_MIN && raw <= BCQ_LOG_STD_MAX) ? dstd * sd : 0.f;  // clamp passes the gradient inside (and at) the bounds
    dy[(int64_t)b * 2 * L + l] = dmean;
    dy[(int64_t)b * 2 * L + L + l] = draw;
}

// latent draws of decode()/sample_action (policies.py:111,122): clip(N(0,1), -0.5, 0.5), (rows, L)
__global__ void __launch_bounds__(256)
bcq_clip_latent_kernel(int64_t rows, int L, const float *__restrict__ z_in, uint64_t seed, uint32_t update_index, const float *__restrict__ dev_scalars,
                       uint32_t stream, float *__restrict__ zc) {
    pdl_enter();
    const int q = L >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * q) return;
    const int64_t r = i / q;
    const int l = (int)(i % q) * 4;
    float4 z;
    if (z_in) z = *reinterpret_cast<const float4 *>(z_in + r * L + l);
    else z = philox_normal4(seed, (uint64_t)r, dev_scalars ? __float_as_uint(dev_scalars[4]) : update_index, stream, (uint32_t)(l >> 2));
    *reinterpret_cast<float4 *>(zc + r * L + l) = make_float4(fminf(fmaxf(z.x, -BCQ_Z_CLIP), BCQ_Z_CLIP), fminf(fmaxf(z.y, -BCQ_Z_CLIP), BCQ_Z_CLIP),
                                                              fminf(fmaxf(z.z, -BCQ_Z_CLIP), BCQ_Z_CLIP), fminf(fmaxf(z.w, -BCQ_Z_CLIP), BCQ_Z_CLIP));
}

// perturbation (policies.py:152-160): pre = a + phi * xi, out = clamp(pre, -1, 1)
__global__ void __launch_bounds__(256)
bcq_perturb_kernel(int64_t rows, const float2 *__restrict__ a, const float2 *__restrict__ xi, float phi, float2 *__restrict__ out, float2 *__restrict__ pre_out) {
    pdl_enter();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float2 av = a[r], x = xi[r];
    const float p0 = __fadd_rn(av.x, __fmul_rn(x.x, phi)), p1 = __fadd_rn(av.y, __fmul_rn(x.y, phi));
    out[r] = make_float2(fminf(fmaxf(p0, -1.f), 1.f), fminf(fmaxf(p1, -1.f), 1.f));
    if (pre_out) pre_out[r] = make_float2(p0, p1);
}

// target (bcq.py:165-172): q = min over the critics of the (K*B, 1) column, reshaped to (B, K) ROW-MAJOR as the reference writes it —
// so row b takes the max over the K CONSECUTIVE entries b*K .. b*K+K-1 of the tiled batch — then r + (1 - done) * gamma * max
__global__ void __launch_bounds__(256)
bcq_target_kernel(int B, int K, int n_critics, const float *__restrict__ q, const float *__restrict__ rewards, const float *__restrict__ dones, float gamma,
                  float *__restrict__ target) {
    pdl_enter();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t R = (int64_t)B * K;
    float best = -INFINITY;
    for (int j = 0; j < K; ++j) {
        const int64_t r = (int64_t)b * K + j;
        float v = q[r];
        if (n_critics > 1) v = fminf(v, q[R + r]);
        best = fmaxf(best, v);
    }
    target[b] = rewards[b] + (1.f - dones[b]) * gamma * best;
}

// gradient entering the perturbation net's tanh head: da (dQ1/da) gated by the clamp, times phi, times tanh'
__global__ void __launch_bounds__(256)
bcq_pert_grad_kernel(int B, const float2 *__restrict__ da, const float2 *__restrict__ pre, const float2 *__restrict__ xi, float phi, float2 *__restrict__ dy) {
    pdl_enter();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float2 g = da[b], p = pre[b], x = xi[b];
    const float g0 = (p.x >= -1.f && p.x <= 1.f) ? g.x : 0.f, g1 = (p.y >= -1.f && p.y <= 1.f) ? g.y : 0.f;
    dy[b] = make_float2(g0 * phi * (1.f - x.x * x.x), g1 * phi * (1.f - x.y * x.y));
}

// ---- multi-agent DDPG (two reactors = two agents; one action dimension per agent) ----------------------------------------------
// next_actions[:, i] = clip(tanh-head_i + clip(noise_i, -c, c), -1, 1)   (maddpg.py:132-144); y = (n_agents, B) head outputs
__global__ void __launch_bounds__(256)
ma_next_action_kernel(int B, int n_agents, const float *__restrict__ y, const float *__restrict__ noise, float sigma, float clip, uint64_t seed,
                      uint32_t update_index, const float *__restrict__ dev_scalars, float *__restrict__ next_act) {
    pdl_enter();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!noise) {
        nz = philox_normal4(seed, (uint64_t)b, dev_scalars ? __float_as_uint(dev_scalars[4]) : update_index, STREAM_MA, 0);
        nz.x *= sigma, nz.y *= sigma, nz.z *= sigma, nz.w *= sigma;
    }
    const float nv[4] = {nz.x, nz.y, nz.z, nz.w};
    for (int i = 0; i < n_agents; ++i) {
        const float n = noise ? noise[(int64_t)i * B + b] : nv[i & 3];
        next_act[(int64_t)b * n_agents + i] = fminf(fmaxf(y[(int64_t)i * B + b] + fminf(fmaxf(n, -clip), clip), -1.f), 1.f);
    }
}

// joint[b][j] = y[j][b]: the (B, n_agents) action matrix a critic reads, from the z-batched actor heads
__global__ void __launch_bounds__(256) ma_joint_kernel(int B, int n_agents, const float *__restrict__ y, float *__restrict__ joint) {
    pdl_enter();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int j = 0; j < n_agents; ++j) joint[(int64_t)b * n_agents + j] = y[(int64_t)j * B + b];
}

// dy[b] = da[b] * (1 - a[b]^2): into agent i's tanh head
__global__ void __launch_bounds__(256) ma_actor_grad_kernel(int B, const float *__restrict__ da, const float *__restrict__ a, float *__restrict__ dy) {
    pdl_enter();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float v = a[b];
    dy[b] = da[b] * (1.f - v * v);
}

}  // namespace cstr
