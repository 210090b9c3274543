// Shape-generic pieces for the update steps whose networks are not TD3's fixed (4|6 -> H1 -> H2 -> 1|2) tables: the BCQ VAE / decoder /
// perturbation nets (core/bcq/policies.py:21-166) and the per-agent actors and critics of MADDPG / IDDPG (core/maddpg/policies.py:21-121,
// core/iddpg/policies.py).  Included by cstr_td3.cu after its kernels: the hidden-layer GEMM (FFMA or tcgen05), the skinny weight-gradient
// kernels, the finalize and the Adam/polyak apply kernels are already shape-generic and are reused as they are; what is added here is
//   * a layer-1 kernel with a run-time input width fed from up to two row-major sources (optionally a source whose rows repeat with a
//     period: the reference's `state.repeat(n, 1)` without materialising the tiled batch),
//   * a linear head (forward, optional tanh) and its backward into the second hidden layer,
//   * the input gradient of layer 1 for a column range (what flows on into the network that produced those inputs),
//   * host helpers forward_mlp / backward_mlp that chain them the way forward_hidden / backward_hidden do for TD3.
// Arithmetic conventions are those of the TD3 kernels (bias first, k ascending FMA; warp row-dots), so a net evaluated through either
// path gives the same bits.
#pragma once

namespace cstr {

struct Src {  // layer-1 input of row b for net z = [(x0 + z*x0_z)[(b % x0_rows)][0:n0] | (x1 + z*x1_z)[b][0:n1]]  (x0_rows = 0: no repetition)
    const float *x0;
    int n0, ld0, x0_rows;
    const float *x1;
    int n1, ld1;
    int x0_z, x1_z;  // per-net offsets in floats: z-batched nets that read different column slices of the same rows (per-agent observations)
};

constexpr int L1G_ROWS = 8;

__global__ void __launch_bounds__(256)
mlp_layer1_kernel(int B, int H1, Src s, const float *__restrict__ W1, const float *__restrict__ b1, int64_t w_stride_z, float *__restrict__ h1,
                  int64_t h_stride_z) {
    pdl_enter();
    const int q = H1 >> 2, IN = s.n0 + s.n1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int row_groups = (B + L1G_ROWS - 1) / L1G_ROWS;
    if (i >= (int64_t)row_groups * q) return;
    const int b0 = (int)(i / q) * L1G_ROWS, j = (int)(i % q) * 4, z = blockIdx.y;
    const float *W = W1 + z * w_stride_z + (int64_t)j * IN;
    const float4 bias = *reinterpret_cast<const float4 *>(b1 + z * w_stride_z + j);
    float acc[L1G_ROWS][4];
    const float *r0[L1G_ROWS], *r1[L1G_ROWS];
#pragma unroll
    for (int r = 0; r < L1G_ROWS; ++r) {
        acc[r][0] = bias.x, acc[r][1] = bias.y, acc[r][2] = bias.z, acc[r][3] = bias.w;
        const int b = min(b0 + r, B - 1);  // rows past the end recompute the last row and are not stored
        r0[r] = s.x0 + z * s.x0_z + (int64_t)(s.x0_rows ? b % s.x0_rows : b) * s.ld0;
        r1[r] = s.x1 ? s.x1 + z * s.x1_z + (int64_t)b * s.ld1 : nullptr;
    }
    for (int k = 0; k < IN; ++k) {
        const float w0 = __ldg(W + k), w1 = __ldg(W + IN + k), w2 = __ldg(W + 2 * IN + k), w3 = __ldg(W + 3 * IN + k);
#pragma unroll
        for (int r = 0; r < L1G_ROWS; ++r) {
            const float x = k < s.n0 ? r0[r][k] : r1[r][k - s.n0];
            acc[r][0] = fmaf(x, w0, acc[r][0]), acc[r][1] = fmaf(x, w1, acc[r][1]);
            acc[r][2] = fmaf(x, w2, acc[r][2]), acc[r][3] = fmaf(x, w3, acc[r][3]);
        }
    }
    float *out = h1 + z * h_stride_z + j;
#pragma unroll
    for (int r = 0; r < L1G_ROWS; ++r) {
        const int b = b0 + r;
        if (b >= B) break;
        *reinterpret_cast<float4 *>(out + (int64_t)b * H1) =
            make_float4(fmaxf(acc[r][0], 0.f), fmaxf(acc[r][1], 0.f), fmaxf(acc[r][2], 0.f), fmaxf(acc[r][3], 0.f));
    }
}

// y[z][b][o] = h2[z][b] . W3[z][o] + b3[z][o]   (tanh when squash)                 one warp per row, OUT dot products
__global__ void __launch_bounds__(256)
mlp_head_fwd_kernel(int B, int H2, int OUT, const float *__restrict__ h2, int64_t h_z, const float *__restrict__ W3, const float *__restrict__ b3,
                    int64_t w_z, float *__restrict__ y, int squash) {
    pdl_enter();
    const int lane = threadIdx.x & 31, b = blockIdx.x * 8 + (threadIdx.x >> 5), z = blockIdx.y;
    if (b >= B) return;
    const float *h = h2 + z * h_z + (int64_t)b * H2, *w = W3 + z * w_z, *bb = b3 + z * w_z;
    float *out = y + ((int64_t)z * B + b) * OUT;
    for (int o = 0; o < OUT; ++o) {
        const float v = row_dot(h, w + (int64_t)o * H2, H2, lane) + bb[o];
        if (lane == 0) out[o] = squash ? tanhf(v) : v;
    }
}

// dz2[b][k] = (sum_o dy[b][o] W3[o][k]) * (h2[b][k] > 0)
__global__ void __launch_bounds__(256)
mlp_head_bwd_kernel(int B, int H2, int OUT, const float *__restrict__ dy, const float *__restrict__ W3, const float *__restrict__ h2,
                    float *__restrict__ dz2) {
    pdl_enter();
    const int lane = threadIdx.x & 31, b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float *d = dy + (int64_t)b * OUT, *h = h2 + (int64_t)b * H2;
    float *o = dz2 + (int64_t)b * H2;
    for (int k = lane * 4; k < H2; k += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < OUT; ++r) {
            const float p = d[r];
            const float4 w = *reinterpret_cast<const float4 *>(W3 + (int64_t)r * H2 + k);
            acc.x = fmaf(p, w.x, acc.x), acc.y = fmaf(p, w.y, acc.y), acc.z = fmaf(p, w.z, acc.z), acc.w = fmaf(p, w.w, acc.w);
        }
        const float4 hh = *reinterpret_cast<const float4 *>(h + k);
        *reinterpret_cast<float4 *>(o + k) = make_float4(hh.x > 0.f ? acc.x : 0.f, hh.y > 0.f ? acc.y : 0.f, hh.z > 0.f ? acc.z : 0.f, hh.w > 0.f ? acc.w : 0.f);
    }
}

// dx[b][c] = sum_j dz1[b][j] * W1[j][col0 + c]   for c < ncols                      (the gradient w.r.t. a slice of the layer-1 input)
__global__ void __launch_bounds__(256)
mlp_dx_kernel(int B, int H1, int IN, const float *__restrict__ dz1, const float *__restrict__ W1, int col0, int ncols, float *__restrict__ dx) {
    pdl_enter();
    const int lane = threadIdx.x & 31, b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float *d = dz1 + (int64_t)b * H1;
    for (int c = 0; c < ncols; ++c) {
        float s = 0.f;
        for (int j = lane; j < H1; j += 32) s = fmaf(d[j], __ldg(W1 + (int64_t)j * IN + col0 + c), s);
        s = warp_sum(s);
        if (lane == 0) dx[(int64_t)b * ncols + c] = s;
    }
}

// 4 standard normals from one Philox block (two Box-Muller pairs, the recipe of td3_actor_head_kernel)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t row, uint32_t c2, uint32_t stream, uint32_t call) {
    const uint4 r = philox_env(seed, row, c2, stream, call);
    const float u1 = fmaf(u24(r.x), 1.0f, 5.9604644775390625e-08f), u2 = u24(r.y);
    const float u3 = fmaf(u24(r.z), 1.0f, 5.9604644775390625e-08f), u4 = u24(r.w);
    const float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
    float s1, c1, s2, c2f;
    sincospif(2.0f * u2, &s1, &c1);
    sincospif(2.0f * u4, &s2, &c2f);
    return make_float4(ra * c1, ra * s1, rb * c2f, rb * s2);
}

}  // namespace cstr
