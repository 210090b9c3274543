// Tensor-core version of the TD3/SAC hidden-layer GEMM (same three roles and epilogues as td3_gemm_kernel): tcgen05.mma
// (kind::f16, bf16 operands, fp32 accumulators in TMEM) with each fp32 operand split into three bf16 planes by truncation
//     x = x1 + x2 + x3,   x1 = hi16(x), x2 = hi16(x - x1), x3 = hi16(x - x1 - x2)          (24 mantissa bits in total, exact residuals)
// and six products per K-step  a1b1 + a1b2 + a2b1 + a2b2 + a1b3 + a3b1  (the dropped terms are <= 2^-24 |a||b|), i.e.
// float32-grade accuracy (measured 3e-6 of the max against the FFMA path; the tensor core's accumulation truncates).
//
// CTA = 128 output rows x n_tile (<= 256) output columns; 16 producer warps + 1 issuer warp.  The producers load fp32 from
// global (either operand may be K-contiguous or MN-contiguous in memory — the transposes of the dgrad / wgrad roles happen
// here), split, and store the planes into shared memory in the UMMA no-swizzle K-major core-matrix layout
// ([k-group of 8][row-group of 8][8 rows x 16 B]); they arrive on the stage's `full` mbarrier after a proxy fence.  Lane 0 of
// the issuer warp waits for `full`, issues the six MMAs of the stage and commits them to the stage's `empty` mbarrier, which
// frees the stage for reuse (3 stages of K=32, global loads TD3_TC_DEPTH stages ahead in registers).  Epilogue: tcgen05.ld ->
// bias+relu / relu-mask / slab store, 32-byte row pieces straight to global.
// Build-time switches: -DTD3_TC_ACC=1|3 (one accumulator, or one per product pair summed in the epilogue: n_tile <= 160 then),
// -DTD3_TC_DEPTH=2|3 (register prefetch depth).  Measured within 10 % of each other; defaults 1 / 2.
#pragma once
#include <cuda_bf16.h>

namespace cstr {
namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TC5_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TC5_WAIT_DONE;\n\t"
        "bra TC5_WAIT_LOOP;\n\t"
        "TC5_WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, SWIZZLE_NONE shared-memory descriptor (see cstr_rollout_tc.cu): LBO between the two 8-element K halves, SBO between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ inline uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// {lo -> bits 0..15, hi -> bits 16..31}, round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// split 8 consecutive-k floats into the three bf16 planes and store one 16-byte core-matrix row per plane.
// Planes by truncation (x1 = x & 0xffff0000, r = x - x1 exact, ...): three planes still carry 24 mantissa bits, and the split is
// LOP/FADD/PRMT only (cvt.rn.bf16x2 runs on the quarter-rate conversion pipe and was the producer bottleneck).
template <int PLANES>
__device__ __forceinline__ void split_store(const float v[8], uint32_t addr, uint32_t plane_stride) {
    if constexpr (PLANES == 1) {  // plain bf16 operands (round to nearest): the reduced-precision throughput mode
        uint32_t q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]) : "memory");
    } else {
    uint32_t p1[4], p2[4], p3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t a = __float_as_uint(v[2 * i]), b = __float_as_uint(v[2 * i + 1]);
        p1[i] = __byte_perm(a, b, 0x7632);  // {hi16(a) -> bits 0..15, hi16(b) -> bits 16..31}
        const float ra = v[2 * i] - __uint_as_float(a & 0xffff0000u), rb = v[2 * i + 1] - __uint_as_float(b & 0xffff0000u);
        const uint32_t a2 = __float_as_uint(ra), b2 = __float_as_uint(rb);
        p2[i] = __byte_perm(a2, b2, 0x7632);
        const float sa = ra - __uint_as_float(a2 & 0xffff0000u), sb = rb - __uint_as_float(b2 & 0xffff0000u);
        p3[i] = __byte_perm(__float_as_uint(sa), __float_as_uint(sb), 0x7632);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(p1[0]), "r"(p1[1]), "r"(p1[2]), "r"(p1[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + plane_stride), "r"(p2[0]), "r"(p2[1]), "r"(p2[2]), "r"(p2[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 2 * plane_stride), "r"(p3[0]), "r"(p3[1]), "r"(p3[2]), "r"(p3[3]) : "memory");
    }
}

// the same for 4 consecutive-k floats (half a core-matrix row, 8 bytes per plane)
template <int PLANES>
__device__ __forceinline__ void split_store4(const float v[4], uint32_t addr, uint32_t plane_stride) {
    if constexpr (PLANES == 1) {
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(pack_bf16x2(v[0], v[1])), "r"(pack_bf16x2(v[2], v[3])) : "memory");
    } else {
        uint32_t p1[2], p2[2], p3[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t a = __float_as_uint(v[2 * i]), b = __float_as_uint(v[2 * i + 1]);
            p1[i] = __byte_perm(a, b, 0x7632);
            const float ra = v[2 * i] - __uint_as_float(a & 0xffff0000u), rb = v[2 * i + 1] - __uint_as_float(b & 0xffff0000u);
            const uint32_t a2 = __float_as_uint(ra), b2 = __float_as_uint(rb);
            p2[i] = __byte_perm(a2, b2, 0x7632);
            const float sa = ra - __uint_as_float(a2 & 0xffff0000u), sb = rb - __uint_as_float(b2 & 0xffff0000u);
            p3[i] = __byte_perm(__float_as_uint(sa), __float_as_uint(sb), 0x7632);
        }
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(p1[0]), "r"(p1[1]) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + plane_stride), "r"(p2[0]), "r"(p2[1]) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + 2 * plane_stride), "r"(p3[0]), "r"(p3[1]) : "memory");
    }
}

}  // namespace tc5

#ifndef TD3_TC_PRODUCERS
#define TD3_TC_PRODUCERS 512
#endif
#ifndef TD3_TC_ACC
#define TD3_TC_ACC 1
#endif
#ifndef TD3_TC_DEPTH
#define TD3_TC_DEPTH 2
#endif
#ifdef CSTR_TD3_TIMING
__device__ __forceinline__ long long tc5_clock() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
#define TC5_T(var) const long long var = tc5_clock();
#define TC5_ACC(i, a, b) tacc[i] += (b) - (a);
#else
#define TC5_T(var)
#define TC5_ACC(i, a, b)
#endif
constexpr int TC_BM = 128, TC_KC = 32, TC_KG = TC_KC / 8, TC_STAGES = 3, TC_PRODUCERS = 512, TC_NACC = TD3_TC_ACC, TC_DEPTH = TD3_TC_DEPTH,
              TC_GEMM_THREADS = TC_PRODUCERS + 32;

// One producer item.  KMAJ operand (memory rows = operand rows, k contiguous): 4 consecutive k of one row = one float4; consecutive
// threads take consecutive quarters of the SAME row, so a warp reads four whole 128-byte lines (the first version gave every thread
// its own row: 32 lines per warp-load, and the L1 tag stage — not the MMAs, not the split — bounded the kernel).
// MN operand (memory rows = k, operand rows contiguous): 8 consecutive k of one operand row, consecutive threads = consecutive rows.
template <bool KMAJ>
__device__ __forceinline__ void load_item(const float *__restrict__ p, int64_t ld, bool ok, int k_left, float *v) {
    if (KMAJ) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && k_left > 0) a = *reinterpret_cast<const float4 *>(p);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (ok && j < k_left) ? p[j * ld] : 0.f;
    }
}

template <bool KMAJ, int PLANES>
__device__ __forceinline__ void store_item(const float *v, uint32_t addr, uint32_t plane_stride) {
    if (KMAJ) tc5::split_store4<PLANES>(v, addr, plane_stride);
    else tc5::split_store<PLANES>(v, addr, plane_stride);
}

template <int MODE, int PLANES = 3>  // PLANES = 3: fp32-grade split; 1: plain bf16 operands
__global__ void __launch_bounds__(TC_GEMM_THREADS) td3_gemm_tc_kernel(GemmArgs g, int n_tile, int tmem_cols) {
    using namespace tc5;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr bool A_KMAJ = MODE != G_WGRAD, B_KMAJ = MODE == G_FWD;
    constexpr int A_ITEMS = A_KMAJ ? 2 : 1, A_W = A_KMAJ ? 4 : 8;  // 128 rows: 1024 quarters or 512 chunks over 512 producers
    constexpr int B_ITEMS = B_KMAJ ? 4 : 2, B_W = B_KMAJ ? 4 : 8;  // n_tile <= 256 rows
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bars = smem_base;  // empty[TC_STAGES], full[TC_STAGES]
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + 96);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z / g.splits, split = blockIdx.z % g.splits;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * n_tile;
    const float *A = g.A + z * g.a_z, *Bm = g.Bm + z * g.b_z;
    int k_begin = 0, k_end = g.K;
    if (MODE == G_WGRAD) {
        k_begin = split * g.k_per_split;
        k_end = min(g.K, k_begin + g.k_per_split);
    }
    const int n_iter = max(0, (k_end - k_begin + TC_KC - 1) / TC_KC);
    // one stage: A planes 3 x [4 k-groups][16 row-groups][128 B], then B planes 3 x [4][n_tile/8][128 B]
    const uint32_t lbo_a = (TC_BM / 8) * 128u, lbo_b = (uint32_t)(n_tile / 8) * 128u;
    const uint32_t a_plane = TC_KG * lbo_a, b_plane = TC_KG * lbo_b;
    const uint32_t stage_bytes = 3u * (a_plane + b_plane), stage0 = smem_base + 128u;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);                            // empty[s]: one tcgen05.commit
            mbar_init(bars + 8 * (TC_STAGES + s), TC_PRODUCERS);   // full[s]: every producer thread
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)), "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_enter();  // barrier init and the TMEM allocation above overlap the tail of the previous kernel; operands are read below
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = umma_idesc_bf16(TC_BM, n_tile);

    if (warp == TC_PRODUCERS / 32) {
        // ---- issuer warp: wait for a full stage, issue its MMAs (two K=16 steps), commit to the stage's empty barrier ----
        if (lane == 0) {
            for (int it = 0; it < n_iter; ++it) {
                const int s = it % TC_STAGES, use = it / TC_STAGES;
                mbar_wait(bars + 8 * (TC_STAGES + s), (uint32_t)(use & 1));
                tc_fence_after();
                const uint32_t sa = stage0 + (uint32_t)s * stage_bytes, sb = sa + 3u * a_plane;
                const uint32_t d0 = tmem_base, d1 = TC_NACC == 3 ? tmem_base + (uint32_t)n_tile : d0, d2 = TC_NACC == 3 ? tmem_base + 2u * (uint32_t)n_tile : d0;
#pragma unroll
                for (int ks = 0; ks < TC_KC / 16; ++ks) {
                    uint64_t da[3], db[3];
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        da[p] = umma_desc(sa + p * a_plane + ks * 2u * lbo_a, lbo_a, 128u);
                        db[p] = umma_desc(sb + p * b_plane + ks * 2u * lbo_b, lbo_b, 128u);
                    }
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    tc_mma_bf16(d0, da[0], db[0], idesc, acc);
                    if (PLANES == 3) {
                        tc_mma_bf16(d1, da[0], db[1], idesc, TC_NACC == 3 ? acc : 1u);
                        tc_mma_bf16(d2, da[1], db[0], idesc, TC_NACC == 3 ? acc : 1u);
                        tc_mma_bf16(d0, da[1], db[1], idesc, 1);
                        tc_mma_bf16(d1, da[2], db[0], idesc, 1);
                        tc_mma_bf16(d2, da[0], db[2], idesc, 1);
                    }
                }
                tc_commit(bars + 8 * s);
            }
        }
    } else {
        // ---- producers: everything that does not change from stage to stage is computed once per item ----
        const int64_t lda = g.lda, ldb = g.ldb;
        const float *a_ptr[A_ITEMS], *b_ptr[B_ITEMS];
        uint32_t a_off[A_ITEMS], b_off[B_ITEMS];
        bool a_ok[A_ITEMS], b_ok[B_ITEMS], b_live[B_ITEMS];
        int a_k[A_ITEMS], b_k[B_ITEMS];
#pragma unroll
        for (int i = 0; i < A_ITEMS; ++i) {
            const int id = tid + i * TC_PRODUCERS;
            if (A_KMAJ) {
                const int r = id >> 3, kq = id & 7;
                a_ok[i] = m0 + r < g.M, a_k[i] = kq * 4;
                a_off[i] = (uint32_t)(kq >> 1) * lbo_a + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(kq & 1) * 8u;
                a_ptr[i] = A + (int64_t)(m0 + r) * lda + (k_begin + kq * 4);
            } else {
                const int r = id & (TC_BM - 1), kg = id >> 7;
                a_ok[i] = m0 + r < g.M, a_k[i] = kg * 8;
                a_off[i] = (uint32_t)kg * lbo_a + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
                a_ptr[i] = A + (int64_t)(k_begin + kg * 8) * lda + (m0 + r);
            }
        }
#pragma unroll
        for (int i = 0; i < B_ITEMS; ++i) {
            const int id = tid + i * TC_PRODUCERS;
            if (B_KMAJ) {
                const int r = id >> 3, kq = id & 7;
                b_live[i] = r < n_tile, b_ok[i] = b_live[i] && n0 + r < g.N, b_k[i] = kq * 4;
                b_off[i] = 3u * a_plane + (uint32_t)(kq >> 1) * lbo_b + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(kq & 1) * 8u;
                b_ptr[i] = Bm + (int64_t)(n0 + r) * ldb + (k_begin + kq * 4);
            } else {
                const int r = id % n_tile, kg = id / n_tile;
                b_live[i] = kg < TC_KG, b_ok[i] = b_live[i] && n0 + r < g.N, b_k[i] = kg * 8;
                b_off[i] = 3u * a_plane + (uint32_t)kg * lbo_b + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
                b_ptr[i] = Bm + (int64_t)(k_begin + kg * 8) * ldb + (n0 + r);
            }
        }
        const int64_t a_step = A_KMAJ ? TC_KC : TC_KC * lda, b_step = B_KMAJ ? TC_KC : TC_KC * ldb;
        float va[TC_DEPTH][A_ITEMS][A_W], vb[TC_DEPTH][B_ITEMS][B_W];  // register sets: loads of stage it+TC_DEPTH are issued once stage it is written
        auto load_stage = [&](int it, float (&xa)[A_ITEMS][A_W], float (&xb)[B_ITEMS][B_W]) {
            const int k_left = k_end - (k_begin + it * TC_KC);
#pragma unroll
            for (int i = 0; i < A_ITEMS; ++i) load_item<A_KMAJ>(a_ptr[i] + it * a_step, lda, a_ok[i], k_left - a_k[i], xa[i]);
#pragma unroll
            for (int i = 0; i < B_ITEMS; ++i)
                if (b_live[i]) load_item<B_KMAJ>(b_ptr[i] + it * b_step, ldb, b_ok[i], k_left - b_k[i], xb[i]);
        };
        auto body = [&](int it, float (&xa)[A_ITEMS][A_W], float (&xb)[B_ITEMS][B_W]) {
            const int s = it % TC_STAGES, use = it / TC_STAGES;
            if (use > 0) {  // the MMAs that read this stage have finished; one lane per warp polls
                if (lane == 0) mbar_wait(bars + 8 * s, (uint32_t)((use - 1) & 1));
                __syncwarp();
            }
            const uint32_t st = stage0 + (uint32_t)s * stage_bytes;
#pragma unroll
            for (int i = 0; i < A_ITEMS; ++i) store_item<A_KMAJ, PLANES>(xa[i], st + a_off[i], a_plane);
#pragma unroll
            for (int i = 0; i < B_ITEMS; ++i)
                if (b_live[i]) store_item<B_KMAJ, PLANES>(xb[i], st + b_off[i], b_plane);
            if (it + TC_DEPTH < n_iter) load_stage(it + TC_DEPTH, xa, xb);
            fence_proxy_async();
            mbar_arrive(bars + 8 * (TC_STAGES + s));
        };
#pragma unroll
        for (int d = 0; d < TC_DEPTH; ++d)
            if (d < n_iter) load_stage(d, va[d], vb[d]);
        for (int it = 0; it < n_iter; it += TC_DEPTH) {
#pragma unroll
            for (int d = 0; d < TC_DEPTH; ++d)
                if (it + d < n_iter) body(it + d, va[d], vb[d]);
        }
    }
    if (n_iter > 0) {
        const int last = n_iter - 1;
        mbar_wait(bars + 8 * (last % TC_STAGES), (uint32_t)((last / TC_STAGES) & 1));
    }
    tc_fence_after();

    // ---- epilogue: thread = output row (TMEM lane); the four producer warps of a lane quarter take the 8-column groups
    // round-robin, two groups per batch so that the TMEM loads and the mask loads of a batch are all in flight together
    if (warp < TC_PRODUCERS / 32) {
        const int row = m0 + (warp & 3) * 32 + lane;
        const bool row_ok = row < g.M;
        float *C = g.C + z * g.c_z + (MODE == G_WGRAD ? split * g.c_split : 0);
        const float *aux = g.aux + z * g.aux_z + (MODE == G_DGRAD ? (int64_t)(row_ok ? row : 0) * g.ldaux : 0);
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        constexpr int WQ = TC_PRODUCERS / 128;  // producer warps per TMEM lane quarter
        for (int c0 = (warp >> 2) * 8; c0 < n_tile; c0 += 16 * WQ) {
            float v[2][8], u[2][8], w[2][8];
            float4 x[2][2];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int c = c0 + 8 * WQ * b;
                if (c < n_tile && n_iter > 0) {  // warp-uniform
                    tmem_ld8(t_lane + (uint32_t)c, v[b]);
                    if (TC_NACC == 3) {
                        tmem_ld8(t_lane + (uint32_t)(n_tile + c), u[b]);
                        tmem_ld8(t_lane + (uint32_t)(2 * n_tile + c), w[b]);
                    }
                }
                if (MODE != G_WGRAD) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int nn = n0 + c + 4 * h;
                        x[b][h] = (c < n_tile && nn < g.N && (MODE == G_FWD || row_ok)) ? *reinterpret_cast<const float4 *>(aux + nn) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            tmem_ld_wait();
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int c = c0 + 8 * WQ * b;
                if (c >= n_tile || !row_ok) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int nn = n0 + c + 4 * h;
                    if (nn >= g.N) break;
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = n_iter > 0 ? (TC_NACC == 3 ? v[b][4 * h + i] + (u[b][4 * h + i] + w[b][4 * h + i]) : v[b][4 * h + i]) : 0.f;
                    const float4 a = x[b][h];
                    float4 r = make_float4(o[0], o[1], o[2], o[3]);
                    if (MODE == G_FWD) r = make_float4(fmaxf(o[0] + a.x, 0.f), fmaxf(o[1] + a.y, 0.f), fmaxf(o[2] + a.z, 0.f), fmaxf(o[3] + a.w, 0.f));
                    else if (MODE == G_DGRAD) r = make_float4(a.x > 0.f ? o[0] : 0.f, a.y > 0.f ? o[1] : 0.f, a.z > 0.f ? o[2] : 0.f, a.w > 0.f ? o[3] : 0.f);
                    *reinterpret_cast<float4 *>(C + (int64_t)row * g.ldc + nn) = r;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
}

struct TcTile {
    int n_tile, n_tiles, tmem_cols;
    uint32_t smem_bytes;
};

inline TcTile tc_tile(int N) {
    TcTile t;
    t.n_tiles = TC_NACC == 3 ? (N + 159) / 160 : (N + 255) / 256;  // TC_NACC accumulators of n_tile columns each in the 512 TMEM columns
    t.n_tile = (((N + t.n_tiles - 1) / t.n_tiles) + 15) & ~15;
    t.tmem_cols = 32;
    while (t.tmem_cols < TC_NACC * t.n_tile) t.tmem_cols <<= 1;
    const uint32_t a_plane = TC_KG * (TC_BM / 8) * 128u, b_plane = TC_KG * (uint32_t)(t.n_tile / 8) * 128u;
    t.smem_bytes = 128u + TC_STAGES * 3u * (a_plane + b_plane);
    return t;
}

}  // namespace cstr
