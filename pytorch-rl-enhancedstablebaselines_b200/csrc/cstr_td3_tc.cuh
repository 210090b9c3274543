// Tensor-core version of the TD3 hidden-layer GEMM (same three roles and epilogues as td3_gemm_kernel): tcgen05.mma
// (kind::f16, bf16 operands, fp32 accumulators in TMEM) with each fp32 operand split into three bf16 planes
//     x = x1 + x2 + x3,   x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2)          (24 mantissa bits in total)
// and six products per K-step  a1b1 + a1b2 + a2b1 + a2b2 + a1b3 + a3b1  (the dropped terms are <= 2^-24 |a||b|), i.e.
// float32-grade accuracy at 1/6 of the bf16 tensor rate — 5x the FFMA rate.
//
// CTA = 128 output rows x n_tile (<= 256) output columns, 256 threads.  All threads are producers: they load fp32 from
// global (either operand may be K-contiguous or MN-contiguous in memory — the transposes of the dgrad / wgrad roles are
// done by the producers), split, and store the planes into shared memory in the UMMA no-swizzle K-major core-matrix layout
// ([k-group of 8][row-group of 8][8 rows x 16 B]); thread 0 then issues the 6 MMAs of the stage and commits them to the
// stage's mbarrier, which frees the stage for reuse (3 stages of K=16).  Epilogue: tcgen05.ld -> bias+relu / relu-mask /
// slab store, 32-byte row pieces straight to global.
#pragma once
#include <cuda_bf16.h>

namespace cstr {
namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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

// split 8 consecutive-k floats into the three bf16 planes and store one 16-byte core-matrix row per plane
__device__ __forceinline__ void split_store(const float v[8], uint32_t addr, uint32_t plane_stride) {
    uint32_t p1[4], p2[4], p3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        p1[i] = pack_bf16x2(a, b);
        const float ra = a - __uint_as_float(p1[i] << 16), rb = b - __uint_as_float(p1[i] & 0xffff0000u);
        p2[i] = pack_bf16x2(ra, rb);
        const float sa = ra - __uint_as_float(p2[i] << 16), sb = rb - __uint_as_float(p2[i] & 0xffff0000u);
        p3[i] = pack_bf16x2(sa, sb);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(p1[0]), "r"(p1[1]), "r"(p1[2]), "r"(p1[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + plane_stride), "r"(p2[0]), "r"(p2[1]), "r"(p2[2]), "r"(p2[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 2 * plane_stride), "r"(p3[0]), "r"(p3[1]), "r"(p3[2]), "r"(p3[3]) : "memory");
}

}  // namespace tc5

constexpr int TC_BM = 128, TC_KC = 16, TC_STAGES = 3, TC_GEMM_THREADS = 256, TC_MAX_BCHUNKS = 2;  // n_tile <= 256 -> <= 512 B chunks / 256 threads

// 8 floats = one (row r, k-group) chunk of an operand tile.  KMAJ: memory rows are operand rows (k contiguous);
// otherwise memory rows are k (operand rows contiguous).
template <bool KMAJ>
__device__ __forceinline__ void load_chunk(const float *__restrict__ src, int ld, int r, int r_end, int k, int k_end, float v[8]) {
    if (KMAJ) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 a = z, b = z;
        if (r < r_end) {
            const float *p = src + (int64_t)r * ld + k;
            if (k < k_end) a = *reinterpret_cast<const float4 *>(p);
            if (k + 4 < k_end) b = *reinterpret_cast<const float4 *>(p + 4);
        }
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (r < r_end && k + j < k_end) ? src[(int64_t)(k + j) * ld + r] : 0.f;
    }
}

template <int MODE>
__global__ void __launch_bounds__(TC_GEMM_THREADS) td3_gemm_tc_kernel(GemmArgs g, int n_tile, int tmem_cols) {
    using namespace tc5;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr bool A_KMAJ = MODE != G_WGRAD, B_KMAJ = MODE == G_FWD;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bars = smem_base;  // TC_STAGES mbarriers
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + 64);
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
    // one stage: A planes 3 x [2 k-groups][16 row-groups][128 B], then B planes 3 x [2][n_tile/8][128 B]
    const uint32_t a_plane = 2u * (TC_BM / 8) * 128u, b_plane = 2u * (uint32_t)(n_tile / 8) * 128u;
    const uint32_t stage_bytes = 3u * (a_plane + b_plane), stage0 = smem_base + 128u;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) mbar_init(bars + 8 * s, 1);
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
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = umma_idesc_bf16(TC_BM, n_tile);

    // this thread's chunks: A tile = 128 rows x 2 k-groups = 256 chunks (one each); B tile = n_tile x 2 (<= 2 each)
    const int a_r = tid & 127, a_kg = tid >> 7;
    const uint32_t a_off = (uint32_t)a_kg * (TC_BM / 8) * 128u + (uint32_t)(a_r >> 3) * 128u + (uint32_t)(a_r & 7) * 16u;
    const int a_rows_end = MODE == G_WGRAD ? g.M : g.M;  // operand-row bound (rows of C)
    const int b_chunks = n_tile * 2;
    float va[8], vb[TC_MAX_BCHUNKS][8];
    auto load_stage = [&](int it) {
        const int k0 = k_begin + it * TC_KC;
        load_chunk<A_KMAJ>(A, g.lda, m0 + a_r, a_rows_end, k0 + a_kg * 8, k_end, va);
#pragma unroll
        for (int c = 0; c < TC_MAX_BCHUNKS; ++c) {
            const int ch = tid + c * TC_GEMM_THREADS;
            if (ch < b_chunks) {
                const int r = ch % n_tile, kg = ch / n_tile;
                load_chunk<B_KMAJ>(Bm, g.ldb, n0 + r, g.N, k0 + kg * 8, k_end, vb[c]);
            }
        }
    };
    auto store_stage = [&](int s) {
        const uint32_t sa = stage0 + (uint32_t)s * stage_bytes, sb = sa + 3u * a_plane;
        split_store(va, sa + a_off, a_plane);
#pragma unroll
        for (int c = 0; c < TC_MAX_BCHUNKS; ++c) {
            const int ch = tid + c * TC_GEMM_THREADS;
            if (ch < b_chunks) {
                const int r = ch % n_tile, kg = ch / n_tile;
                split_store(vb[c], sb + (uint32_t)kg * (uint32_t)(n_tile / 8) * 128u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u, b_plane);
            }
        }
    };

    if (n_iter > 0) load_stage(0);
    for (int it = 0; it < n_iter; ++it) {
        const int s = it % TC_STAGES, use = it / TC_STAGES;
        if (use > 0) mbar_wait(bars + 8 * s, (uint32_t)((use - 1) & 1));  // the MMAs that read this stage have finished
        store_stage(s);
        if (it + 1 < n_iter) load_stage(it + 1);  // global loads in flight across the barrier and the MMA issue
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t sa = stage0 + (uint32_t)s * stage_bytes, sb = sa + 3u * a_plane;
            uint64_t da[3], db[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                da[p] = umma_desc(sa + p * a_plane, (TC_BM / 8) * 128u, 128u);
                db[p] = umma_desc(sb + p * b_plane, (uint32_t)(n_tile / 8) * 128u, 128u);
            }
            tc_mma_bf16(tmem_base, da[0], db[2], idesc, it > 0);  // small terms first
            tc_mma_bf16(tmem_base, da[2], db[0], idesc, 1);
            tc_mma_bf16(tmem_base, da[1], db[1], idesc, 1);
            tc_mma_bf16(tmem_base, da[0], db[1], idesc, 1);
            tc_mma_bf16(tmem_base, da[1], db[0], idesc, 1);
            tc_mma_bf16(tmem_base, da[0], db[0], idesc, 1);
            tc_commit(bars + 8 * s);
        }
    }
    if (n_iter > 0) {
        const int last = n_iter - 1;
        mbar_wait(bars + 8 * (last % TC_STAGES), (uint32_t)((last / TC_STAGES) & 1));
    }
    tc_fence_after();

    // ---- epilogue: thread = output row (TMEM lane), two warps share a lane quarter and split the columns -------------------
    const int row = m0 + (warp & 3) * 32 + lane;
    const int half_cols = n_tile / 2, c_begin = (warp >> 2) * half_cols;
    float *C = g.C + z * g.c_z + (MODE == G_WGRAD ? split * g.c_split : 0);
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = c_begin; c < c_begin + half_cols; c += 8) {
        float v[8];
        if (n_iter > 0) {
            tmem_ld8(t_lane + (uint32_t)c, v);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        const int n = n0 + c;
        if (row >= g.M || n >= g.N) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int nn = n + 4 * h;
            if (nn >= g.N) break;
            float4 o = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
            if (MODE == G_FWD) {
                const float4 b = *reinterpret_cast<const float4 *>(g.aux + z * g.aux_z + nn);
                o = make_float4(fmaxf(o.x + b.x, 0.f), fmaxf(o.y + b.y, 0.f), fmaxf(o.z + b.z, 0.f), fmaxf(o.w + b.w, 0.f));
            } else if (MODE == G_DGRAD) {
                const float4 hh = *reinterpret_cast<const float4 *>(g.aux + z * g.aux_z + (int64_t)row * g.ldaux + nn);
                o = make_float4(hh.x > 0.f ? o.x : 0.f, hh.y > 0.f ? o.y : 0.f, hh.z > 0.f ? o.z : 0.f, hh.w > 0.f ? o.w : 0.f);
            }
            *reinterpret_cast<float4 *>(C + (int64_t)row * g.ldc + nn) = o;
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
    t.n_tiles = (N + 255) / 256;
    t.n_tile = (((N + t.n_tiles - 1) / t.n_tiles) + 15) & ~15;
    t.tmem_cols = 32;
    while (t.tmem_cols < t.n_tile) t.tmem_cols <<= 1;
    const uint32_t a_plane = 2u * (TC_BM / 8) * 128u, b_plane = 2u * (uint32_t)(t.n_tile / 8) * 128u;
    t.smem_bytes = 128u + TC_STAGES * 3u * (a_plane + b_plane);
    return t;
}

}  // namespace cstr
