// Pipe-peak microbenchmarks: the roofline denominators bench.py reports K1 against are MEASURED in
// the same run (MEASURED_PEAKS.json has no non-tensor peak).
#include <string.h>

#include "cstr_abi.cuh"

namespace cstr {

__device__ __forceinline__ uint64_t probe_globaltimer() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// out[0] (thread 0 of block 0) additionally carries the SM clock the chains actually ran at, in MHz: clock64 cycles / globaltimer ns
// over the thread's own loop — the pipe peak a launch can reach is SMs x lanes x 2 x THIS clock, not the nominal maximum
template <int KIND>
__global__ void __launch_bounds__(256) probe_kernel(int64_t iters, float *out) {
    const float seed = (float)(threadIdx.x & 7) * 1e-3f;
    const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
    const uint64_t t0 = timer ? probe_globaltimer() : 0;
    const long long c0 = timer ? clock64() : 0;
    if (KIND == 1) {
        double a[8], b = 1.0000001, c = 1e-9 + seed;
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 1.0 + j * 1e-3;
        // 8 trips unrolled: 64 DFMA per backward branch, so the loop's add/compare/branch take < 5 % of the issue slots
#pragma unroll 8
        for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
        }
        double s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += a[j];
        out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
        if (timer) {
            const long long c1 = clock64();
            const uint64_t t1 = probe_globaltimer();
            out[0] = (s != s ? 1.f : 0.f) + (float)(1e3 * (double)(c1 - c0) / (double)(t1 - t0 ? t1 - t0 : 1));
        }
    } else {
        float a[8], b = 1.0000001f + seed * 1e-3f, c = 1e-9f + seed;
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 1.0f + j * 1e-3f;
        // 8 trips unrolled (64 FFMA per branch): the rolled loop of round 1 spent ~12 % of its issue slots on loop overhead and read
        // 65.5 TFLOP/s where the clock-derived peak (SMs x 128 lanes x 2 x f_clk) is 74.4
#pragma unroll 8
        for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (KIND == 0) a[j] = __fmaf_rn(a[j], b, c);
                if (KIND == 2) a[j] = __fadd_rn(__fmul_rn(a[j], b), c);
                if (KIND == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
            }
        }
        float s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += a[j];
        out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
        if (timer) {
            const long long c1 = clock64();
            const uint64_t t1 = probe_globaltimer();
            out[0] = (s != s ? 1.f : 0.f) + (float)(1e3 * (double)(c1 - c0) / (double)(t1 - t0 ? t1 - t0 : 1));
        }
    }
}

}  // namespace cstr

using namespace cstr;

extern "C" int cstr_probe_pipe(int kind, int64_t iters, int grid, int block, float *out, void *stream) {
    if (!out || iters < 0 || grid <= 0 || block <= 0 || block > 256) return fail_arg(CSTR_EINVAL, "probe: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case 0: probe_kernel<0><<<grid, block, 0, st>>>(iters, out); break;
        case 1: probe_kernel<1><<<grid, block, 0, st>>>(iters, out); break;
        case 2: probe_kernel<2><<<grid, block, 0, st>>>(iters, out); break;
        case 3: probe_kernel<3><<<grid, block, 0, st>>>(iters, out); break;
        default: return fail_arg(CSTR_EINVAL, "probe: unknown kind");
    }
    return check_launch("probe_kernel");
}

// ---------------------------------------------------------------------------------------------------
// Exhaustive self-tests of the strict kernel's arithmetic shortcuts (each must be bit-identical to the
// plain IEEE operation it replaces).  Results: number of mismatching inputs, must be 0.
//   0  div_neg_e(y) == __fdiv_rn(-83140, y)        for EVERY float32 y in [2200, 3400]  (R*T, T in [273.15,400])
//   1  expf_shared_normal(x) == expf_shared(x)     for EVERY float32 x in [-40, -20]     (-E/(R T) range)
//   2  CSTR_DIV_CONST(x, c) == __fdiv_rn(x, c)      for EVERY float32 x with |x| in [2^-100, 2^100] and every
//      constant divisor c the step uses
// ---------------------------------------------------------------------------------------------------
#include "cstr_device.cuh"

namespace cstr {

template <typename F>
__global__ void sweep_kernel(uint32_t lo_bits, uint32_t hi_bits, unsigned long long *mismatches, F f) {
    unsigned long long bad = 0;
    for (uint64_t b = (uint64_t)lo_bits + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= hi_bits;
         b += (uint64_t)gridDim.x * blockDim.x)
        bad += f(__uint_as_float((uint32_t)b)) ? 0ull : 1ull;
    if (bad) atomicAdd(mismatches, bad);
}

struct CheckDivNegE {
    __device__ bool operator()(float y) const { return __float_as_uint(div_neg_e(y)) == __float_as_uint(__fdiv_rn(-83140.0f, y)); }
};
struct CheckExpNormal {
    __device__ bool operator()(float x) const { return __float_as_uint(expf_shared_normal(x)) == __float_as_uint(expf_shared(x)); }
};
template <int WHICH>
struct CheckDivConst {
    __device__ bool operator()(float x) const {
        float a, b;
        switch (WHICH) {
            case 0: a = CSTR_DIV_CONST(x, 239.0f); b = __fdiv_rn(x, 239.0f); break;
            case 1: a = CSTR_DIV_CONST(x, CSTR_HALF_C); b = __fdiv_rn(x, CSTR_HALF_C); break;
            case 2: a = CSTR_DIV_CONST(x, CSTR_HALF_T); b = __fdiv_rn(x, CSTR_HALF_T); break;
            case 3: a = CSTR_DIV_CONST(x, 0.4f); b = __fdiv_rn(x, 0.4f); break;
            case 4: a = CSTR_DIV_CONST(x, 280.0f); b = __fdiv_rn(x, 280.0f); break;
            default: a = CSTR_DIV_CONST(x, 350.0f); b = __fdiv_rn(x, 350.0f); break;
        }
        return __float_as_uint(a) == __float_as_uint(b) && __float_as_uint(CSTR_DIV_CONST(-x, 239.0f)) == __float_as_uint(__fdiv_rn(-x, 239.0f));
    }
};

}  // namespace cstr

extern "C" int cstr_selftest(int which, unsigned long long *mismatches, void *stream) {
    if (!mismatches) return fail_arg(CSTR_EINVAL, "selftest: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = sm_count() * 8, block = 256;
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    switch (which) {
        case 0: sweep_kernel<<<grid, block, 0, st>>>(bits(2200.0f), bits(3400.0f), mismatches, CheckDivNegE()); break;
        case 1: sweep_kernel<<<grid, block, 0, st>>>(bits(20.0f) | 0x80000000u, bits(40.0f) | 0x80000000u, mismatches, CheckExpNormal()); break;
        case 2: {
            const uint32_t lo = bits(7.888609052210118e-31f) /* 2^-100 */, hi = bits(1.2676506002282294e30f) /* 2^100 */;
            sweep_kernel<<<grid, block, 0, st>>>(lo, hi, mismatches, CheckDivConst<0>());
            sweep_kernel<<<grid, block, 0, st>>>(lo, hi, mismatches, CheckDivConst<1>());
            sweep_kernel<<<grid, block, 0, st>>>(lo, hi, mismatches, CheckDivConst<2>());
            sweep_kernel<<<grid, block, 0, st>>>(lo, hi, mismatches, CheckDivConst<3>());
            sweep_kernel<<<grid, block, 0, st>>>(lo, hi, mismatches, CheckDivConst<4>());
            sweep_kernel<<<grid, block, 0, st>>>(lo, hi, mismatches, CheckDivConst<5>());
            break;
        }
        default: return fail_arg(CSTR_EINVAL, "selftest: unknown test");
    }
    return check_launch("selftest");
}
