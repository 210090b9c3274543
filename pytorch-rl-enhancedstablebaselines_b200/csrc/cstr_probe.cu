// Pipe-peak microbenchmarks: the roofline denominators bench.py reports K1 against are MEASURED in
// the same run (MEASURED_PEAKS.json has no non-tensor peak).
#include "cstr_abi.cuh"

namespace cstr {

template <int KIND>
__global__ void __launch_bounds__(256) probe_kernel(int64_t iters, float *out) {
    const float seed = (float)(threadIdx.x & 7) * 1e-3f;
    if (KIND == 1) {
        double a[8], b = 1.0000001, c = 1e-9 + seed;
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 1.0 + j * 1e-3;
        for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
        }
        double s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += a[j];
        out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
    } else {
        float a[8], b = 1.0000001f + seed * 1e-3f, c = 1e-9f + seed;
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 1.0f + j * 1e-3f;
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
