// Host-side helpers shared by the C-ABI translation units: error reporting and launch geometry.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cstr_b200.h"

namespace cstr {

char *last_error_buf();  // thread-local, defined in cstr_step.cu

inline int fail_arg(int code, const char *msg) {
    snprintf(last_error_buf(), 256, "%s", msg);
    return code;
}

inline int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    snprintf(last_error_buf(), 256, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

inline int check_launch(const char *what) { return check_cuda(cudaGetLastError(), what); }

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device

// One reactor per thread.  Small batches (65,536 reactors = 443 threads per SM) get 64-thread CTAs so
// the grid spreads evenly over the 148 SMs; large batches get 256-thread CTAs.
inline void env_launch_geometry(int64_t n, int &grid, int &block) {
    const int64_t sms = sm_count();
    block = 256;
    while (block > 64 && (n + block - 1) / block < 8 * sms) block >>= 1;
    grid = (int)((n + block - 1) / block);
}

}  // namespace cstr
