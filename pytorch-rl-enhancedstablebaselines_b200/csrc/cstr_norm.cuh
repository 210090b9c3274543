// VecNormalize value maps (core/common/vec_env/vec_normalize.py:225-259), float64 arithmetic as in the reference.
#pragma once
#include <cuda_runtime.h>

namespace cstr {

// np.clip((obs - mean) / np.sqrt(var + eps), -clip, clip).astype(float32)
__device__ __forceinline__ float normalize_obs_value(float x, double mean, double var, double eps, double clip) {
    const double z = ((double)x - mean) / sqrt(var + eps);
    return (float)fmin(fmax(z, -clip), clip);
}
// np.clip(reward / np.sqrt(ret_var + eps), -clip, clip).astype(float32)
__device__ __forceinline__ float normalize_reward_value(float r, double ret_var, double eps, double clip) {
    const double z = (double)r / sqrt(ret_var + eps);
    return (float)fmin(fmax(z, -clip), clip);
}

}  // namespace cstr
