// K2, tensor-core actor path (placeholder until the tcgen05 kernel lands in this file).
#include "cstr_abi.cuh"

using namespace cstr;

int cstr_rollout_tc_launch(const cstr_env_params *, int64_t, int64_t, int, const cstr_actor_f32 *, const void *, float, const float *, int,
                           uint32_t, float *, int32_t *, int32_t *, double *, int64_t, int64_t, float *, double *, void *) {
    return fail_arg(CSTR_EINVAL, "rollout: actor_mode 1 (tcgen05) is not available in this build");
}

extern "C" int64_t cstr_actor_pack_bf16(const cstr_actor_f32 *, void *, void *) {
    fail_arg(CSTR_EINVAL, "actor_pack_bf16: not available in this build");
    return -1;
}
