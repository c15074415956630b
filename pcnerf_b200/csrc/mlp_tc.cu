// K3 (bf16 tensor-core path) -- placeholder until the tcgen05 kernels land.
#include "common.cuh"
#include "mlp_layout.h"

int mlp_tc_forward(const pcnerf_mlp_params*, const void*, int64_t, float*, void*, size_t, void*, size_t, cudaStream_t) {
    pcn_set_error("mlp: precision 1 (bf16 tcgen05) is not built yet");
    return PCNERF_ERR_UNSUPPORTED;
}
int mlp_tc_backward(const pcnerf_mlp_params*, const pcnerf_mlp_grads*, const void*, int64_t, const float*, const float*,
                    void*, size_t, void*, size_t, cudaStream_t) {
    pcn_set_error("mlp: precision 1 (bf16 tcgen05) is not built yet");
    return PCNERF_ERR_UNSUPPORTED;
}
