// Shared helpers for the pcnerf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/pcnerf_b200.h"

#define PCN_SM_COUNT 148

void pcn_set_error(const char* fmt, ...);

#define PCN_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            pcn_set_error(__VA_ARGS__);          \
            return PCNERF_ERR_ARG;               \
        }                                        \
    } while (0)

#define PCN_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            pcn_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PCNERF_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)

#define PCN_LAUNCH_CHECK()                                                                    \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            pcn_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PCNERF_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)

// kernel classes of the built-in profiler (pcnerf_prof_*); `work` is algorithmic FLOPs (GEMMs, the closed-form engine's
// moments kernel and float64 algebra) or bytes (the rest)
enum {
    PCN_K_GEMM_FWD = 0, PCN_K_GEMM_DGRAD, PCN_K_GEMM_WGRAD, PCN_K_MLP_SMALL, PCN_K_SAMPLE_ENCODE, PCN_K_COMPOSITE_FWD,
    PCN_K_COMPOSITE_BWD, PCN_K_AABB, PCN_K_SEARCH, PCN_K_AFFINE, PCN_K_AFFINE_MOMENTS, PCN_K_AFFINE_ALGEBRA, PCN_K_COUNT
};

// Counts kernel launches (always) and, when pcnerf_prof_enable(1) is active, brackets them with CUDA events on the
// launching stream.  Declare one in the scope of the launch(es) it covers.
struct PcnScope {
    int id;
    cudaStream_t st;
    cudaEvent_t a, b;
    bool on;
    PcnScope(int id, cudaStream_t st, double work = 0.0, int nlaunch = 1);
    ~PcnScope();
};

#define PCN_TIMED(id, st, work, ...)       \
    do {                                   \
        PcnScope _ps((id), (st), (work));  \
        __VA_ARGS__;                       \
    } while (0)

static inline int64_t pcn_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

#define FULL_MASK 0xffffffffu

// The warp's index within its block as a value the compiler KNOWS to be warp-uniform.  With the plain
// `threadIdx.x >> 5` a warp-per-ray loop counts as potentially divergent and every shuffle / vote inside it is wrapped in
// WARPSYNC ... ENDCOLLECTIVE (ncu on k_composite_fwd_r: 7.7 warp instructions per shuffle-add, 36 % of the kernel).
__device__ __forceinline__ int warp_in_block() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
