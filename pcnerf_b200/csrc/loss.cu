// Masked-mean range losses and logging metrics on (N,) vectors of rendered depths (nof/criteria/loss.py:7-50 -- the
// SmoothL1 / MSE / L1 wrappers behind `nof_loss`, train_kitti.py:145-146 -- and nof/criteria/metrics.py:5-21).
//
// The reference runs each as 3-6 eager torch kernels (boolean gather, subtraction, pointwise loss, mean) plus their autograd
// mirrors.  Here: one forward kernel that reduces sum(elementwise loss) and the number of selected elements into two
// doubles (per-block partials in shared memory, one fp64 atomic pair per block), a one-thread finaliser, and one backward
// kernel that writes dL/dpred (and -dL/dtarget) directly.  kind: 0 SmoothL1 (beta = 1), 1 MSE, 2 L1, 3 abs_error
// (= L1 without a backward), 4 acc_thres (percentage of |pred - gt| < 0.2).
#include "common.cuh"

__device__ __forceinline__ float loss_elem(int kind, float d) {
    const float a = fabsf(d);
    switch (kind) {
        case 0: return a < 1.f ? 0.5f * d * d : a - 0.5f;
        case 1: return d * d;
        case 4: return a < 0.2f ? 1.f : 0.f;
        default: return a;
    }
}
__device__ __forceinline__ float loss_elem_grad(int kind, float d) {
    switch (kind) {
        case 0: return fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f);
        case 1: return 2.f * d;
        default: return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    }
}

__global__ void __launch_bounds__(256) k_masked_loss_fwd(int kind, const float* __restrict__ pred, const float* __restrict__ target,
                                                         const uint8_t* __restrict__ mask, int64_t n, double* __restrict__ acc) {
    __shared__ double rs[8], rc[8];
    double s = 0.0, c = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        s += (double)loss_elem(kind, pred[i] - target[i]);
        c += 1.0;
    }
    s = warp_sum_d(s);
    c = warp_sum_d(c);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rc[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { s += rs[k]; c += rc[k]; }
        atomicAdd(&acc[0], s);
        atomicAdd(&acc[1], c);
    }
}

// out[0] = mean over the selected elements (x 100 for acc_thres); NaN for an empty selection, like torch's mean
__global__ void k_masked_loss_finish(int kind, const double* __restrict__ acc, float* __restrict__ out) {
    const double v = acc[0] / acc[1];
    out[0] = (float)(kind == 4 ? v * 100.0 : v);
}

__global__ void __launch_bounds__(256) k_masked_loss_bwd(int kind, const float* __restrict__ pred, const float* __restrict__ target,
                                                         const uint8_t* __restrict__ mask, int64_t n, const double* __restrict__ acc,
                                                         const float* __restrict__ g_out, float* __restrict__ g_pred,
                                                         float* __restrict__ g_target) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = 0.f;
    if (!mask || mask[i]) g = (*g_out) * loss_elem_grad(kind, pred[i] - target[i]) / (float)acc[1];
    if (g_pred) g_pred[i] = g;
    if (g_target) g_target[i] = -g;
}

extern "C" int pcnerf_masked_loss_fwd(int kind, const float* pred, const float* target, const uint8_t* mask, int64_t n,
                                      double* acc2, float* out, void* stream) {
    PCN_CHECK_ARG(kind >= 0 && kind <= 4 && n >= 0 && acc2 && out && (n == 0 || (pred && target)), "masked_loss_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCN_CUDA(cudaMemsetAsync(acc2, 0, 2 * sizeof(double), st));
    PcnScope ps(PCN_K_COMPOSITE_FWD, st, (double)n * 9.0, 2);
    if (n > 0) {
        int64_t grid = pcn_cdiv(n, 256);
        if (grid > 2 * PCN_SM_COUNT) grid = 2 * PCN_SM_COUNT;
        k_masked_loss_fwd<<<(int)grid, 256, 0, st>>>(kind, pred, target, mask, n, acc2);
    }
    k_masked_loss_finish<<<1, 1, 0, st>>>(kind, acc2, out);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_masked_loss_bwd(int kind, const float* pred, const float* target, const uint8_t* mask, int64_t n,
                                      const double* acc2, const float* g_out, float* g_pred, float* g_target, void* stream) {
    PCN_CHECK_ARG(kind >= 0 && kind <= 2 && n >= 0 && acc2 && g_out, "masked_loss_bwd: bad arguments (kinds 0..2 have a backward)");
    if (n == 0) return 0;
    PCN_CHECK_ARG(pred && target && (g_pred || g_target), "masked_loss_bwd: null argument");
    PcnScope ps(PCN_K_COMPOSITE_BWD, (cudaStream_t)stream, (double)n * 17.0);
    k_masked_loss_bwd<<<(int)pcn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(kind, pred, target, mask, n, acc2, g_out, g_pred,
                                                                               g_target);
    PCN_LAUNCH_CHECK();
    return 0;
}
