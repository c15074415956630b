// Optimizer step for the two occupancy nets on ONE flat buffer (nof/nof_utils.py:162-173 builds
// torch.optim.Adam(lr, eps = 1e-8, weight_decay) over 68 parameter tensors).  The parameters, their gradients (the same flat
// buffer the NCCL all-reduce runs on, parallel.GradBucket) and the two moments are 994,818-element fp32 vectors: one
// kernel reads 16 B and writes 12 B per parameter.  The 1/world scaling of the summed gradients is folded into the same
// pass.  Step counter and learning rate live on the device (the step can be replayed from a CUDA graph).
// Math = torch.optim.Adam (amsgrad False, maximize False, L2 weight decay):
//   g = grad * grad_scale + wd * p;  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

__global__ void k_adam_tick(int64_t* step, const float* lr, float b1, float b2, float* coef) {
    const int64_t t = *step + 1;
    *step = t;
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    coef[0] = (float)((double)*lr / bc1);        // step size
    coef[1] = (float)sqrt(bc2);                  // sqrt of the second-moment bias correction
}

__global__ void __launch_bounds__(256) k_adam_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, const float* __restrict__ coef, float b1,
                                                   float b2, float eps, float wd, float grad_scale) {
    const float step_size = coef[0], bc2_sqrt = coef[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        float gi = g[i] * grad_scale;
        gi = fmaf(wd, pi, gi);
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

extern "C" int pcnerf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t* step,
                                const float* lr, float* coef2, float beta1, float beta2, float eps, float weight_decay,
                                float grad_scale, void* stream) {
    PCN_CHECK_ARG(n >= 0 && step && lr && coef2, "adam_step: null argument");
    if (n == 0) return 0;
    PCN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adam_step: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    PcnScope ps(PCN_K_MLP_SMALL, st, (double)n * 28.0, 2);
    k_adam_tick<<<1, 1, 0, st>>>(step, lr, beta1, beta2, coef2);
    int64_t blocks = pcn_cdiv(n, 256 * 4);
    if (blocks > PCN_SM_COUNT * 8) blocks = PCN_SM_COUNT * 8;
    k_adam_flat<<<(int)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, coef2, beta1, beta2, eps, weight_decay, grad_scale);
    PCN_LAUNCH_CHECK();
    return 0;
}
