// Small (non-GEMM) kernels of the occupancy MLP, shared by the fp32 path (mlp.cu: TH = TG = float) and the
// tensor-core path (mlp_tc.cu: activations TH = __half, gradients TG = __nv_bfloat16).
#pragma once
#include "common.cuh"
#include "mlp_layout.h"
#include <cuda_fp16.h>

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------------------

struct PrepArgs {
    const float* W[8];
    float* Wp[8];
};

// Padded copies of the hidden Linear weights: layer 0 -> [256,64] (col 63 = 0); layer 4 -> [256,320]
// ([0,63) encoding cols, col 63 = 0, [64,320) hidden cols); others [256,256].
static __global__ void k_prep_weights(PrepArgs a) {
    const int l = blockIdx.y;
    const int kin = mlp_kin(l), kpad = mlp_kpad(l);
    const int total = 256 * kpad;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int o = idx / kpad, c = idx - o * kpad;
        float v = 0.f;
        if (l == 0) { if (c < 63) v = a.W[0][o * kin + c]; }
        else if (l == 4) { if (c < 63) v = a.W[4][o * kin + c]; else if (c >= 64) v = a.W[4][o * kin + c - 1]; }
        else v = a.W[l][o * kin + c];
        a.Wp[l][idx] = v;
    }
}

// BN(l) statistics -> (mean, invstd, a, s) and fold into layer l+1 (or the output layer when l == 7).
// grid: 256 blocks (output feature o of the next layer) x 256 threads (hidden input feature i); l == 7: 1 block.
static __global__ void __launch_bounds__(256) k_bn_fold(int l, int training, int64_t rows, const double* __restrict__ sum,
                                                 const double* __restrict__ sumsq, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float* __restrict__ running_mean,
                                                 float* __restrict__ running_var, int64_t* __restrict__ nbt,
                                                 float momentum, float eps, float* __restrict__ stats /* [4][256] */,
                                                 const float* __restrict__ Wp_next, const float* __restrict__ b_next,
                                                 float* __restrict__ Wf_next, float* __restrict__ bf_next,
                                                 __half* __restrict__ Wh_next /* fp16 copy of Wf_next or NULL */) {
    __shared__ float red[8];
    const int i = threadIdx.x, o = blockIdx.x;
    float mean, var;
    if (training) {
        const double m = sum[i] / (double)rows;
        double v = sumsq[i] / (double)rows - m * m;
        if (v < 0) v = 0;
        mean = (float)m;
        var = (float)v;
    } else {
        mean = running_mean[i];
        var = running_var[i];
    }
    const float invstd = 1.f / sqrtf(var + eps);
    const float a = gamma[i] * invstd;
    const float s = beta[i] - mean * a;
    if (o == 0) {
        stats[i] = mean; stats[256 + i] = invstd; stats[512 + i] = a; stats[768 + i] = s;
        if (training) {
            const float unbiased = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
            running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * mean;
            running_var[i] = (1.f - momentum) * running_var[i] + momentum * unbiased;
            if (i == 0 && nbt) *nbt += 1;
        }
    }
    const int nl = l + 1;
    float part;
    if (nl < 8) {
        const int kpad = mlp_kpad(nl), off = nl == 4 ? 64 : 0;
        const float w = Wp_next[o * kpad + off + i];
        Wf_next[o * kpad + off + i] = w * a;
        if (nl == 4 && i < 64) Wf_next[o * kpad + i] = Wp_next[o * kpad + i];
        if (Wh_next) {
            Wh_next[o * kpad + off + i] = __float2half_rn(w * a);
            if (nl == 4 && i < 64) Wh_next[o * kpad + i] = __float2half_rn(Wp_next[o * kpad + i]);
        }
        part = w * s;
    } else {
        const float w = Wp_next[i];            // occ_out weight (1,256)
        Wf_next[i] = w * a;
        part = w * s;
    }
    part = warp_sum(part);
    if ((i & 31) == 0) red[i >> 5] = part;
    __syncthreads();
    if (i == 0) {
        float t = 0;
        for (int k = 0; k < 8; ++k) t += red[k];
        bf_next[o] = b_next[o] + t;
    }
}

// p = sigmoid(h8 . wout_f + bout_f): one warp per row
template <class TH>
__device__ __forceinline__ void ld8(const TH* p, float* o);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float* o) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<__half>(const __half* p, float* o) {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}

template <>
__device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float* o) {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void st8(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        pk[i] = *reinterpret_cast<const uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

template <class TH>
__global__ void k_logit_sigmoid(const TH* __restrict__ H, int64_t rows, const float* __restrict__ wf,
                                const float* __restrict__ bf, float* __restrict__ out_p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = wf[lane * 8 + j];
    const float b = bf[0];
    for (int64_t r = warp; r < rows; r += nw) {
        float hv[8];
        ld8<TH>(H + r * 256 + lane * 8, hv);
        float t = hv[0] * wv[0] + hv[1] * wv[1] + hv[2] * wv[2] + hv[3] * wv[3] + hv[4] * wv[4] + hv[5] * wv[5] +
                  hv[6] * wv[6] + hv[7] * wv[7];
        t = warp_sum(t);
        if (lane == 0) out_p[r] = 1.f / (1.f + expf(-(t + b)));
    }
}

#define STRIP 64

// g = dL/dp * p(1-p);  acc[j] += sum_r g_r H8[r,j];  acc[256] += sum_r g_r.
// One warp per row (lane = 8 consecutive columns, 16/32-byte loads), STRIP rows per block, block reduction in shared
// memory, one fp64 atomic per column per block.
template <class TH>
__global__ void __launch_bounds__(256) k_out_bwd_reduce(const float* __restrict__ grad_p, const float* __restrict__ p,
                                                        const TH* __restrict__ H, int64_t rows,
                                                        float* __restrict__ gvec, double* __restrict__ acc) {
    __shared__ float red[8][257];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sg = 0.f;
    // Grid-stride over strips of STRIP rows (the launch caps the grid at a few CTAs per SM: every CTA ends with one fp64
    // atomic per column, and thousands of same-address atomics serialise in L2 -- 58 us for 4096 CTAs, measured).
    // Within a strip warp w owns rows r0 + w*8 .. + 7; four independent row loads in flight per lane.
    for (int64_t r0 = (int64_t)blockIdx.x * STRIP; r0 < rows; r0 += (int64_t)gridDim.x * STRIP)
#pragma unroll
    for (int b = 0; b < STRIP / 8; b += 4) {
        float hv[4][8], gr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = r0 + w * (STRIP / 8) + b + u;
            gr[u] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) hv[u][i] = 0.f;
            if (r < rows) {
                const float pv = p[r];
                gr[u] = grad_p[r] * pv * (1.f - pv);
                ld8<TH>(H + r * 256 + lane * 8, hv[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = r0 + w * (STRIP / 8) + b + u;
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(gr[u], hv[u][i], a[i]);
            sg += gr[u];
            if (lane == 0 && r < rows) gvec[r] = gr[u];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = a[i];
    if (lane == 0) red[w][256] = sg;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(acc + threadIdx.x, (double)t);
    if (threadIdx.x == 0) {
        float ts = 0.f;
        for (int k = 0; k < 8; ++k) ts += red[k][256];
        atomicAdd(acc + 256, (double)ts);
    }
}

// Output-layer parameter grads, BN(7) grads and the coefficient vectors of
//   DH8[r,j] = g_r*u[j] - c1[j] - (H8[r,j] - mean[j])*c2[j]
static __global__ void __launch_bounds__(256) k_out_bwd_finalize(const double* __restrict__ acc, int64_t rows,
                                                          const float* __restrict__ w_out,
                                                          const float* __restrict__ stats, float* __restrict__ dw_out,
                                                          float* __restrict__ db_out, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, float* __restrict__ coef) {
    const int j = threadIdx.x;
    const float q = (float)acc[j], sg = (float)acc[256];
    const float mean = stats[j], invstd = stats[256 + j], a = stats[512 + j], s = stats[768 + j];
    const float w = w_out[j];
    dw_out[j] += a * q + s * sg;
    if (j == 0) db_out[0] += sg;
    const float db = w * sg;
    const float dg = w * invstd * (q - mean * sg);
    dgamma[j] += dg;
    dbeta[j] += db;
    const float B = (float)rows;
    coef[j] = a * w;                       // u
    coef[256 + j] = a * db / B;            // c1
    coef[512 + j] = a * invstd * dg / B;   // c2
}

// LAST: DH = g (x) u - c1 - (H - mean) c2     (Gy never materialised for the last BN)
// else: DH = Gy*a - c1 - (H - mean) c2        in place on Gy
// plus column sums of DH (the Linear bias gradient; zero in exact arithmetic).
// One warp per row, lane = 8 consecutive columns.
template <bool LAST, class TH, class TG>
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const float* __restrict__ gvec, TG* __restrict__ G,
                                                      const TH* __restrict__ H, int64_t rows,
                                                      const float* __restrict__ coef, const float* __restrict__ stats,
                                                      double* __restrict__ colsum) {
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float c0[8], c1[8], c2[8], mean[8], cs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = lane * 8 + i;
        c0[i] = coef[j]; c1[i] = coef[256 + j]; c2[i] = coef[512 + j]; mean[i] = stats[j]; cs[i] = 0.f;
    }
    // grid-stride over strips (see k_out_bwd_reduce); warp w owns rows r0 + w*8 .. + 7, four row loads in flight per lane
    for (int64_t r0 = (int64_t)blockIdx.x * STRIP; r0 < rows; r0 += (int64_t)gridDim.x * STRIP)
#pragma unroll
    for (int b = 0; b < STRIP / 8; b += 4) {
        float hv[4][8], up[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = r0 + w * (STRIP / 8) + b + u;
            if (r < rows) {
                ld8<TH>(H + r * 256 + lane * 8, hv[u]);
                if (LAST) {
                    const float gr = gvec[r];
#pragma unroll
                    for (int i = 0; i < 8; ++i) up[u][i] = gr;
                } else {
                    ld8<TG>(G + r * 256 + lane * 8, up[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t r = r0 + w * (STRIP / 8) + b + u;
            if (r < rows) {
                float dh[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    dh[i] = up[u][i] * c0[i] - c1[i] - (hv[u][i] - mean[i]) * c2[i];
                    cs[i] += dh[i];
                }
                st8(G + r * 256 + lane * 8, dh);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = cs[i];
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(colsum + threadIdx.x, (double)t);
}

// BN(l) backward coefficients from the dgrad epilogue sums: st0 = sum Gy, st1 = sum Gy*H
static __global__ void __launch_bounds__(256) k_bn_bwd_coef(const double* __restrict__ st0, const double* __restrict__ st1,
                                                     int64_t rows, const float* __restrict__ stats,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ coef) {
    const int j = threadIdx.x;
    const float mean = stats[j], invstd = stats[256 + j], a = stats[512 + j];
    const float db = (float)st0[j];
    const float dg = invstd * (float)(st1[j] - (double)mean * st0[j]);
    dgamma[j] += dg;
    dbeta[j] += db;
    const float B = (float)rows;
    coef[j] = a;
    coef[256 + j] = a * db / B;
    coef[512 + j] = a * invstd * dg / B;
}

// dW_l[o, real col] += (sum_splits partial[o, c]) * a_prev[c] + dbias[o] * s_prev[c];  db_l[o] += dbias[o]
// partial is [splits][256][kpad].  prev_stats == NULL for encoding columns (a = 1, s = 0).
static __global__ void k_wgrad_finalize(int l, const float* __restrict__ partial, int splits,
                                 const double* __restrict__ colsum, const float* __restrict__ prev_stats,
                                 float* __restrict__ dW, float* __restrict__ db) {
    const int kin = mlp_kin(l), kpad = mlp_kpad(l);
    const int total = 256 * kpad;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int o = idx / kpad, c = idx - o * kpad;
        int real = c, hid = c;
        bool is_hidden = true;
        if (l == 0) { is_hidden = false; if (c >= 63) continue; }
        else if (l == 4) {
            if (c < 64) { is_hidden = false; if (c == 63) continue; }
            else { real = c - 1; hid = c - 64; }
        }
        double t = 0;
        for (int s = 0; s < splits; ++s) t += (double)partial[(size_t)s * total + idx];
        float v = (float)t;
        const float dbias = (float)colsum[o];
        if (is_hidden) v = v * prev_stats[512 + hid] + dbias * prev_stats[768 + hid];
        dW[o * kin + real] += v;
        if (c == 0) db[o] += dbias;
    }
}

