// K3' -- the occupancy MLP in closed form ("affine" mode, precision 2).
//
// As the reference builds it (nof/networks/models.py:152,172: every LeakyReLU has negative_slope == 1.0, layer2 has no
// activation) the network is Linear -> BN (x8) -> Linear -> Sigmoid, i.e. for one BatchNorm batch (= one `chunk` of
// nof/render.py:47-49) the logit is EXACTLY affine in the 63-d encoding x:   logit_r = alpha . x_r + c,
// where (alpha, c) depend on the parameters and on the batch only through its first two moments
//     m = mean_r x_r,   C = mean_r x_r x_r^T - m m^T          (train mode; running statistics in eval mode).
// Hence dL/dtheta = (d alpha/d theta)^T sum_r g_r x_r + (d c/d theta) sum_r g_r  with  g_r = dL/dlogit_r.
// This file holds the three data-sized kernels of that formulation; the parameter-sized algebra
// (moments, theta) -> (alpha, c) and its backward is O(params) work done once per chunk by the host mirror
// (pcnerf_b200/nof/networks/models.py, float64).  Nothing here approximates: results differ from the layered path by
// rounding only (tests/test_gpu_affine.py gates 1e-5 against the reference fixtures).
//
//   pcnerf_affine_moments : per chunk, partial sums of (x - s) and (x - s)(x - s)^T, s = first row of the chunk
//   pcnerf_affine_apply   : p_r = sigmoid(alpha_chunk . x_r + c_chunk)
//   pcnerf_affine_grad    : per chunk, partial sums of g_r x_r and g_r,  g_r = dL/dp_r * p_r (1 - p_r)
#include "common.cuh"

#define AFF_PARTS 64            // CTAs (partial results) per chunk: enough CTAs to fill the machine several times over
#define AFF_SUB 32              // rows per shared-memory sub-tile
#define AFF_FLUSH 256           // rows accumulated in fp32 before being folded into the fp64 accumulators

// out partial layout per (chunk, part): [65][64] doubles: rows 0..63 = sum (x-s)(x-s)^T, row 64 = sum (x-s)
__global__ void __launch_bounds__(256) k_affine_moments(const float* __restrict__ enc, int64_t rows, int64_t chunk,
                                                        double* __restrict__ part) {
    __shared__ __align__(16) float xs[AFF_SUB][64];
    __shared__ float shift[64];
    const int ci = blockIdx.y, pi = blockIdx.x;
    const int64_t c_beg = (int64_t)ci * chunk, c_end = min(rows, c_beg + chunk), c_rows = c_end - c_beg;
    const int64_t per = (c_rows + AFF_PARTS - 1) / AFF_PARTS;
    const int64_t r_beg = c_beg + pi * per, r_end = min(c_end, r_beg + per);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    if (tid < 64) shift[tid] = enc[c_beg * 64 + tid];
    __syncthreads();
    double acc[4][4], acc1 = 0.0;
    float f[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = 0.0; f[i][j] = 0.f; }
    int since = 0;
    for (int64_t r0 = r_beg; r0 < r_end; r0 += AFF_SUB) {
        // 32 rows x 64 floats = 512 float4: two per thread, coalesced; rows past the end contribute zeros
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = tid + h * 256, rr = idx >> 4, c4 = (idx & 15) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + rr < r_end) {
                v = *reinterpret_cast<const float4*>(enc + (r0 + rr) * 64 + c4);
                v.x -= shift[c4]; v.y -= shift[c4 + 1]; v.z -= shift[c4 + 2]; v.w -= shift[c4 + 3];
            }
            *reinterpret_cast<float4*>(&xs[rr][c4]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < AFF_SUB; ++rr) {
            const float4 a = *reinterpret_cast<const float4*>(&xs[rr][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&xs[rr][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) f[i][j] = fmaf(av[i], bv[j], f[i][j]);
        }
        if (tid < 64) {                      // first moment: fp32 over one sub-tile only, then fp64
            float t1 = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < AFF_SUB; ++rr) t1 += xs[rr][tid];
            acc1 += (double)t1;
        }
        __syncthreads();
        since += AFF_SUB;
        if (since >= AFF_FLUSH) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { acc[i][j] += (double)f[i][j]; f[i][j] = 0.f; }
            since = 0;
        }
    }
    double* out = part + ((size_t)ci * AFF_PARTS + pi) * (65 * 64);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[(ty * 4 + i) * 64 + tx * 4 + j] = acc[i][j] + (double)f[i][j];
    if (tid < 64) out[64 * 64 + tid] = acc1;
}

// half-warp per row: lane holds 4 consecutive columns
__global__ void __launch_bounds__(256) k_affine_apply(const float* __restrict__ enc, int64_t rows, int64_t chunk,
                                                      const float* __restrict__ alpha, const float* __restrict__ cc,
                                                      float* __restrict__ out_p) {
    const int lane = threadIdx.x & 31, hl = lane & 15, sub = lane >> 4;
    const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int64_t nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
    int64_t cur_chunk = -1;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float c = 0.f;
    (void)sub;
    for (int64_t r = hw; r < rows; r += nhw) {
        const int64_t ch = r / chunk;
        if (ch != cur_chunk) {
            cur_chunk = ch;
            a = *reinterpret_cast<const float4*>(alpha + ch * 64 + hl * 4);
            c = cc[ch];
        }
        const float4 x = *reinterpret_cast<const float4*>(enc + r * 64 + hl * 4);
        float t = x.x * a.x + x.y * a.y + x.z * a.z + x.w * a.w;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(FULL_MASK, t, o);
        if (hl == 0) out_p[r] = 1.f / (1.f + expf(-(t + c)));
    }
}

// out partial layout per (chunk, part): [65] doubles: 0..63 = sum g x, 64 = sum g
__global__ void __launch_bounds__(256) k_affine_grad(const float* __restrict__ enc, const float* __restrict__ p,
                                                     const float* __restrict__ grad_p, int64_t rows, int64_t chunk,
                                                     double* __restrict__ part) {
    const int ci = blockIdx.y, pi = blockIdx.x;
    const int64_t c_beg = (int64_t)ci * chunk, c_end = min(rows, c_beg + chunk), c_rows = c_end - c_beg;
    const int64_t per = (c_rows + AFF_PARTS - 1) / AFF_PARTS;
    const int64_t r_beg = c_beg + pi * per, r_end = min(c_end, r_beg + per);
    const int tid = threadIdx.x, hl = tid & 15, hw = tid >> 4;       // 16 half-warps per block
    double accd[4] = {0, 0, 0, 0}, gd = 0.0;
    float acc[4] = {0, 0, 0, 0}, gs = 0.f;
    int since = 0;
    for (int64_t r = r_beg + hw; r < r_end; r += 16) {
        const float pv = p[r];
        const float g = grad_p[r] * pv * (1.f - pv);
        const float4 x = *reinterpret_cast<const float4*>(enc + r * 64 + hl * 4);
        acc[0] = fmaf(g, x.x, acc[0]); acc[1] = fmaf(g, x.y, acc[1]);
        acc[2] = fmaf(g, x.z, acc[2]); acc[3] = fmaf(g, x.w, acc[3]);
        gs += g;
        if (++since == 256) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { accd[i] += (double)acc[i]; acc[i] = 0.f; }
            gd += (double)gs;
            gs = 0.f;
            since = 0;
        }
    }
    __shared__ double redd[16][65];
#pragma unroll
    for (int i = 0; i < 4; ++i) redd[hw][hl * 4 + i] = accd[i] + (double)acc[i];
    if (hl == 0) redd[hw][64] = gd + (double)gs;
    __syncthreads();
    if (tid < 65) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) t += redd[k][tid];
        part[((size_t)ci * AFF_PARTS + pi) * 65 + tid] = t;
    }
}

extern "C" int pcnerf_affine_parts(void) { return AFF_PARTS; }

extern "C" int pcnerf_affine_moments(const float* enc, int64_t rows, int64_t chunk, double* out_part, void* stream) {
    PCN_CHECK_ARG(enc && out_part && rows >= 1 && chunk >= 1, "affine_moments: bad arguments");
    const int64_t nchunk = pcn_cdiv(rows, chunk);
    PCN_CHECK_ARG(nchunk <= 65535, "affine_moments: too many chunks (%lld)", (long long)nchunk);
    cudaStream_t st = (cudaStream_t)stream;
    PcnScope ps(PCN_K_AFFINE, st, (double)rows * 256.0);
    k_affine_moments<<<dim3(AFF_PARTS, (unsigned)nchunk), 256, 0, st>>>(enc, rows, chunk, out_part);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_affine_apply(const float* enc, int64_t rows, int64_t chunk, const float* alpha, const float* c,
                                   float* out_p, void* stream) {
    PCN_CHECK_ARG(enc && alpha && c && out_p && rows >= 1 && chunk >= 1, "affine_apply: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = pcn_cdiv(rows, 16);
    if (blocks > PCN_SM_COUNT * 16) blocks = PCN_SM_COUNT * 16;
    PcnScope ps(PCN_K_AFFINE, st, (double)rows * 260.0);
    k_affine_apply<<<(int)blocks, 256, 0, st>>>(enc, rows, chunk, alpha, c, out_p);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_affine_grad(const float* enc, const float* p, const float* grad_p, int64_t rows, int64_t chunk,
                                  double* out_part, void* stream) {
    PCN_CHECK_ARG(enc && p && grad_p && out_part && rows >= 1 && chunk >= 1, "affine_grad: bad arguments");
    const int64_t nchunk = pcn_cdiv(rows, chunk);
    PCN_CHECK_ARG(nchunk <= 65535, "affine_grad: too many chunks (%lld)", (long long)nchunk);
    cudaStream_t st = (cudaStream_t)stream;
    PcnScope ps(PCN_K_AFFINE, st, (double)rows * 264.0);
    k_affine_grad<<<dim3(AFF_PARTS, (unsigned)nchunk), 256, 0, st>>>(enc, p, grad_p, rows, chunk, out_part);
    PCN_LAUNCH_CHECK();
    return 0;
}
