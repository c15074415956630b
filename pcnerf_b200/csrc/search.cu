// K5: parent-then-child depth-inference search (nof/render.py:229-368) and the output points (:674-684).
//
// Per candidate row (one warp per row): normalised weights, strict child mask with expand-until-non-empty,
// Gaussian smoothing of the weights (scipy.ndimage.gaussian_filter sigma=5: radius 20, 'reflect', fp64 accumulation in
// NI_Correlate1D's symmetric order, rounded to fp32) + first-max argmax, peak-in-child flag, in-child weight sum,
// depth by method 1 or 2.  A second kernel picks one winner per candidate group.  The fp64 smoothing is compiled
// with explicit __dmul_rn/__dadd_rn so that the argmax is bit-identical to SciPy's.
#include "common.cuh"

#define SRCH_MAX_SMEM (200 * 1024)
#define GAUSS_R 20

struct GaussK { double k[GAUSS_R + 1]; };   // k[j] = weight at offset j-GAUSS_R (j = 0..R), symmetric

struct MaskBounds { float lo, hi; };

__device__ __forceinline__ MaskBounds strict_bounds(const float* zs, int P, float cn, float cf, int lane) {
    double g = 0.01;                          // render.py:253
    MaskBounds b;
    for (int it = 0; it < 1000000; ++it) {
        b.lo = __fsub_rn(cn, (float)g);
        b.hi = __fadd_rn(cf, (float)g);
        int any = 0;
        for (int i = lane; i < P; i += 32) { const float z = zs[i]; any |= (b.lo < z && z < b.hi); }
        if (__any_sync(FULL_MASK, any)) break;
        g = g + 0.01;
    }
    return b;
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // scipy 'reflect' (d c b a | a b c d | d c b a), valid for any extension length
    const int per = 2 * n;
    int m = i % per;
    if (m < 0) m += per;
    return m >= n ? per - 1 - m : m;
}

// row_ray != NULL ("grouped" form): p, z and w are indexed by PHYSICAL ray, row r reads ray row_ray[r] (all candidate rows of
// a group share origin, direction and the parent segment, nof/render.py:619-628, so their samples, occupancies and
// weights are identical: they are evaluated once per ray); the head row of every group (the first row that maps to the
// ray) writes the ray's weights.  row_ray == NULL: one (p, z, w) row per candidate row, as the reference evaluates them.
__global__ void k_search_rows(const float* __restrict__ p, const float* __restrict__ z, const float* __restrict__ rays,
                              int ld, int64_t n, int P, int cnear_col, int cfar_col, float epsilon, int method,
                              GaussK gk, const int32_t* __restrict__ row_ray, float* __restrict__ w,
                              float* __restrict__ depth, uint8_t* __restrict__ peak_in, float* __restrict__ wsum_child,
                              double* __restrict__ sums) {
    extern __shared__ double smd[];
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, wib = warp_in_block(), wpb = blockDim.x >> 5;
    // per warp: the weights as doubles with GAUSS_R reflected samples on either side | z | w
    double* sd = smd + (size_t)wib * (2 * P + 2 * GAUSS_R);
    float* sz = reinterpret_cast<float*>(sd + P + 2 * GAUSS_R);
    float* sw = sz + P;
    double acc_op = 0;
    for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < n; r += (int64_t)gridDim.x * wpb) {
        const int64_t g = row_ray ? (int64_t)row_ray[r] : r;               // row of p / z / w
        const bool put_w = w != nullptr && (!row_ray || r == 0 || row_ray[r - 1] != row_ray[r]);
        for (int i = lane; i < P; i += 32) sz[i] = z[g * P + i];
        // weights (render.py:241-246)
        float carry = 1.f, sumv = 0.f, op = 0.f;
        for (int base = 0; base < P; base += 32) {
            const int i = base + lane;
            const float pi = i < P ? p[g * P + i] : 0.f;
            const float fr = __fsub_rn(1.f, pi);
            float incl = fr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl *= t;
            }
            float excl = __shfl_up_sync(FULL_MASK, incl, 1);
            if (lane == 0) excl = 1.f;
            const float v = carry * excl * pi;
            carry = carry * __shfl_sync(FULL_MASK, incl, 31);
            if (i < P) {
                sw[i] = v;
                sumv += v;
                op += __fadd_rn(__fadd_rn(logf(__fadd_rn(0.1f, pi)), logf(__fadd_rn(0.1f, fr))), 2.20727f);
            }
        }
        const float denom = __fadd_rn(warp_sum(sumv), epsilon);
        acc_op += (double)warp_sum(op);
        for (int i = lane; i < P; i += 32) {
            const float wi = __fdiv_rn(sw[i], denom);
            sw[i] = wi;
            if (put_w) w[g * P + i] = wi;
        }
        __syncwarp();
        const float cn = rays[r * ld + cnear_col], cf = rays[r * ld + cfar_col];
        const MaskBounds b = strict_bounds(sz, P, cn, cf, lane);
        // Gaussian smoothing + first-max argmax (render.py:303-308)
        // The 'reflect' extension is materialised once per row (fp32 -> fp64 once per sample as well): the 41-tap loop
        // is then 2 loads + 3 fp64 operations per tap pair, same operands in the same order, instead of two integer
        // modulo reductions and two conversions per tap pair.
        for (int i = lane; i < P + 2 * GAUSS_R; i += 32) sd[i] = (double)sw[reflect_idx(i - GAUSS_R, P)];
        __syncwarp();
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int i = lane; i < P; i += 32) {
            const double* c0 = sd + GAUSS_R + i;
            double t = __dmul_rn(c0[0], gk.k[GAUSS_R]);
#pragma unroll
            for (int j = -GAUSS_R; j < 0; ++j)
                t = __dadd_rn(t, __dmul_rn(__dadd_rn(c0[j], c0[-j]), gk.k[j + GAUSS_R]));
            const float s = (float)t;
            if (s > best) { best = s; besti = i; }   // ascending i per lane -> keeps the first maximum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(FULL_MASK, best, o);
            const int oi = __shfl_xor_sync(FULL_MASK, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        // NaN rows: torch.argmax returns the first NaN; not reproduced (weights are finite for finite inputs)
        if (besti == 0x7fffffff) besti = 0;
        float ws = 0.f, dsum = 0.f;
        for (int i = lane; i < P; i += 32) {
            const float zi = sz[i];
            const float m = (b.lo < zi && zi < b.hi) ? 1.f : 0.f;
            ws += sw[i] * m;
            if (method != 2) dsum += sw[i] * zi;
        }
        ws = warp_sum(ws);
        if (method == 2) {
            // render.py:345-348
            const float cden = __fadd_rn(ws, epsilon);
            for (int i = lane; i < P; i += 32) {
                const float zi = sz[i];
                const float m = (b.lo < zi && zi < b.hi) ? 1.f : 0.f;
                dsum += __fdiv_rn(sw[i] * m, cden) * zi;
            }
        }
        dsum = warp_sum(dsum);
        if (lane == 0) {
            const float zp = sz[besti];
            peak_in[r] = (b.lo < zp && zp < b.hi) ? 1 : 0;
            wsum_child[r] = ws;
            depth[r] = dsum;
        }
        __syncwarp();
    }
    if (lane == 0) red[wib] = acc_op;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < wpb; ++k) t += red[k];
        atomicAdd(&sums[0], t);
    }
}

// Group winner (render.py:317-340).  The reference walks the rows sequentially, jumping over the followers of each
// head.  Parallel form for well-formed input (a head row carries its follower count > 0, a singleton 0, follower
// rows 0): (A) every head marks its followers as covered (bit 2); (B) every uncovered row acts as head / singleton
// and sets bit 1 on the winner; (C) the cover bit is cleared.
__global__ void k_select_cover(const int64_t* __restrict__ other, int64_t n, uint8_t* __restrict__ flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t oi = other[i];
    for (int64_t j = 1; j <= oi && i + j < n; ++j) flag[i + j] = 2;
}

__global__ void k_select_winner(const int64_t* __restrict__ other, const uint8_t* __restrict__ peak_in,
                                const float* __restrict__ wsum, int64_t n, volatile uint8_t* flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i] & 2) return;                   // follower: decided by its head
    const int64_t oi = other[i];
    if (oi < 0) return;
    int64_t k = oi;
    if (i + k >= n) k = n - 1 - i;             // the reference would raise IndexError on a truncated group
    int64_t win = i;
    if (k > 0 && !peak_in[i]) {
        bool found = false;
        for (int64_t j = 1; j <= k; ++j)
            if (peak_in[i + j]) { win = i + j; found = true; break; }
        if (!found)
            for (int64_t j = 1; j <= k; ++j)
                if (wsum[i + j] > wsum[win]) win = i + j;
    }
    flag[win] = (win == i) ? 1 : 3;            // single writer per byte: rows i+1..i+k belong to this head only
}

// (rows at or beyond *n_eff were never handed to the renderer by the batch driver: see k_eval_walk)
__global__ void k_select_clear(int64_t n, const int64_t* __restrict__ n_eff, uint8_t* __restrict__ flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (n_eff && i >= *n_eff) ? 0 : (flag[i] & 1);
}

// head_flag[i] = 1 for rows that start a candidate group (everything k_select_cover did not mark as a follower)
__global__ void k_group_heads(int64_t n, uint8_t* __restrict__ flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (flag[i] & 2) ? 0 : 1;
}

// mismatch += number of rows whose ray (columns 0..5 and the parent segment columns) differs from their group head's
__global__ void k_group_uniform(const float* __restrict__ rays, int ld, int64_t n, const int32_t* __restrict__ row_ray,
                                const int64_t* __restrict__ head_row, int c0, int c1, int* __restrict__ mismatch) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = rays + i * ld;
    const float* b = rays + head_row[row_ray[i]] * ld;
    bool same = a[c0] == b[c0] && a[c1] == b[c1];
#pragma unroll
    for (int k = 0; k < 6; ++k) same = same && a[k] == b[k];
    if (!same) atomicAdd(mismatch, 1);
}

// The batch walk of eval_kitti_render.py:979-1005 / :1111-1136 on the device (one thread: ~n / batch steps).  Batches are
// extended so that no candidate group is split (follower rows carry -1 in the tag column); results do not depend on the
// batch boundaries in eval mode (running statistics) EXCEPT for one quirk: when a batch ends exactly one row before the
// end, the loop stops (`if i == n - 1: break`) and that last row is never rendered.  out[0] = number of rows rendered.
__global__ void k_eval_walk(const float* __restrict__ rays, int ld, int tag_col, int64_t n, int64_t batch,
                            int64_t* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int64_t i = 0;
    while (i < n) {
        if (i == n - 1) break;
        if ((double)(i + batch) < (double)n - 0.5 * (double)batch) {
            int64_t extra = 0;
            while (rays[(i + batch + extra) * ld + tag_col] < -0.5f) {
                ++extra;
                if (i + batch + extra == n) break;
            }
            i = i + batch + extra;
        } else {
            i = n;
        }
    }
    out[0] = i;
}

__global__ void k_points(const float* __restrict__ rays, int ld, int64_t n, const float* __restrict__ depth,
                         float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = rays + i * ld;
    const float d = depth[i];
    out[3 * i + 0] = __fadd_rn(r[0], __fmul_rn(d, r[3]));
    out[3 * i + 1] = __fadd_rn(r[1], __fmul_rn(d, r[4]));
    out[3 * i + 2] = __fadd_rn(r[2], __fmul_rn(d, r[5]));
}

static int search_rows_impl(const float* p, const float* z, const float* rays, int ld, int64_t n, int P, int cnear_col,
                            int cfar_col, float epsilon, int method, const int32_t* row_ray, float* w, float* depth,
                            uint8_t* peak_in_child, float* wsum_child, double* sums, void* stream) {
    PCN_CHECK_ARG(n >= 0 && P >= 1 && cnear_col < ld && cfar_col < ld && sums, "search_rows: bad arguments");
    PCN_CHECK_ARG(row_ray || w, "search_rows: the per-row form needs the weights output");
    cudaStream_t st = (cudaStream_t)stream;
    PCN_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(double), st));
    if (n == 0) return 0;
    // scipy.ndimage._filters._gaussian_kernel1d(sigma=5, order=0, radius=int(4*5+0.5)=20): the 21 leading weights
    // (offsets -20..0) as produced by numpy (exp, pairwise sum, divide) -- hex literals so that no host libm
    // difference can change a bit; tests/test_oracle_golden.py::test_gaussian_filter_matches_scipy pins the oracle
    // restatement against SciPy and tests/test_gpu_search.py pins this table against the oracle.
    static const double kGauss[GAUSS_R + 1] = {
        0x1.c113e67a34f9ap-16, 0x1.e9d347af7ba2cp-15, 0x1.00a91aed84201p-13, 0x1.026ceaaef5d9cp-12,
        0x1.f3ffe5366298dp-12, 0x1.d0bb4c23b8d53p-11, 0x1.9f03a798bae24p-10, 0x1.64156b94ff939p-9,
        0x1.258a96c00a508p-8,  0x1.d0fdc1a91a71bp-8,  0x1.61d971cc0d07ep-7,  0x1.02b6d98acd25ap-6,
        0x1.6b7adf708e81bp-6,  0x1.eaa58a4ba7224p-6,  0x1.3e2acd55166dap-5,  0x1.8c75f2fc165cbp-5,
        0x1.daa6517492b60p-5,  0x1.10fd11517a7f6p-4,  0x1.2db2f1c27e704p-4,  0x1.405ae2f8a8257p-4,
        0x1.46d39dcd3d08cp-4};
    GaussK gk;
    for (int j = 0; j <= GAUSS_R; ++j) gk.k[j] = kGauss[j];
    const size_t per_warp = (size_t)(2 * P + 2 * GAUSS_R) * sizeof(double);    // (P + 2R) doubles + 2P floats
    int wpb = (int)(SRCH_MAX_SMEM / per_warp);
    if (wpb < 1) {
        pcn_set_error("search_rows: %d samples per row exceed the shared-memory budget", P);
        return PCNERF_ERR_UNSUPPORTED;
    }
    if (wpb > 8) wpb = 8;
    const size_t smem = per_warp * wpb;
    if (smem > 48 * 1024)
        PCN_CUDA(cudaFuncSetAttribute(k_search_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = pcn_cdiv(n, wpb);
    const int64_t cap = (int64_t)PCN_SM_COUNT * 16;
    if (grid > cap) grid = cap;
    PcnScope ps(PCN_K_SEARCH, st, (double)n * (8.0 + 12.0 * P + 13.0));
    k_search_rows<<<(int)grid, wpb * 32, smem, st>>>(p, z, rays, ld, n, P, cnear_col, cfar_col, epsilon, method, gk, row_ray,
                                                    w, depth, peak_in_child, wsum_child, sums);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_search_rows(const float* p, const float* z, const float* rays, int ld, int64_t n, int P,
                                  int cnear_col, int cfar_col, float epsilon, int method, float* w, float* depth,
                                  uint8_t* peak_in_child, float* wsum_child, double* sums, void* stream) {
    return search_rows_impl(p, z, rays, ld, n, P, cnear_col, cfar_col, epsilon, method, nullptr, w, depth, peak_in_child,
                            wsum_child, sums, stream);
}

extern "C" int pcnerf_search_rows_grouped(const float* p_ray, const float* z_ray, const float* rays, int ld, int64_t n_rows,
                                          int P, int cnear_col, int cfar_col, float epsilon, int method,
                                          const int32_t* row_ray, float* w_ray, float* depth, uint8_t* peak_in_child,
                                          float* wsum_child, double* sums, void* stream) {
    PCN_CHECK_ARG(row_ray, "search_rows_grouped: null row -> ray map");
    return search_rows_impl(p_ray, z_ray, rays, ld, n_rows, P, cnear_col, cfar_col, epsilon, method, row_ray, w_ray, depth,
                            peak_in_child, wsum_child, sums, stream);
}

extern "C" int pcnerf_group_heads(const int64_t* other, int64_t n, uint8_t* head_flag, void* stream) {
    PCN_CHECK_ARG(n >= 0 && (n == 0 || (other && head_flag)), "group_heads: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return 0;
    PCN_CUDA(cudaMemsetAsync(head_flag, 0, (size_t)n, st));
    const int g = (int)pcn_cdiv(n, 256);
    PcnScope ps(PCN_K_SEARCH, st, (double)n * 10.0, 2);
    k_select_cover<<<g, 256, 0, st>>>(other, n, head_flag);
    k_group_heads<<<g, 256, 0, st>>>(n, head_flag);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_group_uniform(const float* rays, int ld, int64_t n, const int32_t* row_ray, const int64_t* head_row,
                                    int pnear_col, int pfar_col, int* mismatch, void* stream) {
    PCN_CHECK_ARG(n >= 0 && ld >= 6 && pnear_col < ld && pfar_col < ld && mismatch, "group_uniform: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCN_CUDA(cudaMemsetAsync(mismatch, 0, sizeof(int), st));
    if (n == 0) return 0;
    PcnScope ps(PCN_K_SEARCH, st, (double)n * 68.0);
    k_group_uniform<<<(int)pcn_cdiv(n, 256), 256, 0, st>>>(rays, ld, n, row_ray, head_row, pnear_col, pfar_col, mismatch);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_eval_rows_rendered(const float* rays, int ld, int tag_col, int64_t n, int64_t batch_size_set,
                                         int64_t* out_n, void* stream) {
    PCN_CHECK_ARG(n >= 0 && batch_size_set >= 1 && tag_col < ld && out_n, "eval_rows_rendered: bad arguments");
    PcnScope ps(PCN_K_SEARCH, (cudaStream_t)stream, (double)n * 4.0 / (double)batch_size_set);
    k_eval_walk<<<1, 32, 0, (cudaStream_t)stream>>>(rays, ld, tag_col, n, batch_size_set, out_n);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_search_select(const int64_t* other, const uint8_t* peak_in_child, const float* wsum_child,
                                    int64_t n, const int64_t* n_rendered, uint8_t* out_flag, void* stream) {
    PCN_CHECK_ARG(n >= 0, "search_select: bad size");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return 0;
    PCN_CUDA(cudaMemsetAsync(out_flag, 0, (size_t)n, st));
    const int g = (int)pcn_cdiv(n, 256);
    PcnScope ps(PCN_K_SEARCH, st, (double)n * 15.0, 3);
    k_select_cover<<<g, 256, 0, st>>>(other, n, out_flag);
    k_select_winner<<<g, 256, 0, st>>>(other, peak_in_child, wsum_child, n, out_flag);
    k_select_clear<<<g, 256, 0, st>>>(n, n_rendered, out_flag);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_points(const float* rays, int ld, int64_t n, const float* depth, float* out_xyz, void* stream) {
    PCN_CHECK_ARG(n >= 0 && ld >= 6, "points: bad arguments");
    if (n == 0) return 0;
    PcnScope ps(PCN_K_SEARCH, (cudaStream_t)stream, (double)n * 40.0);
    k_points<<<(int)pcn_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(rays, ld, n, depth, out_xyz);
    PCN_LAUNCH_CHECK();
    return 0;
}
