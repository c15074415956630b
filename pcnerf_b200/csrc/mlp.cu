// K3 (fp32 path): the occupancy MLP of nof/networks/models.py:125-203 as actually constructed --
//   Linear(63,256) BN  [Linear(256,256) BN]x3  cat(x, .)  Linear(319,256) BN  [Linear(256,256) BN]x3  Linear(256,1) Sigmoid
// with every LeakyReLU(negative_slope=True==1.0) an identity (models.py:152,172).
//
// Design: BN(l) is an affine map y = a*h + s per feature, so it is folded into Linear(l+1):
//   h_{l+1} = (W_{l+1} diag(a_l)) h_l + (b_{l+1} + W_{l+1} s_l).
// Training mode needs the batch statistics of h_l first: each layer is one GEMM whose epilogue accumulates the
// per-feature sum / sum of squares (fp32 per tile, fp64 across tiles), followed by a tiny fold kernel.
// Only the pre-BN activations h_1..h_8 are stored (backward needs them); normalised activations, the identity
// "activation" copies and the skip concat of the eager reference never touch HBM.
//
// This file is the CUDA-core fp32 path (precision 0, the 1e-5 parity gate).  mlp_tc.cu holds the bf16 tcgen05 path.
#include "common.cuh"
#include "mlp_layout.h"

// ---------------------------------------------------------------------------------------------------------------
// fp32 GEMM  C[m,n] = sum_k A(m,k) B(n,k)   (128x64x16 tiles, 8x4 micro-tiles, register-prefetched double buffer)
// ---------------------------------------------------------------------------------------------------------------

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_SPLITK = 2 };

struct GemmArgs {
    const float* A1; int lda1; int K1;     // A segment 1 covers k in [0,K1); non-transposed A only
    const float* A2; int lda2;             // A segment 2 covers k in [K1,K)
    const float* B; int ldb;
    float* C; int ldc;
    int M, N, K;
    const float* bias;                     // EPI_FWD: + bias[n]
    const float* E; int lde;               // EPI_DGRAD: second statistic is sum_m C[m,n]*E[m,n]
    double* stat0; double* stat1;          // per-column accumulators (EPI_FWD / EPI_DGRAD)
    int k_per_split;                       // EPI_SPLITK: partial C for split z goes to C + z*M*ldc
};

#define GBM 128
#define GBN 64
#define GBK 16
#define AS_LD (GBM + 4)
#define BS_LD (GBN + 4)

template <bool AT, bool BT, int EPI>
__global__ void __launch_bounds__(256) k_gemm(GemmArgs g) {
    __shared__ __align__(16) float As[2][GBK][AS_LD];
    __shared__ __align__(16) float Bs[2][GBK][BS_LD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    int kbeg = 0, kend = g.K;
    if (EPI == EPI_SPLITK) {
        kbeg = blockIdx.z * g.k_per_split;
        kend = min(g.K, kbeg + g.k_per_split);
    }
    float4 ra[2], rb;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    auto load_tile = [&](int k0) {
        if (AT) {
            // A(m,k) = A1[k*lda + m]: 16 k-rows x 128 m, float4 along m
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = k0 + (tid >> 5) + h * 8, m = m0 + (tid & 31) * 4;
                ra[h] = (k < kend) ? *reinterpret_cast<const float4*>(g.A1 + (size_t)k * g.lda1 + m)
                                   : make_float4(0, 0, 0, 0);
            }
        } else {
            const float* Ap = g.A1; int lda = g.lda1; int kk = k0;
            if (k0 >= g.K1) { Ap = g.A2; lda = g.lda2; kk = k0 - g.K1; }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = m0 + (tid >> 2) + h * 64, kq = (tid & 3) * 4;
                ra[h] = (m < g.M) ? *reinterpret_cast<const float4*>(Ap + (size_t)m * lda + kk + kq)
                                  : make_float4(0, 0, 0, 0);
            }
        }
        if (BT) {
            // B(n,k) = B[k*ldb + n]: 16 k-rows x 64 n
            const int k = k0 + (tid >> 4), n = n0 + (tid & 15) * 4;
            rb = (k < kend) ? *reinterpret_cast<const float4*>(g.B + (size_t)k * g.ldb + n) : make_float4(0, 0, 0, 0);
        } else {
            const int n = n0 + (tid >> 2), kq = (tid & 3) * 4;
            rb = *reinterpret_cast<const float4*>(g.B + (size_t)n * g.ldb + k0 + kq);
        }
    };
    auto store_tile = [&](int buf) {
        if (AT) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                *reinterpret_cast<float4*>(&As[buf][(tid >> 5) + h * 8][(tid & 31) * 4]) = ra[h];
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = (tid >> 2) + h * 64, kq = (tid & 3) * 4;
                As[buf][kq + 0][m] = ra[h].x; As[buf][kq + 1][m] = ra[h].y;
                As[buf][kq + 2][m] = ra[h].z; As[buf][kq + 3][m] = ra[h].w;
            }
        }
        if (BT) {
            *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = rb;
        } else {
            const int n = tid >> 2, kq = (tid & 3) * 4;
            Bs[buf][kq + 0][n] = rb.x; Bs[buf][kq + 1][n] = rb.y; Bs[buf][kq + 2][n] = rb.z; Bs[buf][kq + 3][n] = rb.w;
        }
    };

    int buf = 0;
    if (kbeg < kend) {
        load_tile(kbeg);
        store_tile(0);
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += GBK) {
        const bool more = k0 + GBK < kend;
        if (more) load_tile(k0 + GBK);
#pragma unroll
        for (int kk = 0; kk < GBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ---- epilogue
    const int n = n0 + tx * 4;
    if (EPI == EPI_SPLITK) {
        float* Cp = g.C + (size_t)blockIdx.z * g.M * g.ldc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + ty * 8 + i;
            if (m < g.M)
                *reinterpret_cast<float4*>(Cp + (size_t)m * g.ldc + n) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        return;
    }
    float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    float bv[4] = {0, 0, 0, 0};
    if (EPI == EPI_FWD) { bv[0] = g.bias[n]; bv[1] = g.bias[n + 1]; bv[2] = g.bias[n + 2]; bv[3] = g.bias[n + 3]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m < g.M) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bv[j];
            *reinterpret_cast<float4*>(g.C + (size_t)m * g.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
            float e[4] = {v[0], v[1], v[2], v[3]};
            if (EPI == EPI_DGRAD) {
                const float4 ev = *reinterpret_cast<const float4*>(g.E + (size_t)m * g.lde + n);
                e[0] = ev.x; e[1] = ev.y; e[2] = ev.z; e[3] = ev.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { s0[j] += v[j]; s1[j] = fmaf(v[j], e[j], s1[j]); }
        }
    }
    // reduce the 16 ty-partials per column through shared memory, then one fp64 atomic per column per CTA
    float* red0 = &As[0][0][0];            // 16 x 64
    float* red1 = red0 + 16 * 64;
#pragma unroll
    for (int j = 0; j < 4; ++j) { red0[ty * 64 + tx * 4 + j] = s0[j]; red1[ty * 64 + tx * 4 + j] = s1[j]; }
    __syncthreads();
    if (tid < 128) {
        const int c = tid & 63;
        const float* rp = tid < 64 ? red0 : red1;
        double t = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) t += (double)rp[k * 64 + c];
        atomicAdd((tid < 64 ? g.stat0 : g.stat1) + n0 + c, t);
    }
}

template <bool AT, bool BT, int EPI>
static void launch_gemm(const GemmArgs& g, int splits, cudaStream_t st) {
    dim3 grid((unsigned)pcn_cdiv(g.M, GBM), (unsigned)(g.N / GBN), (unsigned)splits);
    PcnScope ps(EPI == EPI_FWD ? PCN_K_GEMM_FWD : (EPI == EPI_DGRAD ? PCN_K_GEMM_DGRAD : PCN_K_GEMM_WGRAD), st,
                2.0 * (double)g.M * (double)g.N * (double)g.K);
    k_gemm<AT, BT, EPI><<<grid, 256, 0, st>>>(g);
}

#include "mlp_small.cuh"

// ---------------------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------------------

static int check_common(const pcnerf_mlp_params* P, const void* enc, int64_t rows, const char* who) {
    PCN_CHECK_ARG(P && enc, "%s: null params / enc", who);
    PCN_CHECK_ARG(rows >= 1, "%s: rows must be >= 1", who);
    PCN_CHECK_ARG(rows < (1ll << 31) / 320, "%s: chunk of %lld rows is too large (use a smaller chunk)", who, (long long)rows);
    if (P->training && rows == 1) {
        pcn_set_error("Expected more than 1 value per channel when training, got input size [1, 256]");
        return PCNERF_ERR_ARG;
    }
    return 0;
}

int mlp_tc_forward(const pcnerf_mlp_params*, const void*, int64_t, float*, void*, size_t, void*, size_t, cudaStream_t);
int mlp_tc_backward(const pcnerf_mlp_params*, const pcnerf_mlp_grads*, const void*, int64_t, const float*, const float*,
                    void*, size_t, void*, size_t, cudaStream_t);

extern "C" size_t pcnerf_mlp_saved_bytes(int64_t rows, int precision) { return MlpLayout(rows, precision).saved_bytes; }
extern "C" size_t pcnerf_mlp_scratch_bytes(int64_t rows, int precision) { return MlpLayout(rows, precision).scratch_bytes; }

static void prep_weights(const pcnerf_mlp_params* P, const MlpLayout& L, char* scratch, cudaStream_t st) {
    PrepArgs pa;
    for (int l = 0; l < 8; ++l) { pa.W[l] = P->W[l]; pa.Wp[l] = L.Wp(scratch, l); }
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_prep_weights<<<dim3(64, 8), 256, 0, st>>>(pa));
}

extern "C" int pcnerf_mlp_forward(const pcnerf_mlp_params* P, const void* enc, int64_t rows, float* out_p, void* saved,
                                  size_t saved_bytes, void* scratch_v, size_t scratch_bytes, void* stream) {
    int rc = check_common(P, enc, rows, "mlp_forward");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const MlpLayout L(rows, P->precision);
    // (the fused eval kernel of precision 1 keeps every activation on chip and saves nothing)
    const bool no_saved = P->precision == 1 && !P->training && pcnerf_tc_get_fused_eval();
    PCN_CHECK_ARG(no_saved || (saved && saved_bytes >= L.saved_bytes), "mlp_forward: saved buffer too small (%zu < %zu)", saved_bytes, L.saved_bytes);
    // (... and uses only the rows-independent head of the scratch layout: weight copies, folded biases)
    const size_t need_scratch = no_saved ? MlpLayout(1, 1).scratch_bytes : L.scratch_bytes;
    PCN_CHECK_ARG(scratch_v && scratch_bytes >= need_scratch, "mlp_forward: scratch too small (%zu < %zu)", scratch_bytes, need_scratch);
    if (P->precision == 1) return mlp_tc_forward(P, enc, rows, out_p, saved, saved_bytes, scratch_v, scratch_bytes, st);
    PCN_CHECK_ARG(P->precision == 0, "mlp_forward: precision must be 0 (fp32) or 1 (bf16 tcgen05)");
    char* scratch = (char*)scratch_v;
    char* sv = (char*)saved;
    PCN_CUDA(cudaMemsetAsync(L.dstat(scratch, 0), 0, sizeof(double) * 8 * 512, st));
    if (!P->prepared) prep_weights(P, L, scratch, st);
    const float* encf = (const float*)enc;
    for (int l = 0; l < 8; ++l) {
        GemmArgs g = {};
        float* Hl = L.H(sv, l);
        if (l == 0) { g.A1 = encf; g.lda1 = 64; g.K1 = 64; g.K = 64; }
        else if (l == 4) { g.A1 = encf; g.lda1 = 64; g.K1 = 64; g.A2 = L.H(sv, 3); g.lda2 = 256; g.K = 320; }
        else { g.A1 = L.H(sv, l - 1); g.lda1 = 256; g.K1 = 256; g.K = 256; }
        g.B = l == 0 ? L.Wp(scratch, 0) : L.Wf(scratch, l);
        g.ldb = mlp_kpad(l);
        g.bias = l == 0 ? P->b[0] : L.bf(scratch, l);
        g.C = Hl; g.ldc = 256; g.M = (int)rows; g.N = 256;
        g.stat0 = L.dstat(scratch, l); g.stat1 = g.stat0 + 256;
        launch_gemm<false, false, EPI_FWD>(g, 1, st);
        const bool last = l == 7;
        PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_bn_fold<<<last ? 1 : 256, 256, 0, st>>>(
            l, P->training, rows, g.stat0, g.stat1, P->gamma[l], P->beta[l], P->running_mean[l], P->running_var[l],
            P->num_batches_tracked[l], P->momentum, P->eps, L.stats(sv, l), last ? P->W[8] : L.Wp(scratch, l + 1),
            last ? P->b[8] : P->b[l + 1], last ? L.wout_f(scratch) : L.Wf(scratch, l + 1),
            last ? L.wout_f(scratch) + 256 : L.bf(scratch, l + 1), nullptr));
    }
    int64_t blocks = pcn_cdiv(rows, 8);
    if (blocks > PCN_SM_COUNT * 16) blocks = PCN_SM_COUNT * 16;
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_logit_sigmoid<float><<<(int)blocks, 256, 0, st>>>(L.H(sv, 7), rows, L.wout_f(scratch), L.wout_f(scratch) + 256, out_p));
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_mlp_backward(const pcnerf_mlp_params* P, const pcnerf_mlp_grads* G, const void* enc, int64_t rows,
                                   const float* out_p, const float* grad_p, void* saved, size_t saved_bytes,
                                   void* scratch_v, size_t scratch_bytes, void* stream) {
    int rc = check_common(P, enc, rows, "mlp_backward");
    if (rc) return rc;
    PCN_CHECK_ARG(G && out_p && grad_p, "mlp_backward: null grads / out_p / grad_p");
    PCN_CHECK_ARG(P->training, "mlp_backward: only training-mode (batch-statistics) backward is implemented");
    cudaStream_t st = (cudaStream_t)stream;
    const MlpLayout L(rows, P->precision);
    PCN_CHECK_ARG(saved && saved_bytes >= L.saved_bytes, "mlp_backward: saved buffer too small");
    PCN_CHECK_ARG(scratch_v && scratch_bytes >= L.scratch_bytes, "mlp_backward: scratch too small");
    if (P->precision == 1)
        return mlp_tc_backward(P, G, enc, rows, out_p, grad_p, saved, saved_bytes, scratch_v, scratch_bytes, st);
    PCN_CHECK_ARG(P->precision == 0, "mlp_backward: precision must be 0 (fp32) or 1 (bf16 tcgen05)");
    char* scratch = (char*)scratch_v;
    char* sv = (char*)saved;
    const float* encf = (const float*)enc;
    PCN_CUDA(cudaMemsetAsync(L.dstat(scratch, 0), 0, sizeof(double) * L.n_dstat, st));
    if (!P->prepared) prep_weights(P, L, scratch, st);
    double* acc_out = L.dstat(scratch, 8);            // 257 (+pad) doubles
    float* gvec = L.gvec(scratch);
    float* coef = L.coef(scratch);
    float* Gb[2] = {L.G(scratch, 0), L.G(scratch, 1)};
    const int strips = (int)(pcn_cdiv(rows, STRIP) < 4 * PCN_SM_COUNT ? pcn_cdiv(rows, STRIP) : 4 * PCN_SM_COUNT);

    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_out_bwd_reduce<float><<<strips, 256, 0, st>>>(grad_p, out_p, L.H(sv, 7), rows, gvec, acc_out));
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_out_bwd_finalize<<<1, 256, 0, st>>>(acc_out, rows, P->W[8], L.stats(sv, 7), G->dW[8], G->db[8], G->dgamma[7],
                                          G->dbeta[7], coef));
    int cur = 0;
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_bn_bwd_apply<true, float, float><<<strips, 256, 0, st>>>(gvec, Gb[cur], L.H(sv, 7), rows, coef, L.stats(sv, 7),
                                                 L.colsum(scratch, 7)));
    // split-K factor for the weight-gradient GEMMs
    int splits = (int)pcn_cdiv(rows, 2048);
    if (splits > MLP_MAX_SPLITS) splits = MLP_MAX_SPLITS;
    if (splits < 1) splits = 1;
    int kps = (int)pcn_cdiv(pcn_cdiv(rows, splits), GBK) * GBK;
    splits = (int)pcn_cdiv(rows, kps);
    for (int l = 7; l >= 0; --l) {
        const float* DH = Gb[cur];
        const int kpad = mlp_kpad(l);
        float* part = L.partial(scratch);
        // ---- weight gradient  dWraw[o, c] = sum_r DH[r,o] * U[r,c]
        {
            GemmArgs g = {};
            g.A1 = DH; g.lda1 = 256; g.M = 256; g.K = (int)rows; g.k_per_split = kps;
            g.C = part; g.ldc = kpad;
            if (l == 0 || l == 4) {           // encoding columns
                g.B = encf; g.ldb = 64; g.N = 64;
                launch_gemm<true, true, EPI_SPLITK>(g, splits, st);
            }
            if (l != 0) {                     // hidden columns
                g.B = L.H(sv, l - 1); g.ldb = 256; g.N = 256;
                g.C = part + (l == 4 ? 64 : 0);
                launch_gemm<true, true, EPI_SPLITK>(g, splits, st);
            }
            PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_wgrad_finalize<<<128, 256, 0, st>>>(l, part, splits, L.colsum(scratch, l),
                                                   l == 0 ? nullptr : L.stats(sv, l - 1), G->dW[l], G->db[l]));
        }
        if (l == 0) break;
        // ---- data gradient  Gy_{l-1} = DH_l . W_l[:, hidden cols], with the BN(l-1) reductions in the epilogue
        {
            GemmArgs g = {};
            g.A1 = DH; g.lda1 = 256; g.K1 = 256; g.K = 256; g.M = (int)rows; g.N = 256;
            g.B = L.Wp(scratch, l) + (l == 4 ? 64 : 0); g.ldb = kpad;
            g.C = Gb[cur ^ 1]; g.ldc = 256;
            g.E = L.H(sv, l - 1); g.lde = 256;
            g.stat0 = L.dstat(scratch, 9 + (l - 1)); g.stat1 = g.stat0 + 256;
            launch_gemm<false, true, EPI_DGRAD>(g, 1, st);
            PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_bn_bwd_coef<<<1, 256, 0, st>>>(g.stat0, g.stat1, rows, L.stats(sv, l - 1), G->dgamma[l - 1],
                                             G->dbeta[l - 1], coef));
            cur ^= 1;
            PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_bn_bwd_apply<false, float, float><<<strips, 256, 0, st>>>(nullptr, Gb[cur], L.H(sv, l - 1), rows, coef,
                                                          L.stats(sv, l - 1), L.colsum(scratch, l - 1)));
        }
    }
    PCN_LAUNCH_CHECK();
    return 0;
}
