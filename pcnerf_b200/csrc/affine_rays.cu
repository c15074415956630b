// K3' on rays -- the closed-form ("affine", precision 2) occupancy MLP of a TRAINING pass without an encoding tensor and
// without torch algebra: everything between (rays, z) and p_occ, forward and backward, in this file.
//
// affine.cu explains the formulation: as the reference builds it (nof/networks/models.py:152,172: identity activations) the
// logit of one BatchNorm batch (= one `chunk` of nof/render.py:47-49) is exactly affine in the encoding x of the sample,
// and the affine map depends on the batch only through the first two moments of x.  Two things change here:
//   * rows are (ray, depth) pairs: a kernel that needs x re-derives it (encode.cuh: ~25 instructions per sin/cos pair)
//     instead of reading 256 bytes per row -- the 2.1 GB encoding tensor of a C2 step and its four passes over HBM are gone;
//   * the parameter-sized algebra runs in float64 in repo kernels, in HOMOGENEOUS coordinates: x_63 := 1 (column 63 of the
//     encoding is zero padding), so that every layer is one 256 x 64 matrix per chunk,  H_l = A_l x,  and
//         mean_l = A_l m,  var_l = diag(A_l C A_l^T),  a_l = gamma_l / sqrt(var_l + eps),  s_l = beta_l - a_l mean_l,
//         Ab_l = diag(a_l) A_l + s_l e_63^T            (BN_l o Linear_l as a map of x)
//         A_{l+1} = W_{l+1} Ab_l + b_{l+1} e_63^T      (+ [W_x | 0] for the skip layer, models.py:192-194)
//         alpha = w_out Ab_7 + b_out e_63^T,           logit = alpha . x
//     with m = E[x] (m_63 = 1) and C = Cov[x] (row / column 63 zero).  The backward is the hand-derived reverse of exactly
//     these lines (k_aff_layer_bwd); the three GEMM shapes per layer (A_{l+1} = W Ab_l, G_{l} = W^T G_{l+1},
//     dW = sum_chunks G Ab^T) go through one tiled DFMA kernel with the chunks concatenated along N (or K).
// Data-sized kernels: k_affine_moments_rays (second moments: the only O(rows x 64 x 64) work, upper triangle only),
// k_affine_apply_rays, k_affine_grad_rays.  Test infrastructure compares against the torch formulation of the same algebra
// (NOF._affine_coeffs) and the fp32 engine (tests/test_gpu_affine.py).
#include <cstring>
#include "common.cuh"
#include "encode.cuh"

#define AFR_PARTS 64            // most CTAs (partial results) per chunk of the data-sized kernels (buffer sizing)
#define AFR_T 120               // rows per shared-memory sub-tile of the moments kernel (two buffers); 8 rows per row group
#define AFR_GROUPS 3            // consumer row groups (rows = g mod 3), 36 tile owners each
#define AFR_PROD 128            // producer threads of the moments kernel: warps 0..3 encode one row each per sub-tile (120 used)
#define AFR_TILES 36            // 8 x 8 tiles of the upper triangle of the 64 x 64 moment matrix (8 * 9 / 2)
#define AFR_THREADS 256         // 8 warps: 4 producers + 4 consumers (3 row groups x 36 tile owners, 20 idle lanes)
#define AFR_FLUSH 4             // sub-tiles (= 160 rows per tile owner) accumulated in fp32 before the fold into fp64 (power of two)
#define AFR_MOM_SMEM (2 * AFR_T * 64 * 4 + AFR_GROUPS * AFR_TILES * 64 * 4 + AFR_TILES * 64 * 8)     // dynamic shared memory of the moments kernel: two sub-tile buffers, fp32 scratch, fp64 accumulators
#define AFR_SPLITS 6            // split-K factor of the weight-gradient GEMM (64 tiles x 6 = 384 blocks: one wave of 3 per SM)
#define AFR_SPLITS_BATCHED 2    // ... of the batched launch over the seven layers of a pass (k_aff_wgrad_batched)

struct RayRows {
    const float* rays;          // (n_rays, ld): origin in columns 0..2, direction in 3..5
    int ld;
    const float* z;             // (n_rays, S) depths, row r = ray r / S, sample r % S
    int S;
};

// rays_o + rays_d * z (nof/render.py:458: mul then add, no FMA) -- the position K2 encodes
__device__ __forceinline__ void ray_row_pos(const RayRows& s, int64_t r, float& x0, float& x1, float& x2) {
    const uint32_t ray = (uint32_t)r / (uint32_t)s.S;
    const float* q = s.rays + (int64_t)ray * s.ld;
    const float zz = __ldg(s.z + r);
    x0 = __fadd_rn(__ldg(q), __fmul_rn(__ldg(q + 3), zz));
    x1 = __fadd_rn(__ldg(q + 1), __fmul_rn(__ldg(q + 4), zz));
    x2 = __fadd_rn(__ldg(q + 2), __fmul_rn(__ldg(q + 5), zz));
}

__device__ __forceinline__ void part_range(int64_t rows, int64_t chunk, int ci, int pi, int64_t& c_beg, int64_t& r_beg,
                                           int64_t& r_end) {
    c_beg = (int64_t)ci * chunk;
    const int parts = (int)gridDim.x;
    const int64_t c_end = min(rows, c_beg + chunk), per = (c_end - c_beg + parts - 1) / parts;
    r_beg = min(c_end, c_beg + pi * per);
    r_end = min(c_end, r_beg + per);
}

// Sub-tile layout of the moments kernel: row-major [AFR_T][64] floats; the first four columns of the 8-column group tg sit in
// 16-byte chunk (tg ^ key) of the row's first 128 bytes, the last four in the same chunk of the second 128 bytes, key =
// row & 7.  The 8 rows a quarter-warp writes with one STS.128 land in 8 different bank groups, and the quarter-warp of tile
// owners that reads the first (second) halves of 8 different column groups of one row reads 128 contiguous bytes.
__device__ __forceinline__ int afr_chunk(int row, int tg, int h) { return row * 64 + (h << 5) + ((tg ^ (row & 7)) << 2); }

// d.lo += a * b.lo, d.hi += a * b.hi: the packed FFMA2 with a scalar multiplicand (ptxas folds the {a, a} pair into the
// instruction's .F32 broadcast operand)
__device__ __forceinline__ void ffma2s(unsigned long long& d, float a, unsigned long long b) {
    unsigned long long aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(aa), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

// named barriers (immediates, so that the kernel reserves 6 hardware barriers and not all 16): FULL + b / EMPTY + b hand
// buffer b between producers and consumers (all AFR_THREADS take part), CONS synchronises the consumer warps only
template <int ID, int N>
__device__ __forceinline__ void bar_sync_id() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N>
__device__ __forceinline__ void bar_arrive_id() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID0>
__device__ __forceinline__ void bar_sync(int b) { if (b) bar_sync_id<ID0 + 1, AFR_THREADS>(); else bar_sync_id<ID0, AFR_THREADS>(); }
template <int ID0>
__device__ __forceinline__ void bar_arrive(int b) { if (b) bar_arrive_id<ID0 + 1, AFR_THREADS>(); else bar_arrive_id<ID0, AFR_THREADS>(); }

// ---------------------------------------------------------------------------------------------------------------
// data-sized kernels
// ---------------------------------------------------------------------------------------------------------------

// part[(ci, pi)][64][64] (only the 8 x 8 tiles with tile row <= tile column are written):
//   sum over the part's rows of y y^T,  y = (x_0 - s_0, ..., x_62 - s_62, 1),  s = encoding of the chunk's first row.
// Column 63 of the result holds the first moments, element (63, 63) the row count.
//
// The kernel is bound by shared-memory wavefronts (ncu on the 4 x 4-tile form: 89 % of the LSU-shared peak, every 128-bit
// load of a warp costs four wavefronts whatever its lanes share), so the outer products use the largest register tile that
// fits: 8 x 8 per thread, 36 tiles for the upper triangle, accumulated with packed FFMA2 (scalar x pair) -- 4 loads and 32
// FFMA2 per row and thread.  Warp-specialised: warps 0..3 (one row per thread) encode sub-tile t + 1 into one buffer while
// warps 4..7 accumulate sub-tile t from the other: three row groups (rows = g mod 3) x 36 tile owners.  fp32 accumulation
// over AFR_FLUSH sub-tiles (160 rows per thread), then the groups' partial tiles are added into fp64 accumulators in shared
// memory.
__global__ void __launch_bounds__(AFR_THREADS, 2) k_affine_moments_rays(RayRows src, int64_t rows, int64_t chunk,
                                                                        double* __restrict__ part,
                                                                        double* __restrict__ shift_out, int dbg) {
    extern __shared__ __align__(16) float xs[];             // 2 x [AFR_T][64] | scratch [3][36][64] fp32 | acc [36][64] fp64
    __shared__ float shift[64];
    float* scratch = xs + 2 * AFR_T * 64;
    double* accs = reinterpret_cast<double*>(scratch + AFR_GROUPS * AFR_TILES * 64);
    const int ci = blockIdx.y, pi = blockIdx.x, tid = threadIdx.x;
    int64_t c_beg, r_beg, r_end;
    part_range(rows, chunk, ci, pi, c_beg, r_beg, r_end);
    // shift: the position of the chunk's first row for columns 0..2 (|x| reaches tens of metres); the sin / cos columns are
    // bounded with means near zero and are accumulated unshifted
    if (tid < 64) shift[tid] = 0.f;
    __syncthreads();
    if (tid == 0) ray_row_pos(src, c_beg, shift[0], shift[1], shift[2]);
    for (int e = tid; e < AFR_TILES * 64; e += AFR_THREADS) accs[e] = 0.0;
    __syncthreads();
    if (pi == 0 && tid < 64) shift_out[ci * 64 + tid] = (double)shift[tid];
    const int nt = (int)((r_end - r_beg + AFR_T - 1) / AFR_T);
    enum { FULL = 1, EMPTY = 3, CONS = 5 };

    if (tid < AFR_PROD) {
        const float s0 = shift[0], s1 = shift[1], s2 = shift[2];
        for (int t = 0; t < nt; ++t) {
            const int b = t & 1;
            if (t >= 2) bar_sync<EMPTY>(b);
            float* buf = xs + b * (AFR_T * 64);
            const int64_t r = r_beg + (int64_t)t * AFR_T + tid;
            if (tid < AFR_T) {
                float e[64];
                if (r < r_end && dbg) {                  // timing experiment (PCNERF_AFF_DEBUG=1, wrong results): free producers
                    float x0, x1, x2;
                    ray_row_pos(src, r, x0, x1, x2);
#pragma unroll
                    for (int col = 0; col < 64; ++col) e[col] = x0 * (float)(col + 1);
                } else if (r < r_end) {
                    float x0, x1, x2;
                    ray_row_pos(src, r, x0, x1, x2);
                    enc_visit_poly(x0, x1, x2, [&](int col, float v) { e[col] = v; });
                    e[0] -= s0; e[1] -= s1; e[2] -= s2;
                    e[63] = 1.f;
                } else {
#pragma unroll
                    for (int col = 0; col < 64; ++col) e[col] = 0.f;
                }
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    *reinterpret_cast<float4*>(buf + afr_chunk(tid, c >> 1, c & 1)) =
                        make_float4(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]);
            }
            __threadfence_block();
            bar_arrive<FULL>(b);
        }
        return;
    }
    const int cid = tid - AFR_PROD;                  // 0 .. 127
    const bool worker = cid < AFR_GROUPS * AFR_TILES;
    const int g = worker ? cid / AFR_TILES : 0, id = worker ? cid - g * AFR_TILES : 0;
    int ti = 0, tj = 0;
    {                                               // id -> (ti, tj), ti <= tj < 8, rows of the upper triangle one after the other
        int rem = id;
        while (rem >= 8 - ti) { rem -= 8 - ti; ++ti; }
        tj = ti + rem;
    }
    unsigned long long f[8][4];                     // f[i][jp] = (tile[i][2 jp], tile[i][2 jp + 1])
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) f[i][j] = 0ull;
    // this group's rows are g, g + 3, g + 6, ...: 8 of them span 24 rows, after which the swizzle keys (row & 7) repeat --
    // the offsets of column groups ti / tj in those 8 rows are loop invariants (second halves: + 32 floats)
    int oa[8], ob[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        oa[u] = afr_chunk(g + AFR_GROUPS * u, ti, 0);
        ob[u] = afr_chunk(g + AFR_GROUPS * u, tj, 0);
    }
    auto flush = [&]() {
        // the groups' fp32 partial tiles -> scratch -> fp64 accumulators (the 108 tile owners fold 64 / 3 elements each)
        if (worker) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const float2 p0 = unpack2(f[i][j]), p1 = unpack2(f[i][j + 1]);
                    *reinterpret_cast<float4*>(scratch + (g * AFR_TILES + id) * 64 + i * 8 + 2 * j) = make_float4(p0.x, p0.y, p1.x, p1.y);
                    f[i][j] = 0ull; f[i][j + 1] = 0ull;
                }
            }
        }
        bar_sync_id<CONS, AFR_THREADS - AFR_PROD>();
        for (int e = cid; e < AFR_TILES * 64; e += AFR_THREADS - AFR_PROD) {
            double v = 0.0;
#pragma unroll
            for (int gg = 0; gg < AFR_GROUPS; ++gg) v += (double)scratch[gg * AFR_TILES * 64 + e];
            accs[e] += v;
        }
        bar_sync_id<CONS, AFR_THREADS - AFR_PROD>();
    };
    for (int t = 0; t < nt; ++t) {
        const int b = t & 1;
        bar_sync<FULL>(b);
        if (worker) {
            const float* buf = xs + b * (AFR_T * 64);
            for (int rb = 0; rb < AFR_T; rb += 8 * AFR_GROUPS) {
                const float* base = buf + rb * 64;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 a0 = *reinterpret_cast<const float4*>(base + oa[u]);
                    const float4 a1 = *reinterpret_cast<const float4*>(base + oa[u] + 32);
                    const float4 b0 = *reinterpret_cast<const float4*>(base + ob[u]);
                    const float4 b1 = *reinterpret_cast<const float4*>(base + ob[u] + 32);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const unsigned long long bv[4] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) ffma2s(f[i][j], av[i], bv[j]);
                }
            }
        }
        if (t + 2 < nt) bar_arrive<EMPTY>(b);
        if ((t & (AFR_FLUSH - 1)) == AFR_FLUSH - 1) flush();
    }
    if (nt & (AFR_FLUSH - 1)) flush();
    double* out = part + ((size_t)ci * gridDim.x + pi) * 4096;
    for (int e = cid; e < AFR_TILES * 64; e += AFR_THREADS - AFR_PROD) {
        int tt = e >> 6, t_i = 0;                   // tile index -> (t_i, t_j) as above
        while (tt >= 8 - t_i) { tt -= 8 - t_i; ++t_i; }
        const int t_j = t_i + tt, el = e & 63;
        out[(t_i * 8 + (el >> 3)) * 64 + t_j * 8 + (el & 7)] = accs[e];
    }
}

// ---- the same sums on the legacy tensor path: mma.sync.m16n8k8 with 3 x TF32 operands ---------------------------------
// Measured on a B200 (scripts/micro/mma_sync_rate.cu, profiles/r02_s2_mma_sync_rate.txt): register-fed mma.sync tf32 issues
// 0.5 MMA per clock and SM = 512 FMA/clk, packed FFMA2 1.75 per clock = 112 FMA/clk (below the 128 of plain FFMA: the
// packed form saves issue slots, not pipe cycles).  With every fp32 value split as x = hi + lo (hi = tf32(x), lo = tf32(x - hi):
// 22 mantissa bits) and the products taken as lo hi + hi lo + hi hi (the dropped lo lo term is 2^-22 of the product) the
// tensor path is worth 170 fp32-grade FMA per clock on a pipe of its own -- the encode work of the producer warps no
// longer competes with the outer products for the FMA pipe.
//   * the moment matrix is y^T y with M = N = the 64 columns and K = the rows: A(m, k) = y[k][m], B(k, n) = y[k][n] are the
//     SAME shared-memory tile, and the A fragment of the 16 columns 16 mi .. is the pair of B fragments 2 mi, 2 mi + 1;
//   * upper triangle in 16 x 8 tiles: (mi, nj) with nj >= 2 mi, 20 tiles in four sets of five, one per consumer warp, which
//     owns them for every row: 20 accumulator registers and its own fp64 accumulators (no atomics), fragments of the next
//     8-row step in flight while the 15 MMAs of the current one issue;
//   * producers: warp p fills ring buffer p of four (sub-tiles p, p + 4, ...: 32 rows, one per lane): encode, split, 16 + 16
//     STS.128 into the hi / lo tiles (row stride 72 words: the fragment loads of a warp hit 32 different banks).  A first
//     form with two 64-row buffers, each filled by its own pair of warps, left ONE pair encoding at a time and was
//     producer-latency-bound (1.55 ms per C2 step against 1.33 ms for the FFMA2 kernel);
//   * measured (C2 closed-form step, same box, ms per step for the moment kernels): FFMA2 kernel 1.30; this kernel 1.55 (two
//     64-row buffers), 1.70 (MMAs issued tile by tile behind asm volatile), 1.51 (A quads loaded instead of copied), 1.34
//     (four tile sets, no atomics, fragments one step ahead).  Same parity, same speed -- and not because of the producers:
//     with PCNERF_AFF_DEBUG=1 (producers write a cheap function of the position, results wrong) the FFMA2 kernel takes 1.14
//     and this one 1.17 ms.  Both saturate the shared-memory pipe (26 - 35 wavefronts per row at one per clock and SM: here
//     every element is loaded as a 4-byte fragment by 4.5 of the consumer warps and the hi / lo stores run at a two-way bank
//     conflict); the tensor pipe sits at ~30 %.  Opt-in (PCNERF_AFF_MOMENTS=tc); FFMA2 stays the default.
//   * fp32 accumulation over AMT_FLUSH sub-tiles (128 rows), then added into fp64 accumulators in shared memory.  The tensor core truncates its fp32 accumulator, so sums of like-signed products (the diagonal) come
//     out low by ~1e-6 relative -- a common scaling of the covariance that BatchNorm's own normalisation absorbs layer by
//     layer (parity gates in tests/test_gpu_affine.py are unchanged).
#define AMT_T 32                // rows per sub-tile = one producer warp
#define AMT_NBUF 4              // ring of sub-tile buffers, buffer p is filled by producer warp p
#define AMT_LD 72               // row stride of the hi / lo tiles in 32-bit words
#define AMT_TILE (AMT_T * AMT_LD)
#define AMT_FLUSH 4             // sub-tiles (128 rows) between folds of a warp's fp32 accumulators into its fp64 ones (power of two)
#define AMT_SMEM (AMT_NBUF * 2 * AMT_TILE * 4 + 4096 * 8)          // ring x (hi, lo) + accs [64][64] fp64 = 106,496 bytes

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// A fragment of 16 columns starting at p (= the lane's element of column block 2 mi, step row t4):
// (y[t4][c], y[t4][c + 8], y[t4 + 4][c], y[t4 + 4][c + 8])
__device__ __forceinline__ void amt_lds4(uint32_t (&a)[4], const uint32_t* p, uint32_t zero) {
    // `zero` is a run-time 0 the compiler cannot fold: the address differs formally from the B fragments' and the loads stay
    // (derived from the B registers, ptxas re-assembled the interleaved quads in front of almost every MMA: 2 IMAD.MOV per HMMA)
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(p) + zero;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a[0]) : "r"(sa));
    asm volatile("ld.shared.b32 %0, [%1+32];" : "=r"(a[1]) : "r"(sa));
    asm volatile("ld.shared.b32 %0, [%1+1152];" : "=r"(a[2]) : "r"(sa));          // + 4 rows = 4 * AMT_LD * 4 bytes
    asm volatile("ld.shared.b32 %0, [%1+1184];" : "=r"(a[3]) : "r"(sa));
}

// The 20 tiles (mi, nj), nj >= 2 mi, in four sets of five, one per consumer warp (SET = warp index):
//   0: mi 0, nj 0..4      1: mi 0, nj 5..7 + mi 3, nj 6..7      2: mi 1, nj 2..6      3: mi 1, nj 7 + mi 2, nj 4..7
// A warp owns its tiles for ALL rows (four 8-row steps per sub-tile), so its fp64 accumulators are its own (no atomics).
template <int SET> struct AmtSet;
template <> struct AmtSet<0> { static constexpr int MA = 0, NA0 = 0, NA = 5, MB = 0, NB0 = 0, NB = 0; };
template <> struct AmtSet<1> { static constexpr int MA = 0, NA0 = 5, NA = 3, MB = 3, NB0 = 6, NB = 2; };
template <> struct AmtSet<2> { static constexpr int MA = 1, NA0 = 2, NA = 5, MB = 1, NB0 = 2, NB = 0; };
template <> struct AmtSet<3> { static constexpr int MA = 1, NA0 = 7, NA = 1, MB = 2, NB0 = 4, NB = 4; };

// fragments of one 8-row step: B pairs of the set's column blocks and the A quads of its (one or two) row blocks, hi and lo
template <int SET>
struct AmtFrag {
    uint32_t bh[5][2], bl[5][2], ahA[4], alA[4], ahB[4], alB[4];
    __device__ __forceinline__ void load(const uint32_t* __restrict__ h0, const uint32_t* __restrict__ l0, uint32_t zero) {
        using S = AmtSet<SET>;
#pragma unroll
        for (int q = 0; q < S::NA; ++q) {
            bh[q][0] = h0[8 * (S::NA0 + q)]; bh[q][1] = h0[4 * AMT_LD + 8 * (S::NA0 + q)];
            bl[q][0] = l0[8 * (S::NA0 + q)]; bl[q][1] = l0[4 * AMT_LD + 8 * (S::NA0 + q)];
        }
#pragma unroll
        for (int q = 0; q < S::NB; ++q) {
            bh[S::NA + q][0] = h0[8 * (S::NB0 + q)]; bh[S::NA + q][1] = h0[4 * AMT_LD + 8 * (S::NB0 + q)];
            bl[S::NA + q][0] = l0[8 * (S::NB0 + q)]; bl[S::NA + q][1] = l0[4 * AMT_LD + 8 * (S::NB0 + q)];
        }
        amt_lds4(ahA, h0 + 16 * S::MA, zero); amt_lds4(alA, l0 + 16 * S::MA, zero);
        if (S::NB > 0) { amt_lds4(ahB, h0 + 16 * S::MB, zero); amt_lds4(alB, l0 + 16 * S::MB, zero); }
    }
    // the three products of a tile (lo hi, hi lo, hi hi: small terms first) as three passes over the five tiles
    __device__ __forceinline__ void mma(float (&acc)[5][4]) const {
        using S = AmtSet<SET>;
#define AMT_PASS(AA_, AB_, B_)                                                                              \
        do {                                                                                                \
            _Pragma("unroll") for (int q = 0; q < S::NA; ++q)                                               \
                mma_tf32(acc[q], AA_[0], AA_[1], AA_[2], AA_[3], B_[q][0], B_[q][1]);                       \
            _Pragma("unroll") for (int q = 0; q < S::NB; ++q)                                               \
                mma_tf32(acc[S::NA + q], AB_[0], AB_[1], AB_[2], AB_[3], B_[S::NA + q][0], B_[S::NA + q][1]); \
        } while (0)
        AMT_PASS(alA, alB, bh);
        AMT_PASS(ahA, ahB, bl);
        AMT_PASS(ahA, ahB, bh);
#undef AMT_PASS
    }
};

// one 32-row sub-tile (four steps) of tile set SET; the fragments of step ks + 1 are requested before the MMAs of step ks
template <int SET>
__device__ __forceinline__ void amt_consume(const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo, float (&acc)[5][4],
                                            int g, int t4, uint32_t zero) {
    const uint32_t* h0 = hi + t4 * AMT_LD + g;
    const uint32_t* l0 = lo + t4 * AMT_LD + g;
    AmtFrag<SET> f0, f1;
    f0.load(h0, l0, zero);
    f1.load(h0 + 8 * AMT_LD, l0 + 8 * AMT_LD, zero);
    f0.mma(acc);
    f0.load(h0 + 16 * AMT_LD, l0 + 16 * AMT_LD, zero);
    f1.mma(acc);
    f1.load(h0 + 24 * AMT_LD, l0 + 24 * AMT_LD, zero);
    f0.mma(acc);
    f1.mma(acc);
}

// (mi, nj) of accumulator tile e of tile set `set`
__device__ __forceinline__ void amt_tile_of(int set, int e, int& mi, int& nj) {
    switch (set) {
        case 0: mi = 0; nj = e; break;
        case 1: mi = e < 3 ? 0 : 3; nj = e < 3 ? 5 + e : 3 + e; break;
        case 2: mi = 1; nj = 2 + e; break;
        default: mi = e < 1 ? 1 : 2; nj = e < 1 ? 7 : 3 + e; break;
    }
}

// named barriers 1 + b (FULL) / 5 + b (EMPTY) of ring buffer b: 32 producer threads + 128 consumer threads
__device__ __forceinline__ void amt_full_sync(int b) {
    switch (b) { case 0: bar_sync_id<1, 160>(); break; case 1: bar_sync_id<2, 160>(); break;
                 case 2: bar_sync_id<3, 160>(); break; default: bar_sync_id<4, 160>(); }
}
__device__ __forceinline__ void amt_full_arrive(int b) {
    switch (b) { case 0: bar_arrive_id<1, 160>(); break; case 1: bar_arrive_id<2, 160>(); break;
                 case 2: bar_arrive_id<3, 160>(); break; default: bar_arrive_id<4, 160>(); }
}
__device__ __forceinline__ void amt_empty_sync(int b) {
    switch (b) { case 0: bar_sync_id<5, 160>(); break; case 1: bar_sync_id<6, 160>(); break;
                 case 2: bar_sync_id<7, 160>(); break; default: bar_sync_id<8, 160>(); }
}
__device__ __forceinline__ void amt_empty_arrive(int b) {
    switch (b) { case 0: bar_arrive_id<5, 160>(); break; case 1: bar_arrive_id<6, 160>(); break;
                 case 2: bar_arrive_id<7, 160>(); break; default: bar_arrive_id<8, 160>(); }
}

struct RowRaw { float z, o0, o1, o2, d0, d1, d2; };
__device__ __forceinline__ RowRaw row_raw(const RayRows& s, int64_t r) {
    const uint32_t ray = (uint32_t)r / (uint32_t)s.S;
    const float* q = s.rays + (int64_t)ray * s.ld;
    RowRaw v;
    v.z = __ldg(s.z + r);
    v.o0 = __ldg(q); v.o1 = __ldg(q + 1); v.o2 = __ldg(q + 2);
    v.d0 = __ldg(q + 3); v.d1 = __ldg(q + 4); v.d2 = __ldg(q + 5);
    return v;
}

__global__ void __launch_bounds__(256, 2) k_affine_moments_rays_tc(RayRows src, int64_t rows, int64_t chunk,
                                                                   double* __restrict__ part, double* __restrict__ shift_out,
                                                                   int dbg) {
    extern __shared__ __align__(16) uint32_t xt[];          // [AMT_NBUF][hi, lo][AMT_T][AMT_LD] | accs [64][64] fp64
    __shared__ float shift[4];
    double* accs = reinterpret_cast<double*>(xt + AMT_NBUF * 2 * AMT_TILE);
    const int ci = blockIdx.y, pi = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    int64_t c_beg, r_beg, r_end;
    part_range(rows, chunk, ci, pi, c_beg, r_beg, r_end);
    if (tid == 0) ray_row_pos(src, c_beg, shift[0], shift[1], shift[2]);
    for (int e = tid; e < 4096; e += 256) accs[e] = 0.0;
    __syncthreads();
    if (pi == 0 && tid < 64) shift_out[ci * 64 + tid] = tid < 3 ? (double)shift[tid] : 0.0;
    const int nt = (int)((r_end - r_beg + AMT_T - 1) / AMT_T);
    if (tid < 128) {
        // producer warp p: sub-tiles p, p + 4, ... into ring buffer p, one row per lane; the raw ray / depth values of the
        // NEXT sub-tile are requested before the current one is encoded
        const int p = tid >> 5;
        const float s0 = shift[0], s1 = shift[1], s2 = shift[2];
        uint32_t* hi = xt + p * 2 * AMT_TILE + lane * AMT_LD;
        uint32_t* lo = hi + AMT_TILE;
        int64_t r = r_beg + (int64_t)p * AMT_T + lane;
        RowRaw cur = {};
        if (r < r_end) cur = row_raw(src, r);
        for (int t = p; t < nt; t += AMT_NBUF, r += AMT_NBUF * AMT_T) {
            const int64_t rn = r + AMT_NBUF * AMT_T;
            RowRaw nxt = {};
            if (rn < r_end) nxt = row_raw(src, rn);
            float e[64];
            if (r < r_end && dbg) {                      // timing experiment (PCNERF_AFF_DEBUG=1, wrong results): free producers
                const float x0 = __fadd_rn(cur.o0, __fmul_rn(cur.d0, cur.z));
#pragma unroll
                for (int col = 0; col < 64; ++col) e[col] = x0 * (float)(col + 1);
            } else if (r < r_end) {
                const float x0 = __fadd_rn(cur.o0, __fmul_rn(cur.d0, cur.z)), x1 = __fadd_rn(cur.o1, __fmul_rn(cur.d1, cur.z)),
                            x2 = __fadd_rn(cur.o2, __fmul_rn(cur.d2, cur.z));          // nof/render.py:458, as ray_row_pos
                enc_visit_poly(x0, x1, x2, [&](int col, float v) { e[col] = v; });
                e[0] -= s0; e[1] -= s1; e[2] -= s2;
                e[63] = 1.f;
            } else {
#pragma unroll
                for (int col = 0; col < 64; ++col) e[col] = 0.f;
            }
            if (t >= AMT_NBUF) amt_empty_sync(p);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    h[q] = to_tf32(e[4 * c + q]);
                    l[q] = to_tf32(e[4 * c + q] - __uint_as_float(h[q]));
                }
                *reinterpret_cast<uint4*>(hi + 4 * c) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(lo + 4 * c) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            __threadfence_block();
            amt_full_arrive(p);
            cur = nxt;
        }
        return;
    }
    const int cw = (tid - 128) >> 5, g = lane >> 2, t4 = lane & 3;
    const uint32_t zero = (gridDim.z - 1u) * 4u;             // 0 at run time (amt_lds4)
    float acc[5][4];
#pragma unroll
    for (int e = 0; e < 5; ++e) acc[e][0] = acc[e][1] = acc[e][2] = acc[e][3] = 0.f;
    auto flush = [&]() {
        // the warp's own 5 x 128 elements of accs: plain read-modify-write, 20 independent chains
#pragma unroll
        for (int e = 0; e < 5; ++e) {
            int mi, nj;
            amt_tile_of(cw, e, mi, nj);
            double2* d = reinterpret_cast<double2*>(accs + (16 * mi + g) * 64 + 8 * nj + 2 * t4);
            double2 v0 = d[0], v1 = d[8 * 32];
            v0.x += (double)acc[e][0]; v0.y += (double)acc[e][1];
            v1.x += (double)acc[e][2]; v1.y += (double)acc[e][3];
            d[0] = v0; d[8 * 32] = v1;
            acc[e][0] = acc[e][1] = acc[e][2] = acc[e][3] = 0.f;
        }
    };
    for (int t = 0; t < nt; ++t) {
        const int b = t & (AMT_NBUF - 1);
        amt_full_sync(b);
        const uint32_t* hi = xt + b * 2 * AMT_TILE;
        switch (cw) {
            case 0: amt_consume<0>(hi, hi + AMT_TILE, acc, g, t4, zero); break;
            case 1: amt_consume<1>(hi, hi + AMT_TILE, acc, g, t4, zero); break;
            case 2: amt_consume<2>(hi, hi + AMT_TILE, acc, g, t4, zero); break;
            default: amt_consume<3>(hi, hi + AMT_TILE, acc, g, t4, zero); break;
        }
        if (t + AMT_NBUF < nt) amt_empty_arrive(b);
        if ((t & (AMT_FLUSH - 1)) == AMT_FLUSH - 1) flush();
    }
    if (nt & (AMT_FLUSH - 1)) flush();
    bar_sync_id<9, 128>();
    // every element of the 20 tiles (a superset of the 8 x 8 tiles k_affine_moments_reduce reads); the rest stays zero
    double* out = part + ((size_t)ci * gridDim.x + pi) * 4096;
    for (int e = tid - 128; e < 4096; e += 128) out[e] = accs[e];
}

// S[ci][e] = sum over the chunk's parts (upper-triangle tiles only; fixed order: deterministic)
__global__ void __launch_bounds__(256) k_affine_moments_reduce(const double* __restrict__ part, int parts,
                                                               double* __restrict__ S) {
    const int ci = blockIdx.y, e = blockIdx.x * 256 + threadIdx.x, i = e >> 6, j = e & 63;
    if ((i >> 3) > (j >> 3)) return;
    const double* base = part + (size_t)ci * parts * 4096 + e;
    double s = 0.0;
#pragma unroll 4
    for (int p = 0; p < parts; ++p) s += base[(size_t)p * 4096];
    S[(size_t)ci * 4096 + e] = s;
}

// per chunk: S -> m (64; m_63 = 1), C (64 x 64 covariance; row / column 63 zero), cnt
__global__ void __launch_bounds__(256) k_affine_moments_finish(const double* __restrict__ S, const double* __restrict__ shift,
                                                               double* __restrict__ m, double* __restrict__ C,
                                                               double* __restrict__ cnt) {
    __shared__ double s1[64];
    const int ci = blockIdx.x, tid = threadIdx.x;
    const double* Sc = S + (size_t)ci * 4096;
    const double n = Sc[63 * 64 + 63];
    if (tid < 64) s1[tid] = Sc[tid * 64 + 63] / n;
    __syncthreads();
    for (int e = tid; e < 4096; e += 256) {
        const int i = e >> 6, j = e & 63;
        double v = 0.0;
        if (i < 63 && j < 63) v = Sc[(i >> 3) <= (j >> 3) ? e : j * 64 + i] / n - s1[i] * s1[j];
        C[(size_t)ci * 4096 + e] = v;
    }
    if (tid < 64) m[ci * 64 + tid] = tid < 63 ? shift[ci * 64 + tid] + s1[tid] : 1.0;
    if (tid == 0) cnt[ci] = n;
}

// p_r = sigmoid(alpha[chunk(r)] . (x_r, 1))
__global__ void __launch_bounds__(256) k_affine_apply_rays(RayRows src, int64_t rows, int64_t chunk,
                                                           const float* __restrict__ alpha, float* __restrict__ out_p) {
    __shared__ float sa[64];
    const int ci = blockIdx.y, pi = blockIdx.x, tid = threadIdx.x;
    int64_t c_beg, r_beg, r_end;
    part_range(rows, chunk, ci, pi, c_beg, r_beg, r_end);
    if (tid < 64) sa[tid] = alpha[ci * 64 + tid];
    __syncthreads();
    for (int64_t r = r_beg + tid; r < r_end; r += 256) {
        float x0, x1, x2;
        ray_row_pos(src, r, x0, x1, x2);
        float t0 = fmaf(x0, sa[0], sa[63]), t1 = x1 * sa[1], t2 = x2 * sa[2];      // one chain per coordinate
        enc_visit_coord_poly<0>(x0, [&](int col, float v) { t0 = fmaf(v, sa[col], t0); });
        enc_visit_coord_poly<1>(x1, [&](int col, float v) { t1 = fmaf(v, sa[col], t1); });
        enc_visit_coord_poly<2>(x2, [&](int col, float v) { t2 = fmaf(v, sa[col], t2); });
        const float t = t0 + (t1 + t2);
        out_p[r] = 1.f / (1.f + expf(-t));
    }
}

// part[(ci, pi)][64]: sum over the part's rows of g_r (x_r, 1),  g_r = dL/dp_r * p_r (1 - p_r)
__global__ void __launch_bounds__(256, 2) k_affine_grad_rays(RayRows src, const float* __restrict__ p,
                                                             const float* __restrict__ grad_p, int64_t rows, int64_t chunk,
                                                             double* __restrict__ part) {
    __shared__ double red[8][64];
    const int ci = blockIdx.y, pi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    int64_t c_beg, r_beg, r_end;
    part_range(rows, chunk, ci, pi, c_beg, r_beg, r_end);
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    for (int64_t r = r_beg + tid; r < r_end; r += 256) {
        const float pv = p[r];
        const float g = grad_p[r] * pv * (1.f - pv);
        float x0, x1, x2;
        ray_row_pos(src, r, x0, x1, x2);
        enc_visit_poly(x0, x1, x2, [&](int col, float v) { acc[col] = fmaf(g, v, acc[col]); });
        acc[63] += g;
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        const double v = warp_sum_d((double)acc[i]);
        if (lane == 0) red[wib][i] = v;
    }
    __syncthreads();
    if (tid < 64) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][tid];
        part[((size_t)ci * gridDim.x + pi) * 64 + tid] = t;
    }
}

__global__ void k_affine_grad_finish(const double* __restrict__ part, int parts, double* __restrict__ dalpha) {
    const int ci = blockIdx.x, j = threadIdx.x;
    double s = 0.0;
    for (int p = 0; p < parts; ++p) s += part[((size_t)ci * parts + p) * 64 + j];
    dalpha[ci * 64 + j] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// parameter-sized algebra (float64).  Per-layer matrices are stored [256][nc][64]: the chunks side by side along the
// columns, so that W X is ONE 256 x 256 x (nc 64) GEMM and sum_chunks G X^T one 256 x (nc 64) x 256 GEMM.
// ---------------------------------------------------------------------------------------------------------------

// D (8 x 8) += A (8 x 4, row) B (4 x 8, col) on the fp64 tensor pipe.  Lane l holds A[l / 4][l % 4], B[l % 4][l / 4] and
// D[l / 4][2 (l % 4) + {0, 1}].  (The CUDA-core DFMA form of this GEMM ran at 8 TFLOP/s; cuBLAS's DMMA kernels reach 36.)
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C[m][n] = sum_{k in split} A(m, k) B(k, n), one 32 x 32 tile per block.  The kernel is bound by the latency of its operand
// loads (L2 -> registers -> shared memory), not by the tensor pipe (ncu: long-scoreboard stalls, 10 % issue activity with one
// k pipeline per block), so the block's k range is split over its FOUR WARPS: each warp runs its own
// load -> shared -> DMMA pipeline over a quarter of the k steps for the whole tile (4 x 4 DMMA tiles, 16-wide k steps, no
// block-wide barrier), and the four partial tiles meet in shared memory at the end.
// A(m, k) = A[m a_sm + k a_sk], B(k, n) = B[k b_sk + n b_sn]; *_KC: the operand is contiguous along k (else along m / n) --
// only the lane -> element mapping of the tile loads depends on it.  M, N multiples of 32, K (per split) of 64.
template <typename TA, bool A_KC, bool B_KC>
__device__ __forceinline__ void aff_dgemm_tile(const TA* __restrict__ A, int64_t a_sm, int64_t a_sk,
                                               const double* __restrict__ B, int64_t b_sk, int64_t b_sn,
                                               double* __restrict__ out, int64_t ldc, int k_beg, int k_end) {
    // leading dimension 40: the four k rows of a fragment load start 0 / 8 / 0 / 8 (mod 16) eight-byte banks apart, so the 32
    // lanes of a fragment load touch every bank exactly twice (the minimum for 256 bytes)
    __shared__ double sm[4][2][16][40];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, fr = lane >> 2, fk = lane & 3;
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int kq = (k_end - k_beg) >> 2;                    // k steps of this warp: [k_beg + w kq, + kq)
    double (*As)[40] = sm[w][0];
    double (*Bs)[40] = sm[w][1];
    double c[4][4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[t][u][0] = c[t][u][1] = 0.0;
    for (int k0 = k_beg + w * kq; k0 < k_beg + (w + 1) * kq; k0 += 16) {
        double ra[16], rb[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int ka = A_KC ? (lane & 15) : q, mm = A_KC ? (lane >> 4) + 2 * q : lane;
            ra[q] = (double)A[(int64_t)(m0 + mm) * a_sm + (int64_t)(k0 + ka) * a_sk];
            const int kb = B_KC ? (lane & 15) : q, nn = B_KC ? (lane >> 4) + 2 * q : lane;
            rb[q] = B[(int64_t)(k0 + kb) * b_sk + (int64_t)(n0 + nn) * b_sn];
        }
        __syncwarp();                                        // the previous step's fragment loads are done
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int ka = A_KC ? (lane & 15) : q, mm = A_KC ? (lane >> 4) + 2 * q : lane;
            As[ka][mm] = ra[q];
            const int kb = B_KC ? (lane & 15) : q, nn = B_KC ? (lane >> 4) + 2 * q : lane;
            Bs[kb][nn] = rb[q];
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { a[t] = As[4 * ks + fk][8 * t + fr]; b[t] = Bs[4 * ks + fk][8 * t + fr]; }
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) dmma_884(c[t][u][0], c[t][u][1], a[t], b[u]);
        }
    }
    // the four partial tiles -> shared memory ([w][32][32 + 2]: 8.7 KB each, the operand buffers are free now) -> C
    __syncthreads();
    double* red = &sm[0][0][0][0];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u)
            *reinterpret_cast<double2*>(red + (w * 32 + 8 * t + fr) * 34 + 8 * u + 2 * fk) = make_double2(c[t][u][0], c[t][u][1]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int e = tid + 128 * q, r = e >> 5, cc = e & 31;
        out[(int64_t)(m0 + r) * ldc + n0 + cc] = (red[r * 34 + cc] + red[(32 + r) * 34 + cc]) + (red[(64 + r) * 34 + cc] + red[(96 + r) * 34 + cc]);
    }
}

template <typename TA, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(128) k_aff_dgemm(const TA* __restrict__ A, int64_t a_sm, int64_t a_sk,
                                                   const double* __restrict__ B, int64_t b_sk, int64_t b_sn,
                                                   double* __restrict__ C, int64_t ldc, int64_t split_stride, int K,
                                                   int k_per_split) {
    const int k_beg = blockIdx.z * k_per_split;
    aff_dgemm_tile<TA, A_KC, B_KC>(A, a_sm, a_sk, B, b_sk, b_sn, C + (size_t)blockIdx.z * split_stride, ldc, k_beg,
                                   min(K, k_beg + k_per_split));
}

// The seven weight-gradient products of a backward pass in ONE launch: blockIdx.z = (layer - 1) * splits + split,
//   part[layer - 1][split] (256 x 256) = sum over the split's chunks of G_layer Ab_{layer-1}^T     (K = nc 64 per layer)
struct AffWgradBatch {
    const double* G[7];          // dL/dA_l, l = 1..7   [256][nc][64]
    const double* Ab[7];         // Ab_{l-1}
};
__global__ void __launch_bounds__(128) k_aff_wgrad_batched(AffWgradBatch b, int N, int splits, int k_per_split,
                                                           double* __restrict__ part) {
    const int li = blockIdx.z / splits, sp = blockIdx.z - li * splits, k_beg = sp * k_per_split;
    const double *Gp = b.G[0], *Ap = b.Ab[0];              // (selects, not a dynamically indexed parameter array: no local copy)
#pragma unroll
    for (int i = 1; i < 7; ++i)
        if (li == i) { Gp = b.G[i]; Ap = b.Ab[i]; }
    aff_dgemm_tile<double, true, true>(Gp, N, 1, Ap, 1, N, part + (size_t)blockIdx.z * 65536, 256, k_beg,
                                       min(N, k_beg + k_per_split));
}

// dW_l[i][k] += sum_splits part[l - 1][s][i][k] for the seven layers (grid (256, 7))
struct AffWgradOut { float* dW[7]; int ld[7]; };
__global__ void __launch_bounds__(256) k_aff_wgrad_reduce_batched(const double* __restrict__ part, int splits, AffWgradOut o) {
    const int li = blockIdx.y, e = blockIdx.x * 256 + threadIdx.x;
    const double* src = part + (size_t)li * splits * 65536 + e;
    double s = 0.0;
    for (int q = 0; q < splits; ++q) s += src[(size_t)q * 65536];
    float* dW = o.dW[0];
    int ld = o.ld[0];
#pragma unroll
    for (int i = 1; i < 7; ++i)
        if (li == i) { dW = o.dW[i]; ld = o.ld[i]; }
    dW[(e >> 8) * ld + (e & 255)] += (float)s;
}

// dW[i][k] += sum_splits part[s][i][k]   (256 x 256 block of a weight gradient with leading dimension ld)
__global__ void k_aff_wgrad_reduce(const double* __restrict__ part, int splits, float* __restrict__ dW, int ld) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 65536) return;
    double s = 0.0;
    for (int q = 0; q < splits; ++q) s += part[(size_t)q * 65536 + e];
    dW[(e >> 8) * ld + (e & 255)] += (float)s;
}

struct AffLayer {
    double *A, *AC, *Ab;                    // [256][nc][64]
    double *a, *r, *mean, *var;             // [nc][256]
    const double *m, *C;                    // [nc][64], [nc][64][64]
    const float* Wx; int wx_ld;             // l == 0: W_0 (ld 63);  l == 4: W_4[:, :63] (ld 319);  else null
    int init;                               // 1 (l == 0): A is not read
    const float *bias, *gamma, *beta;
    const float *rmean, *rvar;              // eval mode (running statistics, models.py:183-203 under model.eval()): mean / var of
                                            // BN_l are these instead of A_l m / diag(A_l C A_l^T); m, C, AC and the small vectors unused
    int nc;
    double eps;
};

__device__ __forceinline__ double half_warp_sum(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// grid (16, nc): 16 feature rows per block, 16 lanes per row; lane q owns columns q, q + 16, q + 32, q + 48.
// A <- A (+ [Wx | 0]) + bias e_63^T;  AC = A C;  mean, var, a, r;  Ab = a A + (beta - a mean) e_63^T
__global__ void __launch_bounds__(256) k_aff_layer_fwd(AffLayer g) {
    __shared__ double Cs[64][64];
    __shared__ double As[16][64];
    __shared__ double ms[64];
    const int tid = threadIdx.x, lr = tid >> 4, q = tid & 15, c = blockIdx.y, row = blockIdx.x * 16 + lr;
    const bool eval = g.rvar != nullptr;
    if (!eval) {
#pragma unroll
        for (int e = 0; e < 16; ++e) Cs[0][tid + 256 * e] = g.C[(size_t)c * 4096 + tid + 256 * e];
        if (tid < 64) ms[tid] = g.m[c * 64 + tid];
    }
    const size_t idx = ((size_t)row * g.nc + c) * 64 + q;
    double av[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = q + 16 * e;
        double v = g.init ? 0.0 : g.A[idx + 16 * e];
        if (g.Wx && j < 63) v += (double)g.Wx[(size_t)row * g.wx_ld + j];
        if (j == 63) v += (double)g.bias[row];
        av[e] = v;
        As[lr][j] = v;
    }
    if (eval) {
        const double rr = 1.0 / sqrt((double)g.rvar[row] + g.eps), aa = (double)g.gamma[row] * rr;
        const double s = (double)g.beta[row] - aa * (double)g.rmean[row];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            g.A[idx + 16 * e] = av[e];
            g.Ab[idx + 16 * e] = aa * av[e] + ((q + 16 * e) == 63 ? s : 0.0);
        }
        return;
    }
    __syncthreads();
    double ac[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
        const double ak = As[lr][k];
#pragma unroll
        for (int e = 0; e < 4; ++e) ac[e] = fma(ak, Cs[k][q + 16 * e], ac[e]);
    }
    double pv = 0.0, pm = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) { pv = fma(ac[e], av[e], pv); pm = fma(av[e], ms[q + 16 * e], pm); }
    pv = half_warp_sum(pv);
    pm = half_warp_sum(pm);
    const double var = pv > 0.0 ? pv : 0.0;
    const double rr = 1.0 / sqrt(var + g.eps), aa = (double)g.gamma[row] * rr, s = (double)g.beta[row] - aa * pm;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        g.A[idx + 16 * e] = av[e];
        g.AC[idx + 16 * e] = ac[e];
        g.Ab[idx + 16 * e] = aa * av[e] + ((q + 16 * e) == 63 ? s : 0.0);
    }
    if (q == 0) {
        const int o = c * 256 + row;
        g.a[o] = aa; g.r[o] = rr; g.mean[o] = pm; g.var[o] = var;
    }
}

// ---- one kernel per layer (the default; PCNERF_AFF_FUSED=0 restores the GEMM + per-layer kernel pairs above / below) ----
// The chain is a sequence of small dependent kernels (4.2 GFLOP of float64 per fine pass spread over 62 launches of 5-25 us
// each: launch, cold L2 loads and tail dominate), so each layer is ONE launch per direction: a block owns AFL_ROWS feature
// rows of ONE chunk -- 32 x 64 elements, everything the BatchNorm algebra of a row needs -- computes its tile of the product
// with the previous layer on the fp64 tensor pipe (warp w: the eight columns 8 w .. 8 w + 7, four 8 x 8 DMMA tiles; A
// fragments from the block's W tile in shared memory, B fragments straight from L2 in fragment order) and applies the
// per-row algebra of k_aff_layer_fwd / k_aff_layer_bwd to it.
#define AFL_ROWS 32
#define AFL_AS_LD 66            // tile row stride in doubles: the four rows a warp reads at once land in different banks
#define AFL_WS_LD 260           // forward W tile [32][256] floats: row stride 260 -> A-fragment loads hit 32 different banks
#define AFL_WT_LD 40            // backward W^T tile [256][32] floats: row stride 40 -> likewise
#define AFL_SMEM_FWD (4096 * 8 + AFL_ROWS * AFL_AS_LD * 8 + 64 * 8 + AFL_ROWS * AFL_WS_LD * 4)
#define AFL_SMEM_BWD (AFL_ROWS * AFL_AS_LD * 8 + 256 * AFL_WT_LD * 4)

__device__ __forceinline__ double oct_sum(double v) {            // sum over the 8 lanes that share a row
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// grid (256 / AFL_ROWS, nc).  W: W_l + kOff_l (ld w_ld), Abp: Ab_{l-1} [256][nc][64] (both unused when g.init)
__global__ void __launch_bounds__(256) k_aff_fwd_layer(AffLayer g, const float* __restrict__ W, int w_ld,
                                                       const double* __restrict__ Abp) {
    extern __shared__ __align__(16) unsigned char afl_raw[];
    double* Cs = reinterpret_cast<double*>(afl_raw);             // [64][64]
    double* As = Cs + 4096;                                      // [AFL_ROWS][AFL_AS_LD]
    double* ms = As + AFL_ROWS * AFL_AS_LD;                      // [64]
    float* Ws = reinterpret_cast<float*>(ms + 64);               // [AFL_ROWS][AFL_WS_LD]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, fr = lane >> 2, fk = lane & 3;
    const int c = blockIdx.y, row0 = blockIdx.x * AFL_ROWS, nc = g.nc;
    const bool eval = g.rvar != nullptr;
    if (!g.init) {
#pragma unroll 8
        for (int e = 0; e < AFL_ROWS; ++e) {
            const int idx = tid + 256 * e, r = idx >> 8, k = idx & 255;
            Ws[r * AFL_WS_LD + k] = W[(size_t)(row0 + r) * w_ld + k];
        }
    }
    if (!eval) {
#pragma unroll
        for (int e = 0; e < 16; ++e) Cs[tid + 256 * e] = g.C[(size_t)c * 4096 + tid + 256 * e];
        if (tid < 64) ms[tid] = g.m[c * 64 + tid];
    }
    __syncthreads();
    double acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
    if (!g.init) {
        // eight B fragments are requested at a time and consumed by 32 DMMAs (measured: a register double buffer that ptxas
        // flattens into just-in-time loads is 30 % SLOWER -- one exposed L2 latency per load instead of one per eight)
        const size_t ks = (size_t)nc * 64;
        const double* bp = Abp + (size_t)fk * ks + (size_t)c * 64 + 8 * w + fr;
        const float* ap = Ws + fr * AFL_WS_LD + fk;
#pragma unroll 1
        for (int k0 = 0; k0 < 256; k0 += 32) {
            double b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) b[u] = bp[(size_t)(k0 + 4 * u) * ks];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int t = 0; t < 4; ++t) dmma_884(acc[t][0], acc[t][1], (double)ap[8 * t * AFL_WS_LD + k0 + 4 * u], b[u]);
        }
    }
    // A <- product (+ [Wx | 0]) + bias e_63^T
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = 8 * t + fr, j = 8 * w + 2 * fk + h;
            double v = acc[t][h];
            if (g.Wx && j < 63) v += (double)g.Wx[(size_t)(row0 + r) * g.wx_ld + j];
            if (j == 63) v += (double)g.bias[row0 + r];
            As[r * AFL_AS_LD + j] = v;
        }
    __syncthreads();
    // per-row algebra: 8 lanes per row, lane q owns columns q, q + 8, ..., q + 56
    const int lr = tid >> 3, q = tid & 7, row = row0 + lr;
    const size_t idx = ((size_t)row * nc + c) * 64 + q;
    double av[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) av[e] = As[lr * AFL_AS_LD + q + 8 * e];
    if (eval) {
        const double rr = 1.0 / sqrt((double)g.rvar[row] + g.eps), aa = (double)g.gamma[row] * rr;
        const double sh = (double)g.beta[row] - aa * (double)g.rmean[row];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            g.A[idx + 8 * e] = av[e];
            g.Ab[idx + 8 * e] = aa * av[e] + ((q + 8 * e) == 63 ? sh : 0.0);
        }
        return;
    }
    double ac[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ac[e] = 0.0;
#pragma unroll 4
    for (int k = 0; k < 64; ++k) {
        const double ak = As[lr * AFL_AS_LD + k];
#pragma unroll
        for (int e = 0; e < 8; ++e) ac[e] = fma(ak, Cs[k * 64 + q + 8 * e], ac[e]);
    }
    double pv = 0.0, pm = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { pv = fma(ac[e], av[e], pv); pm = fma(av[e], ms[q + 8 * e], pm); }
    pv = oct_sum(pv);
    pm = oct_sum(pm);
    const double var = pv > 0.0 ? pv : 0.0;
    const double rr = 1.0 / sqrt(var + g.eps), aa = (double)g.gamma[row] * rr, sh = (double)g.beta[row] - aa * pm;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        g.A[idx + 8 * e] = av[e];
        g.AC[idx + 8 * e] = ac[e];
        g.Ab[idx + 8 * e] = aa * av[e] + ((q + 8 * e) == 63 ? sh : 0.0);
    }
    if (q == 0) {
        const int o = c * 256 + row;
        g.a[o] = aa; g.r[o] = rr; g.mean[o] = pm; g.var[o] = var;
    }
}

// alpha[c][j] = sum_i w_out[i] Ab_7[i][c][j] (+ b_out at j = 63)
__global__ void __launch_bounds__(256) k_aff_alpha(const double* __restrict__ Ab, const float* __restrict__ wo,
                                                   const float* __restrict__ bo, int nc, float* __restrict__ alpha) {
    __shared__ double red[4][64];
    const int c = blockIdx.x, j = threadIdx.x & 63, part = threadIdx.x >> 6;
    double s = 0.0;
    for (int i = part * 64; i < part * 64 + 64; ++i) s = fma((double)wo[i], Ab[((size_t)i * nc + c) * 64 + j], s);
    red[part][j] = s;
    __syncthreads();
    if (part == 0) {
        double t = red[0][j] + red[1][j] + red[2][j] + red[3][j];
        if (j == 63) t += (double)bo[0];
        alpha[c * 64 + j] = (float)t;
    }
}

struct AffRunning {
    float* rm[PCNERF_NBN];
    float* rv[PCNERF_NBN];
    int64_t* nbt[PCNERF_NBN];
    const double* mean[PCNERF_NBN];
    const double* var[PCNERF_NBN];
    const double* cnt;
    int nc;
    double mom;
};

// BatchNorm1d's running statistics: one update per chunk, in chunk order (momentum, unbiased variance)
__global__ void k_aff_running(AffRunning g) {
    const int l = blockIdx.x, i = threadIdx.x;
    double rm = (double)g.rm[l][i], rv = (double)g.rv[l][i];
    for (int c = 0; c < g.nc; ++c) {
        const double n = g.cnt[c];
        rm = (1.0 - g.mom) * rm + g.mom * g.mean[l][c * 256 + i];
        rv = (1.0 - g.mom) * rv + g.mom * g.var[l][c * 256 + i] * (n / (n - 1.0));
    }
    g.rm[l][i] = (float)rm;
    g.rv[l][i] = (float)rv;
    if (i == 0) *g.nbt[l] += g.nc;
}

// grid 256 (feature i), block 64 (column j): G_7 = w_out^T dalpha;  dw_out += Ab_7 dalpha;  db_out += sum_c dalpha[c][63]
__global__ void __launch_bounds__(64) k_aff_head_bwd(const double* __restrict__ Ab, const double* __restrict__ dalpha,
                                                     const float* __restrict__ wo, int nc, double* __restrict__ G,
                                                     float* __restrict__ dwo, float* __restrict__ dbo) {
    __shared__ double red[2];
    const int i = blockIdx.x, j = threadIdx.x;
    const double w = (double)wo[i];
    double acc = 0.0, sb = 0.0;
    for (int c = 0; c < nc; ++c) {
        const double da = dalpha[c * 64 + j];
        const size_t idx = ((size_t)i * nc + c) * 64 + j;
        acc = fma(Ab[idx], da, acc);
        G[idx] = w * da;
        if (j == 63) sb += da;
    }
    acc = warp_sum_d(acc);
    if ((j & 31) == 0) red[j >> 5] = acc;
    __syncthreads();
    if (j == 0) dwo[i] += (float)(red[0] + red[1]);
    if (i == 0 && j == 63) dbo[0] += (float)sb;
}

struct AffLayerBwd {
    double* G;                              // k_aff_layer_bwd: in dL/dAb_l, out dL/dA_l (in place)   [256][nc][64]
    const double* Gin;                      // k_aff_bwd_layer: dL/dAb_l when it is not computed by the kernel (l = 7), else null
    double* Gout;                           // k_aff_bwd_layer: dL/dA_l (kept per layer for the batched weight-gradient launch)
    const double *A, *AC;
    const double *a, *r, *mean, *var, *m;
    const float* gamma;
    double *gpart, *bpart;                  // [nc][256]: per-chunk contributions to dgamma_l, dbeta_l
    int nc;
};

// reverse of k_aff_layer_fwd's last four lines (same thread -> element mapping)
__global__ void __launch_bounds__(256) k_aff_layer_bwd(AffLayerBwd g) {
    const int tid = threadIdx.x, lr = tid >> 4, q = tid & 15, c = blockIdx.y, row = blockIdx.x * 16 + lr;
    const size_t idx = ((size_t)row * g.nc + c) * 64 + q;
    double gv[4], av[4], dot = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) { gv[e] = g.G[idx + 16 * e]; av[e] = g.A[idx + 16 * e]; dot = fma(gv[e], av[e], dot); }
    dot = half_warp_sum(dot);
    const double sbar = __shfl_sync(FULL_MASK, gv[3], (threadIdx.x & 16) | 15);      // dL/ds = column 63 of dL/dAb
    const int o = c * 256 + row;
    const double aa = g.a[o], rr = g.r[o];
    const double abar = dot - sbar * g.mean[o];
    const double mbar = -sbar * aa;
    const double vbar = g.var[o] > 0.0 ? -0.5 * abar * (double)g.gamma[row] * rr * rr * rr : 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e)
        g.G[idx + 16 * e] = aa * gv[e] + 2.0 * vbar * g.AC[idx + 16 * e] + mbar * g.m[c * 64 + q + 16 * e];
    if (q == 0) { g.gpart[o] = abar * rr; g.bpart[o] = sbar; }
}

// Backward of layer l in one launch, grid (256 / AFL_ROWS, nc): the block's tile of dL/dAb_l = W'_{l+1}^T dL/dA_{l+1}
// (Wn: W_{l+1} + kOff_{l+1}, ld wn_ld; Gn: dL/dA_{l+1}; l = 7: the tile is read from g.Gin instead), then the reverse of the
// per-row algebra (k_aff_layer_bwd) -> g.Gout = dL/dA_l, g.gpart / g.bpart.
__global__ void __launch_bounds__(256) k_aff_bwd_layer(AffLayerBwd g, const float* __restrict__ Wn, int wn_ld,
                                                       const double* __restrict__ Gn) {
    extern __shared__ __align__(16) unsigned char afl_raw[];
    double* Gs = reinterpret_cast<double*>(afl_raw);             // [AFL_ROWS][AFL_AS_LD]
    float* Wt = reinterpret_cast<float*>(Gs + AFL_ROWS * AFL_AS_LD);   // [256][AFL_WT_LD]: Wt[k][m] = W_{l+1}[k][kOff + row0 + m]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, fr = lane >> 2, fk = lane & 3;
    const int c = blockIdx.y, row0 = blockIdx.x * AFL_ROWS, nc = g.nc;
    if (Gn) {
#pragma unroll 8
        for (int e = 0; e < 32; ++e) {
            const int idx = tid + 256 * e, k = idx >> 5, mm = idx & 31;
            Wt[k * AFL_WT_LD + mm] = Wn[(size_t)k * wn_ld + row0 + mm];
        }
        __syncthreads();
        double acc[4][2];
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t][0] = acc[t][1] = 0.0;
        const size_t ks = (size_t)nc * 64;
        const double* bp = Gn + (size_t)fk * ks + (size_t)c * 64 + 8 * w + fr;
        const float* ap = Wt + fk * AFL_WT_LD + fr;
#pragma unroll 1
        for (int k0 = 0; k0 < 256; k0 += 32) {
            double b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) b[u] = bp[(size_t)(k0 + 4 * u) * ks];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int t = 0; t < 4; ++t) dmma_884(acc[t][0], acc[t][1], (double)ap[(k0 + 4 * u) * AFL_WT_LD + 8 * t], b[u]);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
            *reinterpret_cast<double2*>(Gs + (8 * t + fr) * AFL_AS_LD + 8 * w + 2 * fk) = make_double2(acc[t][0], acc[t][1]);
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + 256 * e, r = idx >> 6, j = idx & 63;
            Gs[r * AFL_AS_LD + j] = g.Gin[((size_t)(row0 + r) * nc + c) * 64 + j];
        }
    }
    __syncthreads();
    const int lr = tid >> 3, q = tid & 7, row = row0 + lr;
    const size_t idx = ((size_t)row * nc + c) * 64 + q;
    double gv[8], av[8], dot = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { gv[e] = Gs[lr * AFL_AS_LD + q + 8 * e]; av[e] = g.A[idx + 8 * e]; dot = fma(gv[e], av[e], dot); }
    dot = oct_sum(dot);
    const double sbar = Gs[lr * AFL_AS_LD + 63];                 // dL/ds = column 63 of dL/dAb
    const int o = c * 256 + row;
    const double aa = g.a[o], rr = g.r[o];
    const double abar = dot - sbar * g.mean[o];
    const double mbar = -sbar * aa;
    const double vbar = g.var[o] > 0.0 ? -0.5 * abar * (double)g.gamma[row] * rr * rr * rr : 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        g.Gout[idx + 8 * e] = aa * gv[e] + 2.0 * vbar * g.AC[idx + 8 * e] + mbar * g.m[c * 64 + q + 8 * e];
    if (q == 0) { g.gpart[o] = abar * rr; g.bpart[o] = sbar; }
}

// the sums over chunks that end in parameter gradients, all eight layers in one launch: grid (256 (i), 8 (layer)), block 64 (j)
struct AffColsumBatch {
    const double* G[8];                     // dL/dA_l
    const double *gpart[8], *bpart[8];      // [nc][256]
    float *dgamma[8], *dbeta[8], *dbias[8];
    float* dWx[8];                          // layers 0 and 4: the block that multiplies the encoding (else null)
    int wx_ld[8];
};
__global__ void __launch_bounds__(64) k_aff_colsum_batched(AffColsumBatch b, int nc) {
    const int i = blockIdx.x, l = blockIdx.y, j = threadIdx.x;
    const double *G = b.G[0], *gp = b.gpart[0], *bp = b.bpart[0];
    float *dg = b.dgamma[0], *dbt = b.dbeta[0], *dbs = b.dbias[0], *dWx = b.dWx[0];
    int wx_ld = b.wx_ld[0];
#pragma unroll
    for (int q = 1; q < 8; ++q)
        if (l == q) {
            G = b.G[q]; gp = b.gpart[q]; bp = b.bpart[q];
            dg = b.dgamma[q]; dbt = b.dbeta[q]; dbs = b.dbias[q]; dWx = b.dWx[q]; wx_ld = b.wx_ld[q];
        }
    if (dWx || j == 63) {
        double s = 0.0;
        for (int c = 0; c < nc; ++c) s += G[((size_t)i * nc + c) * 64 + j];
        if (j == 63) dbs[i] += (float)s;
        else dWx[(size_t)i * wx_ld + j] += (float)s;
    }
    if (j < 2) {
        const double* src = j == 0 ? gp : bp;
        double s = 0.0;
        for (int c = 0; c < nc; ++c) s += src[c * 256 + i];
        (j == 0 ? dg : dbt)[i] += (float)s;
    }
}

// grid 256 (i), block 64 (j): the sums over chunks that end in parameter gradients of layer l
__global__ void __launch_bounds__(64) k_aff_colsum_bwd(const double* __restrict__ G, const double* __restrict__ gpart,
                                                       const double* __restrict__ bpart, int nc, float* __restrict__ dgamma,
                                                       float* __restrict__ dbeta, float* __restrict__ dbias,
                                                       float* __restrict__ dWx, int wx_ld) {
    const int i = blockIdx.x, j = threadIdx.x;
    if (dWx || j == 63) {
        double s = 0.0;
        for (int c = 0; c < nc; ++c) s += G[((size_t)i * nc + c) * 64 + j];
        if (j == 63) dbias[i] += (float)s;
        else dWx[(size_t)i * wx_ld + j] += (float)s;
    }
    if (j < 2) {
        const double* src = j == 0 ? gpart : bpart;
        double s = 0.0;
        for (int c = 0; c < nc; ++c) s += src[c * 256 + i];
        (j == 0 ? dgamma : dbeta)[i] += (float)s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------

namespace {

struct Work {                 // offsets in doubles
    size_t part, shift, m, C, cnt, layer[8], small[8], G0, G1, dalpha, gpart, bpart, wsplit, alpha, total;
    size_t Gl[8], gpl[8], bpl[8];      // fused backward: dL/dA_l and the per-chunk dgamma / dbeta terms of every layer
    size_t mat;               // 256 * nc * 64
};

Work work_layout(int64_t nc) {
    Work w;
    size_t o = 0;
    w.mat = (size_t)256 * nc * 64;
    auto take = [&](size_t n) { size_t at = o; o += (n + 15) & ~(size_t)15; return at; };
    w.part = take((size_t)nc * AFR_PARTS * 4096);
    w.shift = take(nc * 64);
    w.m = take(nc * 64);
    w.C = take(nc * 4096);
    w.cnt = take(nc);
    for (int l = 0; l < 8; ++l) { w.layer[l] = take(3 * w.mat); w.small[l] = take(4 * nc * 256); }
    w.G0 = take(w.mat);
    w.G1 = take(w.mat);
    w.dalpha = take(nc * 64);
    w.gpart = take(nc * 256);
    w.bpart = take(nc * 256);
    w.wsplit = take((size_t)7 * AFR_SPLITS * 65536);
    w.alpha = take(nc * 32);              // nc x 64 floats
    for (int l = 0; l < 8; ++l) { w.Gl[l] = take(w.mat); w.gpl[l] = take(nc * 256); w.bpl[l] = take(nc * 256); }
    w.total = o;
    return w;
}

struct LayerView { double *A, *AC, *Ab, *a, *r, *mean, *var; };
LayerView layer_view(double* base, const Work& w, int l, int64_t nc) {
    LayerView v;
    v.A = base + w.layer[l]; v.AC = v.A + w.mat; v.Ab = v.AC + w.mat;
    v.a = base + w.small[l]; v.r = v.a + nc * 256; v.mean = v.r + nc * 256; v.var = v.mean + nc * 256;
    return v;
}

int check_common(const pcnerf_mlp_params* P, const float* rays, int ld, int64_t n_rays, const float* z, int S, int64_t chunk,
                 void* work, size_t work_bytes, const char* who) {
    PCN_CHECK_ARG(P && rays && z && work, "%s: null argument", who);
    PCN_CHECK_ARG(ld >= 6 && n_rays >= 1 && S >= 1 && chunk >= 2, "%s: bad shape", who);
    const int64_t rows = n_rays * S;
    PCN_CHECK_ARG(rows < ((int64_t)1 << 31), "%s: %lld rows (limit 2^31 - 1)", who, (long long)rows);
    const int64_t nc = pcn_cdiv(rows, chunk);
    PCN_CHECK_ARG(nc <= 65535, "%s: too many chunks (%lld)", who, (long long)nc);
    PCN_CHECK_ARG(rows - (nc - 1) * chunk >= 2, "%s: a BatchNorm batch of one row has no variance", who);
    PCN_CHECK_ARG(work_bytes >= work_layout(nc).total * sizeof(double), "%s: work area too small", who);
    if (!P->training) {
        pcn_set_error("%s: training-mode (batch statistics) only; eval-mode BN goes through pcnerf_affine_apply", who);
        return PCNERF_ERR_UNSUPPORTED;
    }
    return 0;
}

// CTAs per chunk of the data-sized kernels: whole waves of two CTAs per SM, about 48 per chunk, at most AFR_PARTS
int parts_for(int64_t nc) {
    const int slots = 2 * PCN_SM_COUNT;
    int waves = (int)((nc * 48 + slots / 2) / slots);
    if (waves < 1) waves = 1;
    int64_t p = (int64_t)slots * waves / nc;
    return (int)(p < 1 ? 1 : (p > AFR_PARTS ? AFR_PARTS : p));
}

bool aff_fused() {               // one kernel per layer and direction (default) or the GEMM + per-layer kernel pairs
    static int v = -1;
    if (v < 0) { const char* e = getenv("PCNERF_AFF_FUSED"); v = e ? (atoi(e) != 0) : 1; }
    return v != 0;
}

bool aff_moments_tc() {          // second moments on packed fp32 FFMA2 (default) or on mma.sync 3 x TF32 (PCNERF_AFF_MOMENTS=tc)
    static int v = -1;
    if (v < 0) { const char* e = getenv("PCNERF_AFF_MOMENTS"); v = (e && !strcmp(e, "tc")) ? 1 : 0; }
    return v != 0;
}

int aff_attrs() {
    static bool done = false;
    if (!done) {
        PCN_CUDA(cudaFuncSetAttribute(k_affine_moments_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, AFR_MOM_SMEM));
        PCN_CUDA(cudaFuncSetAttribute(k_affine_moments_rays_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, AMT_SMEM));
        PCN_CUDA(cudaFuncSetAttribute(k_aff_fwd_layer, cudaFuncAttributeMaxDynamicSharedMemorySize, AFL_SMEM_FWD));
        PCN_CUDA(cudaFuncSetAttribute(k_aff_bwd_layer, cudaFuncAttributeMaxDynamicSharedMemorySize, AFL_SMEM_BWD));
        done = true;
    }
    return 0;
}

const int kLd[8] = {63, 256, 256, 256, 319, 256, 256, 256};       // leading dimension of W_l
const int kOff[8] = {0, 0, 0, 0, 63, 0, 0, 0};                    // first column of the block that multiplies H_{l-1}

}  // namespace

extern "C" size_t pcnerf_affine_work_bytes(int64_t nchunk) {
    return nchunk < 1 ? 0 : work_layout(nchunk).total * sizeof(double);
}

extern "C" int pcnerf_affine_forward_rays(const pcnerf_mlp_params* P, const float* rays, int ld, int64_t n_rays,
                                          const float* z, int S, int64_t chunk, float* out_p, void* work, size_t work_bytes,
                                          void* stream) {
    int rc = check_common(P, rays, ld, n_rays, z, S, chunk, work, work_bytes, "affine_forward_rays");
    if (rc) return rc;
    PCN_CHECK_ARG(out_p, "affine_forward_rays: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = n_rays * S, nc = pcn_cdiv(rows, chunk);
    const Work w = work_layout(nc);
    double* base = (double*)work;
    const RayRows src{rays, ld, z, S};
    const int parts = parts_for(nc);
    rc = aff_attrs();
    if (rc) return rc;
    const bool fused = aff_fused();
    {
        // work: the FMAs of the upper-triangle outer products (36 tiles x 64 per row), 2 FLOP each
        PcnScope ps(PCN_K_AFFINE_MOMENTS, st, (double)rows * (AFR_TILES * 64) * 2.0, 3);
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("PCNERF_AFF_DEBUG"); dbg = e ? atoi(e) : 0; }
        if (aff_moments_tc())
            k_affine_moments_rays_tc<<<dim3(parts, (unsigned)nc), 256, AMT_SMEM, st>>>(src, rows, chunk, base + w.part,
                                                                                          base + w.shift, dbg);
        else
            k_affine_moments_rays<<<dim3(parts, (unsigned)nc), AFR_THREADS, AFR_MOM_SMEM, st>>>(src, rows, chunk, base + w.part,
                                                                                                    base + w.shift, dbg);
        PCN_LAUNCH_CHECK();
        k_affine_moments_reduce<<<dim3(16, (unsigned)nc), 256, 0, st>>>(base + w.part, parts, base + w.G0);    // (G0: free until backward)
        PCN_LAUNCH_CHECK();
        k_affine_moments_finish<<<(unsigned)nc, 256, 0, st>>>(base + w.G0, base + w.shift, base + w.m, base + w.C,
                                                              base + w.cnt);
        PCN_LAUNCH_CHECK();
    }
    const int N = (int)nc * 64;
    AffRunning run;
    {
        // work: float64 FLOPs of the seven 256 x 256 x (nc 64) products and of A C (256 x 64 x 64 per layer and chunk)
        PcnScope ps(PCN_K_AFFINE_ALGEBRA, st, 7.0 * 2.0 * 256 * 256 * N + 8.0 * 2.0 * 256 * 64 * N, fused ? 10 : 17);
        for (int l = 0; l < 8; ++l) {
            const LayerView v = layer_view(base, w, l, nc);
            if (l > 0 && !fused) {
                const LayerView pv = layer_view(base, w, l - 1, nc);
                k_aff_dgemm<float, true, false><<<dim3(N / 32, 8, 1), 128, 0, st>>>(P->W[l] + kOff[l], kLd[l], 1, pv.Ab, N, 1, v.A,
                                                                                    N, 0, 256, 256);
                PCN_LAUNCH_CHECK();
            }
            AffLayer g;
            g.A = v.A; g.AC = v.AC; g.Ab = v.Ab; g.a = v.a; g.r = v.r; g.mean = v.mean; g.var = v.var;
            g.m = base + w.m; g.C = base + w.C;
            g.Wx = (l == 0 || l == 4) ? P->W[l] : nullptr;
            g.wx_ld = kLd[l];
            g.init = l == 0;
            g.bias = P->b[l]; g.gamma = P->gamma[l]; g.beta = P->beta[l];
            g.rmean = nullptr; g.rvar = nullptr;
            g.nc = (int)nc; g.eps = (double)P->eps;
            if (fused)
                k_aff_fwd_layer<<<dim3(256 / AFL_ROWS, (unsigned)nc), 256, AFL_SMEM_FWD, st>>>(
                    g, P->W[l] + kOff[l], kLd[l], l > 0 ? layer_view(base, w, l - 1, nc).Ab : nullptr);
            else
                k_aff_layer_fwd<<<dim3(16, (unsigned)nc), 256, 0, st>>>(g);
            PCN_LAUNCH_CHECK();
            run.rm[l] = P->running_mean[l]; run.rv[l] = P->running_var[l]; run.nbt[l] = P->num_batches_tracked[l];
            run.mean[l] = v.mean; run.var[l] = v.var;
        }
        run.cnt = base + w.cnt; run.nc = (int)nc; run.mom = (double)P->momentum;
        k_aff_running<<<8, 256, 0, st>>>(run);
        PCN_LAUNCH_CHECK();
        k_aff_alpha<<<(unsigned)nc, 256, 0, st>>>(layer_view(base, w, 7, nc).Ab, P->W[8], P->b[8], (int)nc,
                                                  (float*)(base + w.alpha));
        PCN_LAUNCH_CHECK();
    }
    {
        PcnScope ps(PCN_K_AFFINE, st, (double)rows * 8.0);
        k_affine_apply_rays<<<dim3(parts, (unsigned)nc), 256, 0, st>>>(src, rows, chunk, (const float*)(base + w.alpha), out_p);
        PCN_LAUNCH_CHECK();
    }
    return 0;
}

// Eval mode (model.eval(): BatchNorm on its running statistics): alpha is a function of the parameters alone -- the same
// chain as above with one "chunk" and (mean, var) = (running_mean, running_var); callers cache it per parameter version.
extern "C" int pcnerf_affine_eval_alpha(const pcnerf_mlp_params* P, float* alpha, void* work, size_t work_bytes, void* stream) {
    PCN_CHECK_ARG(P && alpha && work, "affine_eval_alpha: null argument");
    PCN_CHECK_ARG(work_bytes >= work_layout(1).total * sizeof(double), "affine_eval_alpha: work area too small");
    cudaStream_t st = (cudaStream_t)stream;
    const Work w = work_layout(1);
    double* base = (double*)work;
    int rc = aff_attrs();
    if (rc) return rc;
    const bool fused = aff_fused();
    PcnScope ps(PCN_K_AFFINE_ALGEBRA, st, 7.0 * 2.0 * 256 * 256 * 64, fused ? 9 : 16);
    for (int l = 0; l < 8; ++l) {
        const LayerView v = layer_view(base, w, l, 1);
        if (l > 0 && !fused) {
            const LayerView pv = layer_view(base, w, l - 1, 1);
            k_aff_dgemm<float, true, false><<<dim3(2, 8, 1), 128, 0, st>>>(P->W[l] + kOff[l], kLd[l], 1, pv.Ab, 64, 1, v.A, 64, 0,
                                                                           256, 256);
            PCN_LAUNCH_CHECK();
        }
        AffLayer g;
        g.A = v.A; g.AC = v.AC; g.Ab = v.Ab; g.a = v.a; g.r = v.r; g.mean = v.mean; g.var = v.var;
        g.m = nullptr; g.C = nullptr;
        g.Wx = (l == 0 || l == 4) ? P->W[l] : nullptr;
        g.wx_ld = kLd[l];
        g.init = l == 0;
        g.bias = P->b[l]; g.gamma = P->gamma[l]; g.beta = P->beta[l];
        g.rmean = P->running_mean[l]; g.rvar = P->running_var[l];
        g.nc = 1; g.eps = (double)P->eps;
        if (fused)
            k_aff_fwd_layer<<<dim3(256 / AFL_ROWS, 1), 256, AFL_SMEM_FWD, st>>>(g, P->W[l] + kOff[l], kLd[l],
                                                                                l > 0 ? layer_view(base, w, l - 1, 1).Ab : nullptr);
        else
            k_aff_layer_fwd<<<dim3(16, 1), 256, 0, st>>>(g);
        PCN_LAUNCH_CHECK();
    }
    k_aff_alpha<<<1, 256, 0, st>>>(layer_view(base, w, 7, 1).Ab, P->W[8], P->b[8], 1, alpha);
    PCN_LAUNCH_CHECK();
    return 0;
}

// p_r = sigmoid(alpha . (embed(o + d z_r), 1)) for every (ray, depth) row: the eval-mode closed-form MLP of a sampling pass
// without an encoding tensor (alpha: 64 floats on the device, from pcnerf_affine_eval_alpha).
extern "C" int pcnerf_affine_apply_rays(const float* rays, int ld, int64_t n_rays, const float* z, int S, const float* alpha,
                                        float* out_p, void* stream) {
    PCN_CHECK_ARG(rays && z && alpha && out_p, "affine_apply_rays: null argument");
    PCN_CHECK_ARG(ld >= 6 && n_rays >= 0 && S >= 1, "affine_apply_rays: bad shape");
    const int64_t rows = n_rays * S;
    PCN_CHECK_ARG(rows < ((int64_t)1 << 31), "affine_apply_rays: %lld rows (limit 2^31 - 1)", (long long)rows);
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const RayRows src{rays, ld, z, S};
    const int64_t want = pcn_cdiv(rows, 512);               // at least two rows per thread
    const int parts = (int)(want < 1 ? 1 : (want > 8 * PCN_SM_COUNT ? 8 * PCN_SM_COUNT : want));
    PcnScope ps(PCN_K_AFFINE, st, (double)rows * 8.0);
    k_affine_apply_rays<<<dim3(parts, 1), 256, 0, st>>>(src, rows, rows, alpha, out_p);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_affine_backward_rays(const pcnerf_mlp_params* P, const pcnerf_mlp_grads* Gr, const float* rays, int ld,
                                           int64_t n_rays, const float* z, int S, int64_t chunk, const float* out_p,
                                           const float* grad_p, void* work, size_t work_bytes, void* stream) {
    int rc = check_common(P, rays, ld, n_rays, z, S, chunk, work, work_bytes, "affine_backward_rays");
    if (rc) return rc;
    PCN_CHECK_ARG(Gr && out_p && grad_p, "affine_backward_rays: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = n_rays * S, nc = pcn_cdiv(rows, chunk);
    const Work w = work_layout(nc);
    double* base = (double*)work;
    const RayRows src{rays, ld, z, S};
    const int N = (int)nc * 64, parts = parts_for(nc);
    {
        PcnScope ps(PCN_K_AFFINE, st, (double)rows * 16.0, 2);
        k_affine_grad_rays<<<dim3(parts, (unsigned)nc), 256, 0, st>>>(src, out_p, grad_p, rows, chunk, base + w.part);
        PCN_LAUNCH_CHECK();
        k_affine_grad_finish<<<(unsigned)nc, 64, 0, st>>>(base + w.part, parts, base + w.dalpha);
        PCN_LAUNCH_CHECK();
    }
    const bool fused = aff_fused();
    PcnScope ps(PCN_K_AFFINE_ALGEBRA, st, 14.0 * 2.0 * 256 * 256 * N, fused ? 12 : 45);       // two products per layer (dW, dAb)
    double* G = base + w.G0;
    double* Gn = base + w.G1;
    k_aff_head_bwd<<<256, 64, 0, st>>>(layer_view(base, w, 7, nc).Ab, base + w.dalpha, P->W[8], (int)nc, G, Gr->dW[8], Gr->db[8]);
    PCN_LAUNCH_CHECK();
    // chunks per split of the weight-gradient GEMM (K = nc * 64)
    // (the batched launch covers seven layers: two splits already give 7 x 64 x 2 = 896 blocks = two waves of three per SM, and
    // every warp runs 12 k steps instead of 4 at nc = 24)
    static int spl_b = -1;
    if (spl_b < 0) { const char* e = getenv("PCNERF_AFF_SPLITS"); spl_b = e ? atoi(e) : AFR_SPLITS_BATCHED; if (spl_b < 1 || spl_b > AFR_SPLITS) spl_b = AFR_SPLITS_BATCHED; }
    const int cps = (int)pcn_cdiv(nc, fused ? spl_b : AFR_SPLITS), splits = (int)pcn_cdiv(nc, cps);
    if (fused) {
        rc = aff_attrs();
        if (rc) return rc;
        AffWgradBatch wb;
        AffWgradOut wo;
        AffColsumBatch cb;
        for (int l = 7; l >= 0; --l) {
            const LayerView v = layer_view(base, w, l, nc);
            AffLayerBwd g;
            g.G = nullptr; g.Gin = l == 7 ? G : nullptr; g.Gout = base + w.Gl[l];
            g.A = v.A; g.AC = v.AC; g.a = v.a; g.r = v.r; g.mean = v.mean; g.var = v.var; g.m = base + w.m;
            g.gamma = P->gamma[l]; g.gpart = base + w.gpl[l]; g.bpart = base + w.bpl[l]; g.nc = (int)nc;
            k_aff_bwd_layer<<<dim3(256 / AFL_ROWS, (unsigned)nc), 256, AFL_SMEM_BWD, st>>>(
                g, l < 7 ? P->W[l + 1] + kOff[l + 1] : nullptr, l < 7 ? kLd[l + 1] : 0, l < 7 ? base + w.Gl[l + 1] : nullptr);
            PCN_LAUNCH_CHECK();
            cb.G[l] = base + w.Gl[l]; cb.gpart[l] = base + w.gpl[l]; cb.bpart[l] = base + w.bpl[l];
            cb.dgamma[l] = Gr->dgamma[l]; cb.dbeta[l] = Gr->dbeta[l]; cb.dbias[l] = Gr->db[l];
            cb.dWx[l] = (l == 0 || l == 4) ? Gr->dW[l] : nullptr;
            cb.wx_ld[l] = kLd[l];
            if (l > 0) {
                wb.G[l - 1] = base + w.Gl[l]; wb.Ab[l - 1] = layer_view(base, w, l - 1, nc).Ab;
                wo.dW[l - 1] = Gr->dW[l] + kOff[l]; wo.ld[l - 1] = kLd[l];
            }
        }
        // dW'_l += sum_chunks dL/dA_l Ab_{l-1}^T for l = 1..7, and the column sums of every layer
        k_aff_wgrad_batched<<<dim3(8, 8, 7 * splits), 128, 0, st>>>(wb, N, splits, cps * 64, base + w.wsplit);
        PCN_LAUNCH_CHECK();
        k_aff_wgrad_reduce_batched<<<dim3(256, 7), 256, 0, st>>>(base + w.wsplit, splits, wo);
        PCN_LAUNCH_CHECK();
        k_aff_colsum_batched<<<dim3(256, 8), 64, 0, st>>>(cb, (int)nc);
        PCN_LAUNCH_CHECK();
        return 0;
    }
    for (int l = 7; l >= 0; --l) {
        const LayerView v = layer_view(base, w, l, nc);
        AffLayerBwd g;
        g.G = G; g.Gin = nullptr; g.Gout = nullptr;
        g.A = v.A; g.AC = v.AC; g.a = v.a; g.r = v.r; g.mean = v.mean; g.var = v.var; g.m = base + w.m;
        g.gamma = P->gamma[l]; g.gpart = base + w.gpart; g.bpart = base + w.bpart; g.nc = (int)nc;
        k_aff_layer_bwd<<<dim3(16, (unsigned)nc), 256, 0, st>>>(g);
        PCN_LAUNCH_CHECK();
        k_aff_colsum_bwd<<<256, 64, 0, st>>>(G, base + w.gpart, base + w.bpart, (int)nc, Gr->dgamma[l], Gr->dbeta[l], Gr->db[l],
                                             (l == 0 || l == 4) ? Gr->dW[l] : nullptr, kLd[l]);
        PCN_LAUNCH_CHECK();
        if (l == 0) break;
        const LayerView pv = layer_view(base, w, l - 1, nc);
        // dW'_l += sum_chunks G Ab_{l-1}^T
        k_aff_dgemm<double, true, true><<<dim3(8, 8, splits), 128, 0, st>>>(G, N, 1, pv.Ab, 1, N, base + w.wsplit, 256, 65536, N,
                                                                             cps * 64);
        PCN_LAUNCH_CHECK();
        k_aff_wgrad_reduce<<<256, 256, 0, st>>>(base + w.wsplit, splits, Gr->dW[l] + kOff[l], kLd[l]);
        PCN_LAUNCH_CHECK();
        // dL/dAb_{l-1} = W'_l^T G
        k_aff_dgemm<float, false, false><<<dim3(N / 32, 8, 1), 128, 0, st>>>(P->W[l] + kOff[l], 1, kLd[l], G, N, 1, Gn, N, 0, 256,
                                                                             256);
        PCN_LAUNCH_CHECK();
        double* t = G; G = Gn; Gn = t;
    }
    return 0;
}
