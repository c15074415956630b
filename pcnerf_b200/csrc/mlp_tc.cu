// K3 (tensor-core path, precision 1): the occupancy MLP of nof/networks/models.py:125-203 on tcgen05 / TMEM / TMA.
//
// Same algebra as mlp.cu (BN(l) folded into Linear(l+1), batch statistics from the GEMM epilogue), different engine:
//   * every Linear is one persistent, warp-specialised kernel: warp 0 issues TMA loads (SWIZZLE_128B tiles), warp 1
//     issues tcgen05.mma (kind::f16, M=128 x N=256 x K=16 per instruction, fp32 accumulators in TMEM, two 256-column
//     accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1), warps 2..9 are the epilogue
//     (tcgen05.ld -> bias -> 16-bit store -> per-column batch statistics by a shuffle transpose-reduce);
//   * the layer's weights (256 x K, up to 160 KB) stay resident in shared memory for the life of the CTA, only
//     the activations stream through a 3-4 stage ring;
//   * weight gradients are a split-K GEMM over the rows with both operands MN-major (read straight from the row-major
//     activation / gradient matrices), 256x256 fp32 accumulators = all 512 TMEM columns, reduced with vector atomics.
// Number formats (DESIGN.md "precision"): forward operands fp16 (10-bit mantissa: bf16 storage misses the 1e-3 depth
// gate, measured 3-5e-3), gradients bf16 (range), accumulation and statistics fp32 / fp64.
#include "common.cuh"
#include "mlp_layout.h"
#include "mlp_small.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <type_traits>

#define TC_THREADS 320           // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..9: epilogue
#define TC_A_BYTES (128 * 128)   // one A k-block: 128 rows x 64 x 2 B
#define TC_NCTA 128              // output columns per CTA of the row GEMM (two CTAs share a row tile, see k_tc_rowgemm)
#define TC_B_BYTES (TC_NCTA * 128)   // one B k-block: 128 weight rows x 64 x 2 B
#define TC_WG_STAGE (64 * 1024)  // wgrad stage: A 4 boxes (32 KB) + B up to 4 boxes (32 KB)
#define TC_TIMEOUT_CYCLES 6000000000LL

__device__ int g_tc_err;         // set (before a trap) when a pipeline wait times out: a bug, never a valid state

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u) /* suspend-time hint: the warp may sleep until the phase completes */
        : "memory");
    return ok != 0;
}
// latency-critical hand-offs (MMA <-> epilogue): plain try_wait polling, no suspend hint
__device__ __forceinline__ bool mbar_try_wait_spin(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity, int code) {
    if (mbar_try_wait_spin(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_spin(bar, parity)) {
        if (clock64() - t0 > TC_TIMEOUT_CYCLES) {
            atomicExch(&g_tc_err, code);
            __threadfence_system();
            __trap();
        }
    }
}
// (spin != 0: plain polling instead of the suspend-hinted wait -- PCNERF_TC_DEBUG & 8, timing experiments)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code, int spin);
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > TC_TIMEOUT_CYCLES) {
            atomicExch(&g_tc_err, code);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code, int spin) {
    if (spin) mbar_wait_spin(bar, parity, code);
    else mbar_wait(bar, parity, code);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// same, with an L2 eviction-priority hint (policy constants as CUTLASS's CacheHintSm90 encodes them)
#define TC_L2_EVICT_FIRST 0x12F0000000000000ull
#define TC_L2_EVICT_LAST 0x14F0000000000000ull
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d_nocommit(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (quadrant*32 + t), registers = columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B (layout_type 2), descriptor version 1 (Blackwell).
//   K-major operand  (rows of 128 B = 64 elements along K): LBO unused (1), SBO = 1024 B between 8-row groups.
//   MN-major operand (rows of 128 B = 64 elements along M/N, one row per k): LBO = bytes between 64-element M/N
//   blocks, SBO = 1024 B between groups of 8 k.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// instruction descriptor for kind::f16: D = fp32; formats 0 = f16, 1 = bf16; major 0 = K, 1 = MN
__host__ __device__ constexpr uint32_t make_idesc(int afmt, int bfmt, int amaj, int bmaj, int M, int N) {
    return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)amaj << 15) | ((uint32_t)bmaj << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

enum { TC_FWD = 0, TC_DGRAD = 1, TC_DGRAD2 = 2 };     // DGRAD2: pair form with the BN-backward H term on the tensor core

struct RowGemmArgs {
    int rows;
    int kb0, kb_total;          // 64-wide k-blocks taken from A0 / in total (A1 supplies the rest)
    int a0_blocks;              // distinct 64-column blocks of A0: k-block kb < kb0 reads block kb % a0_blocks (the encoding
                                // tile is read once per weight block that multiplies it: [W_hi | W_lo] split, correction block)
    int nstage;
    void* out;                  // [rows][256]: fp16 (FWD) or bf16 (DGRAD)
    __nv_bfloat16* out2;        // FWD: optional bf16 copy of `out` (what the backward pass reads)
    const float* vec;           // FWD: bias[256];  DGRAD: c0 | c1 | c2 | mean, [4][256]
    const __half* E;            // DGRAD: H_{l-1} [rows][256] (fp16, as the forward pass stored it)
    double* stat0;              // [256] column sums of `out` (written once, by the last CTA to finish)
    double* stat1;              // FWD: [256] column sums of squares
    double* partials;           // [gridDim.x][2][128] per-CTA column sums: no same-address fp64 atomics (296 adds per
                                // address at kernel exit serialise in the L2 atomic unit)
    unsigned int* counter;      // zeroed by the caller; counts finished CTAs
    int debug;                  // timing experiments only (PCNERF_TC_DEBUG): 1 = no statistics, 2 = no output stores, 8 = polling waits
    int sched;                  // row-tile order of a CTA pair: 0 interleaved (pair, pair + npairs, ...), 1 contiguous slab
                                // walked upwards, 2 contiguous slab walked downwards (see tile_plan)
    int nostat;                 // 1: eval-mode BN (running statistics): the epilogue skips the column sums
    int stage_pairs;            // pair form: staging-buffer pairs per epilogue warp (2 = one per 64-column round, 1 = shared)
    int hint;                   // 1: the A tiles are not read again soon -> L2 evict_first, so that the freshly WRITTEN
                                // output matrix is what stays in the 126 MB L2 for the next kernel
};

// Which row tiles a CTA pair processes, in which order.  Consecutive kernels of a chunk walk their slabs in opposite
// directions (forward layer l reads what layer l-1 wrote last first; the data-gradient GEMM re-reads DH_l / H_{l-1} in the
// reverse of the weight-gradient kernel's order), so the tail of the previous kernel's traffic is still in L2.
struct TilePlan { int first, step, count; };
__device__ __forceinline__ TilePlan tile_plan(int sched, int ntiles, int pair, int npairs) {
    TilePlan p;
    if (sched == 0) {
        p.first = pair; p.step = npairs; p.count = pair < ntiles ? (ntiles - 1 - pair) / npairs + 1 : 0;
    } else {
        const int per = (ntiles + npairs - 1) / npairs;
        const int lo = pair * per, hi = min(ntiles, lo + per);
        p.count = hi > lo ? hi - lo : 0;
        if (sched == 1) { p.first = lo; p.step = 1; }
        else { p.first = hi - 1; p.step = -1; }
    }
    return p;
}

#define TC_STAGE_BYTES 2048     // one epilogue staging buffer: 32 rows x 64 B (32 x 16-bit), SWIZZLE_64B like its TMA box
#define TC_NBUF 2               // staging buffers per epilogue warp
#define TC_PAIRS_DEFAULT 7      // row-GEMM form: bit 0 = forward on CTA pairs, bit 1 = data gradient on CTA pairs
#define TC_NBUF2 4              // ... of the CTA-pair form (two 64-column rounds per tile, a pair of buffers each)

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// explicit shared-space accesses (32-bit shared addresses): keeps the staging traffic on LDS/STS instead of generic LD/ST
__device__ __forceinline__ uint32_t lds32u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// Staging buffer layout = the SWIZZLE_64B layout of a {32 columns x 32 rows} 16-bit TMA box: row r at r*64 B, its 16-byte
// chunk k at position k ^ ((r >> 1) & 3).  Row-per-thread 16-byte accesses and the column walks below are conflict-free.
__device__ __forceinline__ uint32_t stage_addr(uint32_t buf, int row, int chunk) {
    return buf + (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
// this thread's row (pk[16] = 32 packed 16-bit values) -> staging buffer
__device__ __forceinline__ void stage_put_row(uint32_t buf, const uint32_t* pk, int lane) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sts128(stage_addr(buf, lane, k), make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]));
}
// Column sums (and sums of squares) over the 32 rows of a staged 32 x 32 chunk of 16-bit values.  Lane l walks the
// 32-bit word (l % 16) = columns 2(l%16), 2(l%16)+1 of the even rows (l < 16) or the odd rows (l >= 16): 32 distinct banks
// per access.  After the xor-16 shuffle lanes l and l^16 both hold the totals of their two columns.
template <bool HALF, bool SQ>
__device__ __forceinline__ void stage_col_sums(uint32_t buf, int lane, float* acc0, float* acc1) {
    const int w = lane & 15, par = lane >> 4;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int row = 2 * i + par;
        const uint32_t u = lds32u(stage_addr(buf, row, w >> 2) + (uint32_t)((w & 3) << 2));
        float2 f;
        if (HALF) f = __half22float2(*reinterpret_cast<const __half2*>(&u));
        else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
        s0 += f.x;
        s1 += f.y;
        if (SQ) { q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1); }
    }
    // Per-lane running sums stay in fp32 (a CTA adds <= 30 tile partials of 32 values each: ~1e-6 relative, far below the
    // fp16 rounding of the values themselves); the cross-quadrant / cross-CTA reduction is fp64.  fp64 adds here showed up
    // as stall_math in the epilogue (ncu: DADD 5 % of the stall samples of k_tc_rowgemm2).
    s0 += __shfl_xor_sync(FULL_MASK, s0, 16);
    s1 += __shfl_xor_sync(FULL_MASK, s1, 16);
    acc0[0] += s0;
    acc0[1] += s1;
    if (SQ) {
        q0 += __shfl_xor_sync(FULL_MASK, q0, 16);
        q1 += __shfl_xor_sync(FULL_MASK, q1, 16);
        acc1[0] += q0;
        acc1[1] += q1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// C[rows,256] = A[rows,K] * B[256,K]^T     (A, B K-major)
//   FWD  : fp16 x fp16; out = fp16(C + bias), out2 = bf16(C + bias); stats = column sums of out and out^2 (the fp16
//          values the next layer consumes)
//   DGRAD: bf16 x bf16; out = bf16(c0*C - c1 - (E - mean)*c2)  (BN backward fused, E fp16); stat0 = column sums of out
// Outputs leave through per-warp TMA stores of {32 x 32} boxes; rows beyond `rows` are clipped by the tensor map.
// Work split: CTA b owns the column half (b & 1) of row tiles (b >> 1), (b >> 1) + gridDim.x/2, ...  Its 128 x K slice of
// the weights (<= 80 KB) stays resident, which leaves room for an 7-8 stage ring of A tiles: the kernel is bound by bytes
// in flight (measured: a one-tile ring sustained 2.5 TB/s regardless of the epilogue), and the twin CTA's read of the same
// A tile hits L2.
// ---------------------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_rowgemm(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
             const __grid_constant__ CUtensorMap tmO2, const RowGemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[24];            // full[8] | empty[8] | bfull | tfull[2] | tempty[2]
    __shared__ uint32_t tmem_slot;
    // FWD: bias[256].  DGRAD: c0 | c2 | k, k = c2*mean - c1, so that DH = c0*G - c2*H + k (two FMAs per element)
    __shared__ __align__(16) float cvec[(EPI == TC_FWD ? 1 : 3) * 256];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstage = g.nstage, KB = g.kb_total;
    uint8_t* sB = smem;                                   // KB x 32 KB, resident
    uint8_t* sA = sB + (size_t)KB * TC_B_BYTES;           // nstage x 16 KB ring
    uint8_t* sStage = sA + (size_t)nstage * TC_A_BYTES;   // 8 warps x TC_NBUF x TC_STAGE_BYTES
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 8), bar_bfull = smem_u32(bars + 16);
    const uint32_t bar_tfull = smem_u32(bars + 17), bar_tempty = smem_u32(bars + 19);

    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_bfull, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, 8); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        tma_prefetch_desc(&tmO2);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
    if (EPI == TC_FWD) {
        for (int i = threadIdx.x; i < 256; i += TC_THREADS) cvec[i] = g.vec[i];
    } else {
        for (int i = threadIdx.x; i < 256; i += TC_THREADS) {
            const float c0 = g.vec[i], c1 = g.vec[256 + i], c2 = g.vec[512 + i], mean = g.vec[768 + i];
            cvec[i] = c0;
            cvec[256 + i] = c2;
            cvec[512 + i] = c2 * mean - c1;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int ntiles = (g.rows + 127) >> 7;
    const int nhalf = blockIdx.x & 1;                     // which 128 output columns this CTA owns
    const TilePlan tp = tile_plan(g.sched, ntiles, blockIdx.x >> 1, gridDim.x >> 1);

    if (warp == 0) {
        // ===== TMA producer
        if (lane == 0) {
            mbar_expect_tx(bar_bfull, (uint32_t)KB * TC_B_BYTES);
            for (int kb = 0; kb < KB; ++kb)
                tma_load_2d(smem_u32(sB + (size_t)kb * TC_B_BYTES), &tmB, bar_bfull, kb * 64, nhalf * TC_NCTA);
        }
        // (an extra `cp.async.bulk.prefetch.tensor` of the tiles after the ring into L2 was measured 10 % slower)
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < tp.count; ++it) {
            const int tile = tp.first + it * tp.step;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(bar_empty + 8 * s, ph ^ 1, 1, g.debug & 8);
                if (lane == 0) {
                    mbar_expect_tx(bar_full + 8 * s, TC_A_BYTES);
                    const CUtensorMap* m = kb < g.kb0 ? &tmA0 : &tmA1;
                    const int c0 = (kb < g.kb0 ? kb % g.a0_blocks : kb - g.kb0) * 64;
                    // A0 next to an A1 = the chunk's encoding, which EVERY layer of the chunk reads (correction block /
                    // skip connection): keep its 33.5 MB in L2; the activations are read once: evict first
                    if (g.hint) tma_load_2d_hint(smem_u32(sA + (size_t)s * TC_A_BYTES), m, bar_full + 8 * s, c0, tile * 128,
                                                 (kb < g.kb0 && g.kb_total > g.kb0) ? TC_L2_EVICT_LAST : TC_L2_EVICT_FIRST);
                    else tma_load_2d(smem_u32(sA + (size_t)s * TC_A_BYTES), m, bar_full + 8 * s, c0, tile * 128);
                }
                __syncwarp();
                if (++s == nstage) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        constexpr uint32_t idesc = (EPI == TC_FWD) ? make_idesc(0, 0, 0, 0, 128, TC_NCTA) : make_idesc(1, 1, 0, 0, 128, TC_NCTA);
        mbar_wait(bar_bfull, 0, 2);
        int s = 0, as = 0;
        uint32_t ph = 0, aph = 0;
        for (int it = 0; it < tp.count; ++it) {
            mbar_wait_spin(bar_tempty + 8 * as, aph ^ 1, 3);
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)as * TC_NCTA;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait_spin(bar_full + 8 * s, ph, 4);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a0 = smem_u32(sA + (size_t)s * TC_A_BYTES), b0 = smem_u32(sB + (size_t)kb * TC_B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(dcol, make_desc(a0 + k * 32, 16, 1024), make_desc(b0 + k * 32, 16, 1024), idesc,
                                 (uint32_t)((kb | k) != 0));
                    umma_commit(bar_empty + 8 * s);
                    if (kb == KB - 1) umma_commit(bar_tfull + 8 * as);
                }
                __syncwarp();
                if (++s == nstage) { s = 0; ph ^= 1; }
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    } else {
        // ===== epilogue: warp (2..9) -> TMEM lane quadrant warp%4, 64-column half (warp-2)/4 of the CTA's 128 columns.
        // Both 32-column chunks of a tile are handled together: one TMEM wait (after which the accumulator stage goes
        // straight back to the MMA warp), one staging round trip, one bulk-store group per tile.
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int colb = nhalf * TC_NCTA + half * 64;     // first output column of this warp
        const uint32_t buf0 = smem_u32(sStage) + (uint32_t)(warp - 2) * (TC_NBUF * TC_STAGE_BYTES);
        const uint32_t bufs[2] = {buf0, buf0 + TC_STAGE_BYTES};
        float acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
        int as = 0;
        uint32_t aph = 0;
        // DGRAD: this warp's 32 rows x 128 B piece of H_{l-1} for the NEXT tile is in flight (8 rows x 64 B per load
        // instruction) while the current tile is processed
        uint4 e[2][4];
        auto load_e = [&](int tile_) {
            const int r0_ = tile_ * 128 + q * 32, left_ = g.rows - r0_;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = (lane >> 2) + 8 * i;
                    e[c][i] = make_uint4(0, 0, 0, 0);
                    if (row < left_)
                        e[c][i] = *reinterpret_cast<const uint4*>(g.E + (size_t)(r0_ + row) * 256 + colb + c * 32 + (lane & 3) * 8);
                }
        };
        if (EPI == TC_DGRAD && tp.count > 0) load_e(tp.first);
        for (int it = 0; it < tp.count; ++it) {
            const int tile = tp.first + it * tp.step;
            mbar_wait_spin(bar_tfull + 8 * as, aph, 5);
            tc_fence_after();
            const int row0 = tile * 128 + q * 32;
            const bool valid = row0 + lane < g.rows;
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TC_NCTA + half * 64);
            uint32_t r[2][32];
            tmem_ld32_issue(tbase, r[0]);
            tmem_ld32_issue(tbase + 32, r[1]);
            tmem_ld_wait();
            // every accumulator value of this stage is in registers: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_tempty + 8 * as);
                tma_store_wait_read<0>();                 // last tile's stores have read both staging buffers
            }
            __syncwarp();
            uint32_t pk[2][16];
            if (EPI == TC_FWD) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(&cvec[colb + c * 32 + 4 * j4]);
                        const float v0 = valid ? __uint_as_float(r[c][4 * j4 + 0]) + b4.x : 0.f;
                        const float v1 = valid ? __uint_as_float(r[c][4 * j4 + 1]) + b4.y : 0.f;
                        const float v2 = valid ? __uint_as_float(r[c][4 * j4 + 2]) + b4.z : 0.f;
                        const float v3 = valid ? __uint_as_float(r[c][4 * j4 + 3]) + b4.w : 0.f;
                        const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
                        pk[c][2 * j4] = *reinterpret_cast<const uint32_t*>(&h01);
                        pk[c][2 * j4 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    }
            } else {
                // this thread's row of H_{l-1}: staged through the (currently idle) output buffers, re-read row-wise
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 4; ++i) sts128(stage_addr(bufs[c], (lane >> 2) + 8 * i, lane & 3), e[c][i]);
                __syncwarp();
                if (it + 1 < tp.count) load_e(tile + tp.step);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint4 hv = lds128(stage_addr(bufs[c], lane, k));
                        const __half2* hb = reinterpret_cast<const __half2*>(&hv);
                        float hf[8];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float2 h2 = __half22float2(hb[t]);
                            hf[2 * t] = h2.x;
                            hf[2 * t + 1] = h2.y;
                        }
#pragma unroll
                        for (int t4 = 0; t4 < 2; ++t4) {
                            const int j = k * 8 + t4 * 4, col = colb + c * 32 + j;
                            const float4 a0 = *reinterpret_cast<const float4*>(&cvec[col]);
                            const float4 a2 = *reinterpret_cast<const float4*>(&cvec[256 + col]);
                            const float4 ak = *reinterpret_cast<const float4*>(&cvec[512 + col]);
                            const float v0 = valid ? fmaf(a0.x, __uint_as_float(r[c][j + 0]), fmaf(-a2.x, hf[t4 * 4 + 0], ak.x)) : 0.f;
                            const float v1 = valid ? fmaf(a0.y, __uint_as_float(r[c][j + 1]), fmaf(-a2.y, hf[t4 * 4 + 1], ak.y)) : 0.f;
                            const float v2 = valid ? fmaf(a0.z, __uint_as_float(r[c][j + 2]), fmaf(-a2.z, hf[t4 * 4 + 2], ak.z)) : 0.f;
                            const float v3 = valid ? fmaf(a0.w, __uint_as_float(r[c][j + 3]), fmaf(-a2.w, hf[t4 * 4 + 3], ak.w)) : 0.f;
                            const __nv_bfloat162 b01 = __floats2bfloat162_rn(v0, v1), b23 = __floats2bfloat162_rn(v2, v3);
                            pk[c][k * 4 + t4 * 2] = *reinterpret_cast<const uint32_t*>(&b01);
                            pk[c][k * 4 + t4 * 2 + 1] = *reinterpret_cast<const uint32_t*>(&b23);
                        }
                    }
                __syncwarp();                              // every lane has read its row of E: the buffers can be rewritten
            }
            stage_put_row(bufs[0], pk[0], lane);
            stage_put_row(bufs[1], pk[1], lane);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(g.debug & 2)) {
                tma_store_2d_nocommit(&tmO, bufs[0], colb, row0);
                tma_store_2d_nocommit(&tmO, bufs[1], colb + 32, row0);
                tma_store_commit();
            }
            if (!(g.debug & 1) && !g.nostat) {
                if (EPI == TC_FWD) {
                    stage_col_sums<true, true>(bufs[0], lane, acc0, acc1);
                    stage_col_sums<true, true>(bufs[1], lane, acc0 + 2, acc1 + 2);
                } else {
                    stage_col_sums<false, false>(bufs[0], lane, acc0, acc1);
                    stage_col_sums<false, false>(bufs[1], lane, acc0 + 2, acc1 + 2);
                }
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
        if (lane == 0) tma_store_wait_all();
        // ---- column statistics: quadrant warps -> CTA partial (staging memory is free now) -> global per-CTA slot;
        //      the last CTA to arrive adds the slots up.  No two CTAs ever touch the same address atomically.
        double* sred = reinterpret_cast<double*>(sStage);              // [2 stats][4 quadrants][128 columns]
        asm volatile("bar.sync 1, 256;" ::: "memory");                  // every warp's bulk stores have read the staging
        if (lane < 16) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int cl = half * 64 + c * 32 + 2 * lane + j;   // column within this CTA's 128
                    sred[(0 * 4 + q) * 128 + cl] = (double)acc0[2 * c + j];
                    sred[(1 * 4 + q) * 128 + cl] = (double)acc1[2 * c + j];
                }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = threadIdx.x - 64;                                 // 0..255
        {
            const int st = t >> 7, cl = t & 127;
            const double v = sred[(st * 4 + 0) * 128 + cl] + sred[(st * 4 + 1) * 128 + cl] + sred[(st * 4 + 2) * 128 + cl] +
                             sred[(st * 4 + 3) * 128 + cl];
            g.partials[((size_t)blockIdx.x * 2 + st) * 128 + cl] = v;
        }
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        __shared__ unsigned int s_last;
        if (t == 0) s_last = atomicAdd(g.counter, 1u) == gridDim.x - 1 ? 1u : 0u;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (s_last) {
            __threadfence();
            const int col = t, hsel = col >> 7, cl = col & 127;         // CTAs with (blockIdx & 1) == hsel own this column
            double a0 = 0.0, a1 = 0.0;
            for (int b = hsel; b < (int)gridDim.x; b += 2) {
                a0 += g.partials[((size_t)b * 2 + 0) * 128 + cl];
                if (EPI == TC_FWD) a1 += g.partials[((size_t)b * 2 + 1) * 128 + cl];
            }
            g.stat0[col] = a0;
            if (EPI == TC_FWD) g.stat1[col] = a1;
            if (t == 0) *g.counter = 0;                                 // ready for the next launch on this stream
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// out[256, ldo] (+)= sum_r DH[r, 0:256]^T X[r, 0:N]      (split-K over the rows; DH bf16, X fp16, both MN-major)
// ---------------------------------------------------------------------------------------------------------------
struct WgradArgs {
    int rows;
    int N;                      // 256 or 64 columns of X
    int kb_per_cta;             // 64-row k-blocks per CTA
    float* out;                 // [256][ldo] fp32, accumulated with vector atomics
    int ldo, col_off;
    int convert_b;              // X is fp16: the idle epilogue warps convert each B tile to bf16 in shared memory
    int debug;                  // timing experiments only (PCNERF_TC_DEBUG & 4: skip the atomic reduction)
};

__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_wgrad(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NST = 3;
    uint64_t* bars = (uint64_t*)(smem + (size_t)NST * TC_WG_STAGE);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 4), bar_tfull = smem_u32(bars + 8);
    const uint32_t bar_conv = smem_u32(bars + 10);
    uint32_t* tmem_slot = (uint32_t*)(bars + 9);
    const int nkb = (g.rows + 63) >> 6;
    const int kb_beg = blockIdx.x * g.kb_per_cta;
    const int kb_end = min(nkb, kb_beg + g.kb_per_cta);
    const int N = g.N, nboxB = N >> 6;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 8);
        }
        mbar_init(bar_tfull, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (kb_beg < kb_end) {
        if (warp == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int kb = kb_beg; kb < kb_end; ++kb) {
                mbar_wait(bar_empty + 8 * s, ph ^ 1, 11, g.debug & 8);
                if (lane == 0) {
                    uint8_t* st = smem + (size_t)s * TC_WG_STAGE;
                    mbar_expect_tx(bar_full + 8 * s, (uint32_t)(4 + nboxB) * 8192);
                    for (int j = 0; j < 4; ++j) tma_load_2d(smem_u32(st + j * 8192), &tmA, bar_full + 8 * s, j * 64, kb * 64);
                    for (int j = 0; j < nboxB; ++j)
                        tma_load_2d(smem_u32(st + 32768 + j * 8192), &tmB, bar_full + 8 * s, j * 64, kb * 64);
                }
                __syncwarp();
                if (++s == NST) { s = 0; ph ^= 1; }
            }
        } else if (warp == 1) {
            const uint32_t idesc = make_idesc(1, 1, 1, 1, 128, N);
            int s = 0;
            uint32_t ph = 0;
            for (int kb = kb_beg; kb < kb_end; ++kb) {
                mbar_wait((g.convert_b ? bar_conv : bar_full) + 8 * s, ph, 12, g.debug & 8);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a0 = smem_u32(smem + (size_t)s * TC_WG_STAGE), b0 = a0 + 32768;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = make_desc(b0 + k * 2048, 8192, 1024);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            umma_f16(tmem_base + (uint32_t)(h * N), make_desc(a0 + h * 16384 + k * 2048, 8192, 1024), bd, idesc,
                                     (uint32_t)((kb != kb_beg) | (k != 0)));
                    }
                    umma_commit(bar_empty + 8 * s);
                    if (kb == kb_end - 1) umma_commit(bar_tfull);
                }
                __syncwarp();
                if (++s == NST) { s = 0; ph ^= 1; }
            }
        } else {
            const int q = warp & 3, half = (warp - 2) >> 2;
            if (g.convert_b) {
                // tcgen05 kind::f16 cannot mix fp16 and bf16 operands, and the gradients need bf16's range: while the
                // mainloop runs these eight warps have nothing else to do, so they rewrite every freshly landed fp16
                // B tile as bf16 in place (same size, same swizzled position) before the MMA warp may touch it.
                const int t256 = threadIdx.x - 64, nch = N * 8;          // 16-byte chunks of the B tile
                int s = 0;
                uint32_t ph = 0;
                for (int kb = kb_beg; kb < kb_end; ++kb) {
                    mbar_wait(bar_full + 8 * s, ph, 14, g.debug & 8);
                    const uint32_t b0 = smem_u32(smem + (size_t)s * TC_WG_STAGE) + 32768;
                    for (int i = t256; i < nch; i += 256) {
                        uint4 v = lds128(b0 + (uint32_t)i * 16);
                        uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[t]));
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(f.x, f.y);
                            w[t] = *reinterpret_cast<const uint32_t*>(&b2);
                        }
                        sts128(b0 + (uint32_t)i * 16, v);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_conv + 8 * s);
                    if (++s == NST) { s = 0; ph ^= 1; }
                }
            }
            mbar_wait(bar_tfull, 0, 13);
            tc_fence_after();
            // Split-K reduction: 148 CTAs add their 256 x N fp32 partial into `out` with vector atomics -- 2.4 M 16-byte
            // reductions per launch, all at kernel exit, bound by the L2's atomic units (timing experiment: 4.4 of the
            // 18 ms/step of this kernel class).  (a) Lane pairs swap half of their chunks so that the two lanes of a pair
            // write ADJACENT 16-byte chunks of one row (one 32-byte sector per pair instead of two); (b) every CTA starts at
            // a different (h, c) block so that the CTAs do not walk the same addresses in lockstep.
            const int ncw = (N >> 5) >> 1;                              // 32-column blocks per epilogue warp and h: 4 or 1
            const int nblk = (g.debug & 4) ? 0 : 2 * ncw;
            const bool odd = lane & 1;
            for (int t0 = 0; t0 < nblk; ++t0) {
                const int tt = (t0 + (int)blockIdx.x) % nblk;
                const int h = tt / ncw, c = half + 2 * (tt % ncw);
                const int m_even = h * 128 + q * 32 + (lane & ~1), m_odd = m_even + 1;
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * N + c * 32), r);
                float* base = g.out + g.col_off + c * 32;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    // even lane keeps chunk 2t of its row and sends chunk 2t+1; odd lane keeps 2t+1 and sends 2t
                    float keep[4], recv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float lo = __uint_as_float(r[8 * t + i]), hi = __uint_as_float(r[8 * t + 4 + i]);
                        keep[i] = odd ? hi : lo;
                        recv[i] = __shfl_xor_sync(FULL_MASK, odd ? lo : hi, 1);
                    }
                    float* d0 = base + (size_t)m_even * g.ldo + (2 * t + (odd ? 1 : 0)) * 4;   // row of the even lane
                    float* d1 = base + (size_t)m_odd * g.ldo + (2 * t + (odd ? 1 : 0)) * 4;    // row of the odd lane
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d0), "f"(odd ? recv[0] : keep[0]),
                                 "f"(odd ? recv[1] : keep[1]), "f"(odd ? recv[2] : keep[2]), "f"(odd ? recv[3] : keep[3])
                                 : "memory");
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d1), "f"(odd ? keep[0] : recv[0]),
                                 "f"(odd ? keep[1] : recv[1]), "f"(odd ? keep[2] : recv[2]), "f"(odd ? keep[3] : recv[3])
                                 : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Eval-mode forward, all eight Linear(+folded BN) layers and the output layer in ONE persistent kernel.
//
// With running statistics every fold is known before the first GEMM, so nothing chunk-wide separates the layers: a CTA
// keeps the activations of its row tiles in shared memory from the encoding to the occupancy -- no H_l ever reaches HBM
// (the layered kernels move 1 KB per sample per layer).  Per CTA: two 128-row tiles X, Y in flight ("ping-pong"): while
// the eight epilogue warps turn the TMEM accumulator of (layer l, X) into the fp16 A operand of layer l+1 (written
// straight into the SWIZZLE_128B K-major layout a TMA load would have produced), the tensor core runs (layer l, Y).
//   shared memory: act[2] 2 x 64 KB | operand ring 6 x 16 KB = 224 KB;  TMEM: 2 x (128 lanes x 256 fp32 columns).
//   Everything that is not an activation streams through the ring in the order the MMA warp consumes it: 128 weight rows
//   x 64 k per stage (from the L2-resident folded fp16 weights, ~1 MB per tile pair) and, ahead of layers 0 and 4, the
//   tile's 128 x 64 encoding block (the A operand of layer 0 and of the skip columns of layer 4).  The ring is the
//   kernel's limiter: a stage feeds 256 tensor-core cycles and comes back ~1 us after it is released, so the bytes in flight
//   (not L2 bandwidth: ncu lts 21 %) set the pace -- hence every byte that is not act[] belongs to it.
//   warp 0: TMA producer;  warp 1: tcgen05.mma issuer;  warps 2..9: epilogue.
// ---------------------------------------------------------------------------------------------------------------
// ---- CTA-pair (cta_group::2) forms: the pair's leader (cluster rank 0) issues one MMA for both SMs (M = 256: 128 rows
// from each CTA's shared memory, B split -- each CTA holds half of the N weight rows), TMA loads of either CTA signal the
// LEADER's mbarrier, commits are multicast to the same barrier offset in both CTAs.  Within a pair the shared::cluster
// address of the peer's copy of a shared variable differs in bit 24 only (the even CTA has it clear).
#define TC_PEER_MASK 0xFEFFFFFFu
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(0x1000000000000000ull) /* evict normal */
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in both CTAs of the pair once every MMA issued so far has completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// arrive (release at cluster scope) on the LEADER's copy of a barrier, from either CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & TC_PEER_MASK) : "memory");
}
// The same arrive WITHOUT cluster-scope release semantics (default .release.cta, what CUTLASS's ClusterBarrier::arrive
// emits): for hand-offs that publish nothing through generic memory -- the TMEM-stage hand-back of k_tc_rowgemm2, whose
// ordering is carried by tcgen05.fence::before/after_thread_sync.  With .release.cluster every arrive compiled to a
// cluster-scope MEMBAR that waited for the warp's outstanding global loads (ncu on the first pair form: 25 % of all stall
// samples on ERRBAR / SYNCS.ARRIVE with stall_membar; the data-gradient kernel, whose epilogue keeps H_{l-1} loads in
// flight, was hit hardest: 23.8 vs 19.4 ms per step).
__device__ __forceinline__ void mbar_arrive_leader_cta(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & TC_PEER_MASK) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int code) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > TC_TIMEOUT_CYCLES) {
            atomicExch(&g_tc_err, code);
            __threadfence_system();
            __trap();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_tc_rowgemm on CTA PAIRS (cta_group::2): C[rows,256] = A[rows,K] * B[256,K]^T, same epilogues as k_tc_rowgemm.
//
// Why: timing experiments on the single-CTA form (profiles/README.md, v13) put its floor at ~55 us per 262,144-row launch
// = 2 us per tile, four times the tile's MMA time, with nothing saturated: the kernel is bound by the UNIQUE bytes it keeps
// in flight.  There both CTAs of a column-split pair stream the SAME A tiles (the second read is an L2 hit), so half of
// every ring is a duplicate.  Here the pair splits the M dimension instead: a unit of work is two row tiles, CTA `rank`
// loads only ITS 128 rows of A (every ring byte is unique, L2 -> SM traffic halves), each CTA keeps half of the weight
// rows resident (128 x K, as before), and the leader issues one M = 256, N = 256 MMA per 16 k for both SMs.  Each CTA's
// accumulator is its 128 rows x 256 columns (two stages = all 512 TMEM columns); its eight epilogue warps take 128
// columns each, in two rounds of 64 through the same staging buffers.
// Protocol = the one of k_tc_fused_eval<2>: TMA loads of either CTA complete on the LEADER's full barrier, commits are
// multicast to the same barrier offset in both CTAs, the sixteen epilogue warps report to the leader's tempty barrier.
// Selected by pcnerf_tc_set_row_pairs(1) / PCNERF_TC_PAIRS=1; NOT the default: parity-tested (tests/test_gpu_tc.py runs the
// GEMM and MLP checks in both forms) but measured SLOWER.  Same box, C2 step, class times per step (weight-gradient class
// 16.66 ms in both runs): k_tc_rowgemm forward 16.4 / data gradient 19.3 ms, step 56.5 ms at 1.65 GHz; this kernel 18.5 /
// 22.9 ms, step 59.9 ms at 1.77 GHz (the first version, which held the TMEM stage through round 0's stores and
// statistics, 19.0 / 20.8).  Halving the L2 -> SM traffic and doubling the unique bytes in flight does not shorten the
// ~2 us per tile -- neither is what bounds the row GEMMs; the per-CTA epilogue (two serial 64-column rounds per tile, 168
// registers) now sets the pace.  The SM clock under the power cap rises (less energy per step), which is why the form is
// kept as a starting point for round 2.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d_2sm_hint(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                                     uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
#define TC_L2_EVICT_NORMAL 0x1000000000000000ull

// NEW = epilogue warps per CTA: 8 (each takes 128 columns of its quadrant's 32 rows in two 64-column rounds) or 16 (64
// columns, one round; 576 threads, <= 113 registers).  The pair form is bound by its epilogue (ncu: the epilogue warps wait
// for the MMA 14 % of their time, 1.1 warp instructions issued per cycle and SM: latency of a long dependent chain per
// warp), so more warps per scheduler hide more of it.
template <int EPI, int NEW>
__global__ void __launch_bounds__(64 + 32 * NEW, 1)
k_tc_rowgemm2(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
              const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
              const __grid_constant__ CUtensorMap tmD, const RowGemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[24];            // full[8] | empty[8] | bfull | tfull[2] | tempty[2]
    __shared__ uint32_t tmem_slot;
    __shared__ unsigned int s_last;
    __shared__ __align__(16) float cvec[(EPI == TC_FWD ? 1 : 3) * 256];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstage = g.nstage, KB = g.kb_total;
    // resident B: this CTA's 128 weight rows, KBW k-blocks of 16 KB; DGRAD2: + its 32 rows of the four 64 x 64 diagonal blocks
    constexpr bool D2 = EPI == TC_DGRAD2;
    const int KBW = D2 ? g.kb0 : KB;
    uint8_t* sB = smem;
    uint8_t* sDg = sB + (size_t)KBW * TC_B_BYTES;         // DGRAD2: 4 x 4 KB
    uint8_t* sA = sDg + (D2 ? TC_B_BYTES : 0);            // nstage x 16 KB ring: this CTA's 128 rows of A
    uint8_t* sStage = sA + (size_t)nstage * TC_A_BYTES;   // 8 warps x TC_NBUF2 x TC_STAGE_BYTES
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 8), bar_bfull = smem_u32(bars + 16);
    const uint32_t bar_tfull = smem_u32(bars + 17), bar_tempty = smem_u32(bars + 19);
    const int rank = (int)cluster_ctarank();              // 0 = the pair's leader
    constexpr int ROUNDS = 16 / NEW;                      // 64-column rounds per epilogue warp and tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_bfull, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, 2 * NEW); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32(&tmem_slot), 512);
    if (EPI == TC_FWD) {
        for (int i = threadIdx.x; i < 256; i += (int)blockDim.x) cvec[i] = g.vec[i];
    } else if (D2) {
        for (int i = threadIdx.x; i < 512; i += (int)blockDim.x) cvec[i] = g.vec[i];       // 1 / t | k
    } else {
        for (int i = threadIdx.x; i < 256; i += (int)blockDim.x) {
            const float c0 = g.vec[i], c1 = g.vec[256 + i], c2 = g.vec[512 + i], mean = g.vec[768 + i];
            cvec[i] = c0;
            cvec[256 + i] = c2;
            cvec[512 + i] = c2 * mean - c1;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int ntiles = (g.rows + 127) >> 7, nunits = (ntiles + 1) >> 1;
    const TilePlan tp = tile_plan(g.sched, nunits, blockIdx.x >> 1, gridDim.x >> 1);   // in units of two row tiles
    const uint64_t pol = g.hint ? TC_L2_EVICT_FIRST : TC_L2_EVICT_NORMAL;

    if (warp == 0) {
        // ===== TMA producer (one per CTA: its own rows of A, its own half of B)
        if (lane == 0) {
            if (rank == 0) mbar_expect_tx(bar_bfull, 2u * ((uint32_t)KBW * TC_B_BYTES + (D2 ? 4u * 4096u : 0u)));
            for (int kb = 0; kb < KBW; ++kb)
                tma_load_2d_2sm_hint(smem_u32(sB + (size_t)kb * TC_B_BYTES), &tmB, bar_bfull & TC_PEER_MASK, kb * 64,
                                     rank * TC_NCTA, TC_L2_EVICT_NORMAL);
            if (D2)     // diagonal block kb: rows kb*64 + 32*rank .. +31 of Dg [256][64] (this CTA's half of the N = 64 operand)
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_2d_2sm_hint(smem_u32(sDg + (size_t)kb * 4096), &tmD, bar_bfull & TC_PEER_MASK, 0,
                                         kb * 64 + rank * 32, TC_L2_EVICT_NORMAL);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < tp.count; ++it) {
                const int tile = (tp.first + it * tp.step) * 2 + rank;     // rows past the end read as zero
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait_spin(bar_empty + 8 * s, ph ^ 1, 31);
                    if (rank == 0) mbar_expect_tx(bar_full + 8 * s, 2 * TC_A_BYTES);
                    const CUtensorMap* m = kb < g.kb0 ? &tmA0 : &tmA1;
                    const int c0 = (kb < g.kb0 ? kb % g.a0_blocks : kb - g.kb0) * 64;
                    tma_load_2d_2sm_hint(smem_u32(sA + (size_t)s * TC_A_BYTES), m, (bar_full + 8 * s) & TC_PEER_MASK, c0,
                                         tile * 128, (g.hint && kb < g.kb0 && g.kb_total > g.kb0) ? TC_L2_EVICT_LAST : pol);
                    if (++s == nstage) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===== MMA issuer: the leader's, for both SMs
        constexpr uint32_t idesc = (EPI == TC_FWD) ? make_idesc(0, 0, 0, 0, 256, 256) : make_idesc(1, 1, 0, 0, 256, 256);
        constexpr uint32_t idesc_d = make_idesc(0, 0, 0, 0, 256, 64);     // DGRAD2: fp16 H x fp16 diagonal block, N = 64
        mbar_wait_spin(bar_bfull, 0, 32);
        int s = 0, as = 0;
        uint32_t ph = 0, aph = 0;
        for (int it = 0; it < tp.count; ++it) {
            mbar_wait_cluster(bar_tempty + 8 * as, aph ^ 1, 33);
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)as * 256;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait_spin(bar_full + 8 * s, ph, 34);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a0 = smem_u32(sA + (size_t)s * TC_A_BYTES);
                    if (D2 && kb >= KBW) {
                        // D[:, 64 j .. 64 j + 63] += H[:, 64 j ..] * diag_j   (j = kb - KBW): the -c2 (.) H term of the BN backward
                        const int j = kb - KBW;
                        const uint32_t b0 = smem_u32(sDg + (size_t)j * 4096);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_2sm(dcol + (uint32_t)(64 * j), make_desc(a0 + k * 32, 16, 1024), make_desc(b0 + k * 32, 16, 1024),
                                         idesc_d, 1u);
                    } else {
                        const uint32_t b0 = smem_u32(sB + (size_t)kb * TC_B_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_2sm(dcol, make_desc(a0 + k * 32, 16, 1024), make_desc(b0 + k * 32, 16, 1024), idesc,
                                         (uint32_t)((kb | k) != 0));
                    }
                    umma_commit_2sm(bar_empty + 8 * s);
                    if (kb == KB - 1) umma_commit_2sm(bar_tfull + 8 * as);
                }
                __syncwarp();
                if (++s == nstage) { s = 0; ph ^= 1; }
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    } else if (warp >= 2) {
        // ===== epilogue: warp (2..9) -> TMEM lane quadrant warp%4 (rows q*32 + lane of THIS CTA's tile), 128-column half
        // (warp-2)/4, in two rounds of 64 columns; per round the code of k_tc_rowgemm's epilogue.
        const int q = warp & 3, half = (warp - 2) >> 2;   // TMEM lane quadrant (= warp index mod 4), column group of 64 * ROUNDS
        // one pair of staging buffers PER ROUND: round 1 does not wait for the bulk stores of round 0 to have read theirs
        // (with two buffers per warp every tile exposed the pick-up latency of a TMA store once; the pair form was
        // measured 12-19 % slower than the column-split kernel that way)
        // two 2 KB staging buffers per warp, used alternately by its 32-column chunks
        const uint32_t sbuf0 = smem_u32(sStage) + (uint32_t)(warp - 2) * (2 * TC_STAGE_BYTES), sbuf1 = sbuf0 + TC_STAGE_BYTES;
        float acc0[4 * ROUNDS], acc1[4 * ROUNDS];
#pragma unroll
        for (int i = 0; i < 4 * ROUNDS; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
        int as = 0;
        uint32_t aph = 0;
        static_assert(EPI == TC_FWD || EPI == TC_DGRAD2, "the pair form has a forward and a DGRAD2 epilogue");
        for (int it = 0; it < tp.count; ++it) {
            const int tile = (tp.first + it * tp.step) * 2 + rank;
            mbar_wait_spin(bar_tfull + 8 * as, aph, 35);
            tc_fence_after();
            const int row0 = tile * 128 + q * 32;
            const bool valid = row0 + lane < g.rows;
            // ---- phase A: the accumulator (this warp's 32 rows x 64 ROUNDS columns) -> packed 16-bit registers, in 32-column
            //      chunks with the TMEM load of chunk i + 1 in flight while chunk i is converted (as in k_tc_fused_eval); the
            //      TMEM stage goes back to the MMA warp as soon as the last load has landed.
            uint32_t ra[32], rb[32];
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256 + half * (64 * ROUNDS));
            // one 32-column chunk: accumulator registers -> 16-bit row -> staging buffer -> bulk store + column statistics.
            // Buffer `sb` was last used two chunks ago: at most the previous chunk's store group may still be pending.
            auto chunk = [&](const uint32_t (&rr)[32], const int ch, const uint32_t sb) {
                const int col0 = half * (64 * ROUNDS) + ch * 32;   // first output column of this chunk
                uint32_t po[16];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float v0, v1, v2, v3;
                    if (EPI == TC_FWD) {
                        const float4 b4 = *reinterpret_cast<const float4*>(&cvec[col0 + 4 * j4]);
                        v0 = __uint_as_float(rr[4 * j4 + 0]) + b4.x;
                        v1 = __uint_as_float(rr[4 * j4 + 1]) + b4.y;
                        v2 = __uint_as_float(rr[4 * j4 + 2]) + b4.z;
                        v3 = __uint_as_float(rr[4 * j4 + 3]) + b4.w;
                    } else {
                        // out = D / t + k  (a and t are folded into the weight operand, -c2 t (.) H came from the tensor core)
                        const float4 rt = *reinterpret_cast<const float4*>(&cvec[col0 + 4 * j4]);
                        const float4 kk = *reinterpret_cast<const float4*>(&cvec[256 + col0 + 4 * j4]);
                        v0 = fmaf(__uint_as_float(rr[4 * j4 + 0]), rt.x, kk.x);
                        v1 = fmaf(__uint_as_float(rr[4 * j4 + 1]), rt.y, kk.y);
                        v2 = fmaf(__uint_as_float(rr[4 * j4 + 2]), rt.z, kk.z);
                        v3 = fmaf(__uint_as_float(rr[4 * j4 + 3]), rt.w, kk.w);
                    }
                    if (!valid) { v0 = 0.f; v1 = 0.f; v2 = 0.f; v3 = 0.f; }
                    if (EPI == TC_FWD) {
                        const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
                        po[2 * j4] = *reinterpret_cast<const uint32_t*>(&h01);
                        po[2 * j4 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    } else {
                        const __nv_bfloat162 b01 = __floats2bfloat162_rn(v0, v1), b23 = __floats2bfloat162_rn(v2, v3);
                        po[2 * j4] = *reinterpret_cast<const uint32_t*>(&b01);
                        po[2 * j4 + 1] = *reinterpret_cast<const uint32_t*>(&b23);
                    }
                }
                if (lane == 0) tma_store_wait_read<1>();      // the group that last read `sb` (two chunks ago) is done
                __syncwarp();
                stage_put_row(sb, po, lane);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && !(g.debug & 2)) {
                    tma_store_2d_nocommit(&tmO, sb, col0, row0);
                    tma_store_commit();
                }
                if (!(g.debug & 1) && !g.nostat) {
                    if (EPI == TC_FWD) stage_col_sums<true, true>(sb, lane, acc0 + 2 * ch, acc1 + 2 * ch);
                    else stage_col_sums<false, false>(sb, lane, acc0 + 2 * ch, acc1 + 2 * ch);
                }
            };
            auto release_tmem = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader_cta(bar_tempty + 8 * as);
            };
            // the TMEM load of chunk i + 1 is in flight while chunk i is processed (as in k_tc_fused_eval); the TMEM stage goes
            // back to the MMA warp as soon as the last load has landed
            tmem_ld32_issue(tbase, ra);
            tmem_ld_wait();
            tmem_ld32_issue(tbase + 32, rb);
            chunk(ra, 0, sbuf0);
            tmem_ld_wait();
            if (ROUNDS == 2) {
                tmem_ld32_issue(tbase + 64, ra);
                chunk(rb, 1, sbuf1);
                tmem_ld_wait();
                tmem_ld32_issue(tbase + 96, rb);
                chunk(ra, 2, sbuf0);
                tmem_ld_wait();
                release_tmem();
                chunk(rb, 3, sbuf1);
            } else {
                release_tmem();
                chunk(rb, 1, sbuf1);
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
        if (lane == 0) tma_store_wait_all();
        // ---- column statistics: quadrant warps -> CTA partial (staging memory is free now) -> global per-CTA slot of 256
        //      columns; the last CTA to arrive adds the slots up.
        double* sred = reinterpret_cast<double*>(sStage);              // [2 stats][4 quadrants][256 columns] = 16 KB
        auto epi_sync = [] { asm volatile("bar.sync 1, %0;" ::"r"(32 * NEW) : "memory"); };    // the epilogue warps only
        epi_sync();
        if (lane < 16) {
#pragma unroll
            for (int cc = 0; cc < ROUNDS; ++cc)
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int cl = half * (64 * ROUNDS) + cc * 64 + c * 32 + 2 * lane + j;
                        sred[(0 * 4 + q) * 256 + cl] = (double)acc0[4 * cc + 2 * c + j];
                        sred[(1 * 4 + q) * 256 + cl] = (double)acc1[4 * cc + 2 * c + j];
                    }
        }
        epi_sync();
        const int t = threadIdx.x - 64;                                 // 0 .. 32 NEW - 1; the first 256 = one column each
        if (t < 256) {
#pragma unroll
            for (int st = 0; st < 2; ++st) {
                const double v = sred[(st * 4 + 0) * 256 + t] + sred[(st * 4 + 1) * 256 + t] + sred[(st * 4 + 2) * 256 + t] +
                                 sred[(st * 4 + 3) * 256 + t];
                g.partials[((size_t)blockIdx.x * 2 + st) * 256 + t] = v;
            }
            __threadfence();
        }
        epi_sync();
        if (t == 0) s_last = atomicAdd(g.counter, 1u) == gridDim.x - 1 ? 1u : 0u;
        epi_sync();
        if (s_last && t < 256) {
            __threadfence();
            double a0 = 0.0, a1 = 0.0;
            for (int b = 0; b < (int)gridDim.x; ++b) {
                a0 += g.partials[((size_t)b * 2 + 0) * 256 + t];
                if (EPI == TC_FWD) a1 += g.partials[((size_t)b * 2 + 1) * 256 + t];
            }
            g.stat0[t] = a0;
            if (EPI == TC_FWD) g.stat1[t] = a1;
            if (t == 0) *g.counter = 0;                                 // ready for the next launch on this stream
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                               // the peer may still be reading its accumulators
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// k_tc_wgrad on CTA PAIRS (cta_group::2), N = 256: out[256,256] += DH^T X over the rows of the chunk (split-K).
// The pair owns ONE 256 x 256 accumulator between its two SMs (M = 256: CTA `rank` holds the DH columns / output rows
// 128 rank .. +127 in its TMEM) and walks the 64-row k-blocks of its share of the rows together: per k-block each CTA loads
// only ITS 128 columns of DH and ITS 128 columns of X (32 KB instead of 64 KB: six ring stages), the leader issues one
// M = 256, N = 256 MMA per 16 rows for both SMs.  What this buys: the split-K reduction -- fp32 vector atomics of every
// CTA's whole accumulator at kernel exit, bound by the L2 atomic units (3 of the class's 16.7 ms per step) -- halves: 74
// pairs x 256 KB instead of 148 CTAs x 256 KB per launch.  X is fp16 (the forward's H_{l-1}): each CTA's eight otherwise
// idle epilogue warps rewrite its B half as bf16 in shared memory before the MMA may touch it (as in k_tc_wgrad); the
// converters of BOTH CTAs report to the leader's barrier.
// ---------------------------------------------------------------------------------------------------------------
#define TC_WG2_STAGE (32 * 1024)   // A: 2 boxes of 64 DH columns (16 KB) + B: 2 boxes of 64 X columns (16 KB)
#define TC_WG2_NST 6

__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_wgrad2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NST = TC_WG2_NST;
    uint64_t* bars = (uint64_t*)(smem + (size_t)NST * TC_WG2_STAGE);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 8), bar_conv = smem_u32(bars + 16);
    const uint32_t bar_tfull = smem_u32(bars + 24);
    uint32_t* tmem_slot = (uint32_t*)(bars + 25);
    const int rank = (int)cluster_ctarank();
    const int nkb = (g.rows + 63) >> 6;
    const int pair = (int)blockIdx.x >> 1;
    const int kb_beg = pair * g.kb_per_cta;                       // (kb_per_cta = k-blocks per PAIR here)
    const int kb_end = min(nkb, kb_beg + g.kb_per_cta);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 16);                      // eight converter warps of each CTA
        }
        mbar_init(bar_tfull, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), 256);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (kb_beg < kb_end) {
        if (warp == 0) {
            if (lane == 0) {
                int s = 0;
                uint32_t ph = 0;
                for (int kb = kb_beg; kb < kb_end; ++kb) {
                    mbar_wait_spin(bar_empty + 8 * s, ph ^ 1, 41);
                    uint8_t* st = smem + (size_t)s * TC_WG2_STAGE;
                    // (each CTA's loads complete on ITS OWN full barrier: an mbarrier can only be waited on locally, and the
                    // converter warps of both CTAs wait for their own tiles)
                    mbar_expect_tx(bar_full + 8 * s, 4u * 8192u);
                    for (int j = 0; j < 2; ++j) tma_load_2d(smem_u32(st + j * 8192), &tmA, bar_full + 8 * s, (2 * rank + j) * 64, kb * 64);
                    for (int j = 0; j < 2; ++j)
                        tma_load_2d(smem_u32(st + 16384 + j * 8192), &tmB, bar_full + 8 * s, (2 * rank + j) * 64, kb * 64);
                    if (++s == NST) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (rank == 0) {
                constexpr uint32_t idesc = make_idesc(1, 1, 1, 1, 256, 256);
                int s = 0;
                uint32_t ph = 0;
                for (int kb = kb_beg; kb < kb_end; ++kb) {
                    // the tiles of both CTAs have landed and (fp16 X) both B halves are rewritten as bf16: sixteen arrivals
                    mbar_wait_cluster(bar_conv + 8 * s, ph, 42);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a0 = smem_u32(smem + (size_t)s * TC_WG2_STAGE), b0 = a0 + 16384;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_2sm(tmem_base, make_desc(a0 + k * 2048, 8192, 1024), make_desc(b0 + k * 2048, 8192, 1024), idesc,
                                         (uint32_t)((kb != kb_beg) | (k != 0)));
                        umma_commit_2sm(bar_empty + 8 * s);
                        if (kb == kb_end - 1) umma_commit_2sm(bar_tfull);
                    }
                    __syncwarp();
                    if (++s == NST) { s = 0; ph ^= 1; }
                }
            }
        } else {
            const int q = warp & 3, half = (warp - 2) >> 2;
            {
                // this CTA's B half (64 rows x 128 columns fp16 = 1024 chunks of 16 B) -> bf16 in place (bf16 X: nothing to
                // rewrite, the warps only report that this CTA's tiles have landed)
                const int t256 = threadIdx.x - 64;
                int s = 0;
                uint32_t ph = 0;
                for (int kb = kb_beg; kb < kb_end; ++kb) {
                    mbar_wait(bar_full + 8 * s, ph, 44);
                    const uint32_t b0 = smem_u32(smem + (size_t)s * TC_WG2_STAGE) + 16384;
                    for (int i = t256; g.convert_b && i < 1024; i += 256) {
                        uint4 v = lds128(b0 + (uint32_t)i * 16);
                        uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[t]));
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(f.x, f.y);
                            w[t] = *reinterpret_cast<const uint32_t*>(&b2);
                        }
                        sts128(b0 + (uint32_t)i * 16, v);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    // (CTA-scope release: the rewritten tile is read by THIS SM's tensor core; the arrive only tells the leader
                    // that it may issue -- fence.proxy.async above has made the writes visible to the async proxy)
                    if (lane == 0) mbar_arrive_leader_cta(bar_conv + 8 * s);
                    if (++s == NST) { s = 0; ph ^= 1; }
                }
            }
            mbar_wait_spin(bar_tfull, 0, 43);
            tc_fence_after();
            // split-K reduction of this CTA's 128 output rows x 256 columns (see k_tc_wgrad: pair-coalesced vector atomics,
            // every CTA starts at a rotated block)
            const int nblk = (g.debug & 4) ? 0 : 4;
            const bool odd = lane & 1;
            for (int t0 = 0; t0 < nblk; ++t0) {
                const int c = half + 2 * ((t0 + (int)blockIdx.x) % 4);      // 32-column block 0..7
                const int m_even = rank * 128 + q * 32 + (lane & ~1), m_odd = m_even + 1;
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
                float* base = g.out + g.col_off + c * 32;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float keep[4], recv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float lo = __uint_as_float(r[8 * t + i]), hi = __uint_as_float(r[8 * t + 4 + i]);
                        keep[i] = odd ? hi : lo;
                        recv[i] = __shfl_xor_sync(FULL_MASK, odd ? lo : hi, 1);
                    }
                    float* d0 = base + (size_t)m_even * g.ldo + (2 * t + (odd ? 1 : 0)) * 4;
                    float* d1 = base + (size_t)m_odd * g.ldo + (2 * t + (odd ? 1 : 0)) * 4;
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d0), "f"(odd ? recv[0] : keep[0]),
                                 "f"(odd ? recv[1] : keep[1]), "f"(odd ? recv[2] : keep[2]), "f"(odd ? recv[3] : keep[3])
                                 : "memory");
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d1), "f"(odd ? keep[0] : recv[0]),
                                 "f"(odd ? keep[1] : recv[1]), "f"(odd ? keep[2] : recv[2]), "f"(odd ? keep[3] : recv[3])
                                 : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 256);
}

#define FZ_STAGES 6
#define FZ_BLK (128 * 128)       // 16 KB: 128 rows x 64 fp16 (one k-block of A, or one weight stage)

struct FusedMaps { CUtensorMap w[8]; };
struct FusedArgs {
    int cluster_arrive;         // 1: epilogue -> leader hand-off with mbarrier.arrive.release.cluster (round-1 form)
    int rows;
    float* out_p;               // [rows] sigmoid(logit)
};
// fp32 epilogue constants of the model whose pass is running (uploaded stream-ordered before its first chunk):
// bias[8][256] = b_0 and the folded biases b_l + W_l s_{l-1};  folded output layer w[256], b
__constant__ float c_fz[8 * 256 + 256 + 4];

__device__ __forceinline__ int fz_kblocks(int l) { return l == 0 ? 1 : (l == 4 ? 5 : 4); }

// NCTA = 1: every CTA on its own (two 128-row tiles per unit of work, N = 128 MMAs, 2 ring stages per k-block).
// NCTA = 2: clusters of two CTAs (four tiles per unit); one M = 256, N = 256 MMA covers a tile of each CTA and all 256
//           features with each CTA streaming HALF of every weight k-block (1 ring stage per k-block): shared-memory
//           operand traffic per SM and per MMA cycle halves (128 -> 64 B/clk), which is what bounds the single-CTA form.
template <int NCTA>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_fused_eval(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ FusedMaps wm, const FusedArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * FZ_STAGES + 2 + 2];
    __shared__ uint32_t tmem_slot;
    __shared__ float part[2][128];                        // [tile slot][row]: partial logit of column half 1
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sAct = smem;                                  // [2] x 64 KB (4 k-blocks)
    uint8_t* sW = sAct + 2 * 4 * FZ_BLK;                   // FZ_STAGES x 16 KB
    const uint32_t bar_wf = smem_u32(bars), bar_we = smem_u32(bars + FZ_STAGES);
    const uint32_t bar_accf = smem_u32(bars + 2 * FZ_STAGES), bar_epi = smem_u32(bars + 2 * FZ_STAGES + 2);

    const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;          // 0 = the pair's leader
    if (threadIdx.x == 0) {
        for (int j = 0; j < 2; ++j) {
            mbar_init(bar_accf + 8 * j, 1);
            mbar_init(bar_epi + 8 * j, 8 * NCTA);              // the epilogue warps of BOTH CTAs report to the leader
        }
        for (int s = 0; s < FZ_STAGES; ++s) { mbar_init(bar_wf + 8 * s, 1); mbar_init(bar_we + 8 * s, 1); }
        fence_barrier_init();
        tma_prefetch_desc(&tmE);
        for (int l = 0; l < 8; ++l) tma_prefetch_desc(&wm.w[l]);
    }
    if (warp == 1) {
        if (NCTA == 2) tmem_alloc_2sm(smem_u32(&tmem_slot), 512);
        else tmem_alloc(smem_u32(&tmem_slot), 512);
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA == 2) cluster_sync_all();                                // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // unit of work: 2 * NCTA row tiles; CTA `rank` of the unit owns tiles unit * 2 NCTA + 2 rank + {0, 1}
    const int ntiles = (g.rows + 127) >> 7, npairs = (ntiles + 2 * NCTA - 1) / (2 * NCTA);
    const int unit0 = (int)blockIdx.x / NCTA, ustep = (int)gridDim.x / NCTA;
    auto tile_of = [&](int unit, int j) { return unit * 2 * NCTA + 2 * rank + j; };

    if (warp == 0) {
        // ===== producer (a dedicated warp: it polls, no suspend)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            auto put = [&](const CUtensorMap* m, int c0, int c1) {
                mbar_wait_spin(bar_we + 8 * s, ph ^ 1, 22);
                if (NCTA == 2) {
                    // both CTAs fill the same stage of their own ring; the bytes of both count on the leader's barrier
                    if (rank == 0) mbar_expect_tx(bar_wf + 8 * s, 2 * FZ_BLK);
                    tma_load_2d_2sm(smem_u32(sW + s * FZ_BLK), m, (bar_wf + 8 * s) & TC_PEER_MASK, c0, c1);
                } else {
                    mbar_expect_tx(bar_wf + 8 * s, FZ_BLK);
                    tma_load_2d(smem_u32(sW + s * FZ_BLK), m, bar_wf + 8 * s, c0, c1);
                }
                if (++s == FZ_STAGES) { s = 0; ph ^= 1; }
            };
            for (int pair = unit0; pair < npairs; pair += ustep)
                for (int l = 0; l < 8; ++l) {
                    const int KB = fz_kblocks(l);
                    for (int j = 0; j < 2; ++j) {
                        if (l == 0 || l == 4) put(&tmE, 0, tile_of(pair, j) * 128);      // rows past the end read as zero
                        for (int kb = 0; kb < KB; ++kb) {
                            if (NCTA == 2) put(&wm.w[l], kb * 64, rank * 128);           // this CTA's half of the 256 weight rows
                            else for (int h = 0; h < 2; ++h) put(&wm.w[l], kb * 64, h * 128);
                        }
                    }
                }
        }
    } else if (warp == 1 && rank == 0) {
        // ===== MMA issuer (NCTA = 2: the leader's, for both SMs)
        constexpr uint32_t idesc = NCTA == 2 ? make_idesc(0, 0, 0, 0, 256, 256) : make_idesc(0, 0, 0, 0, 128, 128);
        int s = 0;
        uint32_t ph = 0, ed0 = 0, ed1 = 0;
        for (int pair = unit0; pair < npairs; pair += ustep) {
            for (int l = 0; l < 8; ++l) {
                const int KB = fz_kblocks(l);
                for (int j = 0; j < 2; ++j) {
                    // accumulator j drained and (l > 0) the fp16 activations of layer l-1 written by the epilogue
                    uint32_t& ed = j ? ed1 : ed0;
                    if (NCTA == 2) mbar_wait_cluster(bar_epi + 8 * j, ed ^ 1, 24);
                    else mbar_wait_spin(bar_epi + 8 * j, ed ^ 1, 24);
                    ed ^= 1;
                    tc_fence_after();
                    const uint32_t act_a = smem_u32(sAct + j * 4 * FZ_BLK);
                    uint32_t enc_a = 0, enc_bar = 0;
                    if (l == 0 || l == 4) {                    // this tile's encoding block arrives through the ring
                        mbar_wait_spin(bar_wf + 8 * s, ph, 23);
                        enc_a = smem_u32(sW + s * FZ_BLK);
                        enc_bar = bar_we + 8 * s;
                        if (++s == FZ_STAGES) { s = 0; ph ^= 1; }
                    }
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint32_t a0 = l == 0 ? enc_a : (l == 4 ? (kb == 0 ? enc_a : act_a + (kb - 1) * FZ_BLK) : act_a + kb * FZ_BLK);
                        for (int h = 0; h < (NCTA == 2 ? 1 : 2); ++h) {
                            mbar_wait_spin(bar_wf + 8 * s, ph, 25);
                            tc_fence_after();
                            if (lane == 0) {
                                const uint32_t b0 = smem_u32(sW + s * FZ_BLK);
                                const uint32_t d = tmem_base + (uint32_t)(j * 256 + h * 128);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (NCTA == 2)
                                        umma_f16_2sm(d, make_desc(a0 + k * 32, 16, 1024), make_desc(b0 + k * 32, 16, 1024), idesc,
                                                     (uint32_t)((kb | k) != 0));
                                    else
                                        umma_f16(d, make_desc(a0 + k * 32, 16, 1024), make_desc(b0 + k * 32, 16, 1024), idesc,
                                                 (uint32_t)((kb | k) != 0));
                                }
                                if (NCTA == 2) {
                                    umma_commit_2sm(bar_we + 8 * s);
                                    if (enc_bar && kb == 0) umma_commit_2sm(enc_bar);
                                } else {
                                    umma_commit(bar_we + 8 * s);
                                    if (enc_bar && kb == 0 && h == 1) umma_commit(enc_bar);     // both halves have read the encoding
                                }
                            }
                            __syncwarp();
                            if (++s == FZ_STAGES) { s = 0; ph ^= 1; }
                        }
                    }
                    if (lane == 0) {
                        if (NCTA == 2) umma_commit_2sm(bar_accf + 8 * j);
                        else umma_commit(bar_accf + 8 * j);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 2) {
        // ===== epilogue: warp -> TMEM lane quadrant q (rows q*32 + lane of the tile), column half (128 of the 256 features).
        // The two 64-column TMEM loads of a step are software-pipelined (the second is in flight while the first is
        // converted), biases / output weights come from constant memory (warp-uniform addresses: no load latency on the
        // critical path; ncu on the first version: 28 % of all stall samples behind tcgen05.wait::ld, long-scoreboard
        // stalls on the bias loads, tensor pipe 40 % busy because the epilogue, not the MMA, set the pace).
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        uint32_t af0 = 0, af1 = 0;
        auto step = [&](auto last_tag, int l, int j, int pair) {
            constexpr bool LAST = decltype(last_tag)::value;
            uint32_t& af = j ? af1 : af0;
            mbar_wait_spin(bar_accf + 8 * j, af, 26);
            af ^= 1;
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 256 + half * 128);
            const float* cb = c_fz + l * 256 + half * 128;
            const float* cw = c_fz + 8 * 256 + half * 128;
            const uint32_t act_a = smem_u32(sAct + j * 4 * FZ_BLK);
            float dot = 0.f;
            uint32_t r[2][32];
            auto process = [&](uint32_t (&rr)[32], int ch) {
                const int n0 = ch * 32;                              // first of this thread's 32 columns within the half
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const __half2 h2 = __floats2half2_rn(__uint_as_float(rr[2 * i]) + cb[n0 + 2 * i],
                                                         __uint_as_float(rr[2 * i + 1]) + cb[n0 + 2 * i + 1]);
                    pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                    if (LAST) {
                        // the output layer sees the same fp16-rounded H_7 the layered path stores
                        const float2 f = __half22float2(h2);
                        dot = fmaf(f.x, cw[n0 + 2 * i], dot);
                        dot = fmaf(f.y, cw[n0 + 2 * i + 1], dot);
                    }
                }
                if (!LAST) {
                    // columns half*128 + n0 .. +31 = half a k-block (half*2 + ch/2) of the next layer's A: 16-byte chunks
                    // (ch&1)*4 .. +3 of this row, at their SWIZZLE_128B positions
                    const uint32_t rbase = act_a + (uint32_t)((half * 2 + (ch >> 1)) * FZ_BLK + row * 128);
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        sts128(rbase + (uint32_t)(((((ch & 1) * 4 + m) ^ (row & 7))) << 4),
                               make_uint4(pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]));
                }
            };
            // 32-column TMEM loads, one ahead of the conversion
            tmem_ld32_issue(tbase, r[0]);
            tmem_ld_wait();
            tmem_ld32_issue(tbase + 32, r[1]);
            process(r[0], 0);
            tmem_ld_wait();
            tmem_ld32_issue(tbase + 64, r[0]);
            process(r[1], 1);
            tmem_ld_wait();
            tmem_ld32_issue(tbase + 96, r[1]);
            process(r[0], 2);
            tmem_ld_wait();
            process(r[1], 3);
            if (!LAST) fence_proxy_async();                   // the tensor core reads these stores through the async proxy
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                // (CTA-scope release, like the hand-offs of k_tc_rowgemm2 / k_tc_wgrad2: the activations written above are read
                // by THIS SM's tensor core, fence.proxy.async has published them; PCNERF_TC_EVAL_CLUSTER_ARRIVE=1 restores
                // the cluster-scope release of round 1 for A/B runs)
                if (NCTA == 2) { if (g.cluster_arrive) mbar_arrive_leader(bar_epi + 8 * j); else mbar_arrive_leader_cta(bar_epi + 8 * j); }
                else mbar_arrive(bar_epi + 8 * j);
            }
            if (LAST) {
                if (half == 1) part[j][row] = dot;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (half == 0) {
                    const int64_t grow = (int64_t)tile_of(pair, j) * 128 + row;
                    if (grow < g.rows) {
                        const float t = dot + part[j][row] + c_fz[8 * 256 + 256];
                        g.out_p[grow] = 1.f / (1.f + expf(-t));
                    }
                }
            }
        };
        for (int pair = unit0; pair < npairs; pair += ustep) {
            for (int l = 0; l < 7; ++l) {
                step(std::false_type{}, l, 0, pair);
                step(std::false_type{}, l, 1, pair);
            }
            step(std::true_type{}, 7, 0, pair);
            step(std::true_type{}, 7, 1, pair);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA == 2) {
        cluster_sync_all();                                           // the peer may still be reading its accumulators
        if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
    } else if (warp == 1) {
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 16-bit weight copies
// ---------------------------------------------------------------------------------------------------------------
// Wh0[256][64] = fp16(Wp0);  split != 0: Wh0[256][128] = [hi | lo], hi = fp16(Wp0), lo = fp16(Wp0 - hi) (the encoding
// tile is multiplied by both blocks: the layer-0 weights enter the GEMM with ~22 significant bits)
__global__ void k_tc_prep_fwd(const float* __restrict__ Wp0, __half* __restrict__ Wh0, int split) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 256 * 64) return;
    const float w = Wp0[i];
    const __half hi = __float2half_rn(w);
    if (!split) { Wh0[i] = hi; return; }
    const int o = i >> 6, c = i & 63;
    Wh0[o * 128 + c] = hi;
    Wh0[o * 128 + 64 + c] = __float2half_rn(w - __half2float(hi));
}

// ---------------------------------------------------------------------------------------------------------------
// k_tc_fold: BN(l) batch statistics -> (mean, invstd, a, s), running statistics, and the fp16 operand of layer nl = l + 1
// (l = 0..6) TOGETHER WITH ITS LINEAR CORRECTION BLOCK.
//
// Rounding the folded weights W' = W_nl diag(a_l) to fp16 perturbs every weight by up to 2^-11 relative -- the largest
// single term of the tensor-core path's depth error (scripts/emulate_tc_precision.py: worst ray of a 16,384-ray C2 batch
// 1.06e-3 with it, 0.69e-3 without).  The perturbation is the same for every row of the chunk, and as the reference builds
// the network (identity activations, nof/networks/models.py:152,172) the input of layer nl is an affine function of the
// 64-d encoding x: H_l = T_l x + t_l.  So what the rounding loses, (W' - fp16(W')) H_l, is itself affine in x and is put
// back EXACTLY by one extra K = 64 block of the same GEMM:
//     H_nl = fp16(W') H_l  +  C x  +  bias,   C = (W' - fp16(W')) T_l   [256 x 64, fp16: second-order rounding only],
//     bias = b_nl + W_nl s_l + (W' - fp16(W')) t_l,
// with the chain T_nl = W' T_l (+ the skip block of layer 4), t_nl = W' t_l + b_nl + W_nl s_l carried in fp32 from
// T_0 = W_0, t_0 = b_0.  Layer 4 already reads x: its encoding weights enter as [hi | lo + C].  Cost: one more 64-wide
// k-block per layer (the kernel layer 4 always was), 2 x 16 K FMAs per block of this kernel.
// grid: 256 blocks (output feature o of layer nl) x 256 threads (hidden input feature i).
// Wh_next row layout (ldw): correct ? (nl == 4 ? [hi 64 | lo + C 64 | W' 256] : [C 64 | W' 256])
//                                   : (nl == 4 ? [W_enc 64 | W' 256] : [W' 256]).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tc_fold(int l, int training, int64_t rows, const double* __restrict__ sum,
                                                 const double* __restrict__ sumsq, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float* __restrict__ running_mean,
                                                 float* __restrict__ running_var, int64_t* __restrict__ nbt, float momentum,
                                                 float eps, float* __restrict__ stats /* [4][256] */,
                                                 const float* __restrict__ Wp_next, const float* __restrict__ b_next,
                                                 float* __restrict__ bf_next, __half* __restrict__ Wh_next, int ldw,
                                                 const float* __restrict__ T_prev, const float* __restrict__ t_prev,
                                                 float* __restrict__ T_next, float* __restrict__ t_next, int correct) {
    __shared__ float sh_w[256], sh_d[256];
    __shared__ float red[3][8];
    __shared__ float redc[2][4][64];
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
    const int kpad = mlp_kpad(nl), off = nl == 4 ? 64 : 0;
    const int encb = correct ? (nl == 4 ? 2 : 1) : (nl == 4 ? 1 : 0);          // encoding blocks in front of W'
    const float w = Wp_next[o * kpad + off + i];
    const float wa = w * a;
    const __half wr = __float2half_rn(wa);
    const float d = wa - __half2float(wr);
    Wh_next[(size_t)o * ldw + encb * 64 + i] = wr;
    sh_w[i] = wa;
    sh_d[i] = d;
    const float tp = correct ? t_prev[i] : 0.f;
    const float p0 = warp_sum(w * s), p1 = warp_sum(wa * tp), p2 = warp_sum(d * tp);
    if ((i & 31) == 0) { red[0][i >> 5] = p0; red[1][i >> 5] = p1; red[2][i >> 5] = p2; }
    __syncthreads();
    if (i == 0) {
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        for (int k = 0; k < 8; ++k) { r0 += red[0][k]; r1 += red[1][k]; r2 += red[2][k]; }
        const float bias = b_next[o] + r0;
        bf_next[o] = bias + r2;
        if (correct) t_next[o] = bias + r1;
    }
    if (!correct) {
        if (nl == 4 && i < 64) Wh_next[(size_t)o * ldw + i] = __float2half_rn(Wp_next[o * kpad + i]);
        return;
    }
    const int c = i & 63, part = i >> 6;
    float accT = 0.f, accC = 0.f;
    // this thread's 64 values of column c of T_prev: all loads in flight at once (L2 hits, ~1 latency instead of 8)
    float tv[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) tv[k] = __ldg(T_prev + (part * 64 + k) * 64 + c);
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        accT = fmaf(sh_w[part * 64 + k], tv[k], accT);
        accC = fmaf(sh_d[part * 64 + k], tv[k], accC);
    }
    redc[0][part][c] = accT;
    redc[1][part][c] = accC;
    __syncthreads();
    if (part != 0) return;
    accT = redc[0][0][c] + redc[0][1][c] + redc[0][2][c] + redc[0][3][c];
    accC = redc[1][0][c] + redc[1][1][c] + redc[1][2][c] + redc[1][3][c];
    if (nl == 4) {
        const float we = Wp_next[o * kpad + c];                  // skip block (column 63 is the zero pad)
        const __half hi = __float2half_rn(we);
        T_next[o * 64 + c] = we + accT;
        Wh_next[(size_t)o * ldw + c] = hi;
        Wh_next[(size_t)o * ldw + 64 + c] = __float2half_rn((we - __half2float(hi)) + accC);
    } else {
        T_next[o * 64 + c] = accT;
        Wh_next[(size_t)o * ldw + c] = __float2half_rn(accC);
    }
}
// BN(l-1) backward coefficients WITHOUT a pass over the data-gradient G = DH_l W_l:
//   sum_r G[r,n]             = sum_o colsum_l[o] W_l[o,n]                 (colsum_l = column sums of DH_l)
//   sum_r G[r,n] H_{l-1}[r,n] = sum_o W_l[o,n] (DH_l^T H_{l-1})[o,n]        (the raw weight gradient of layer l)
// so the data-gradient GEMM can apply the BN backward in its own epilogue.
//
// Per layer, after the weight-gradient GEMM: dW_l / db_l from the raw product (x a_{l-1}, + db (x) s_{l-1}) AND the BN(l-1)
// backward coefficients (see above) in one pass over `part`.  One block per padded input column, one thread per output row.
__global__ void __launch_bounds__(256) k_tc_wgrad_finish(int l, const float* __restrict__ part, const double* __restrict__ colsum,
                                                         const float* __restrict__ prev_stats, const float* __restrict__ Wp,
                                                         int64_t rows, float* __restrict__ dW, float* __restrict__ db,
                                                         float* __restrict__ dgamma_prev, float* __restrict__ dbeta_prev,
                                                         float* __restrict__ coef, __nv_bfloat16* __restrict__ B1,
                                                         __half* __restrict__ Dg, float* __restrict__ rtk) {
    __shared__ double r0[8], r1[8];
    __shared__ float s_c2t[2];                          // c2, t of this block's column (DGRAD2 operands)
    const int kin = mlp_kin(l), kpad = mlp_kpad(l);
    const int c = blockIdx.x, o = threadIdx.x;
    int real = c, hid = c;
    bool is_hidden = true, live = true;
    if (l == 0) { is_hidden = false; live = c < 63; }
    else if (l == 4) {
        if (c < 64) { is_hidden = false; live = c != 63; }
        else { real = c - 1; hid = c - 64; }
    }
    const float v = part[(size_t)o * kpad + c];
    const double cs = colsum[o];
    const float dbias = (float)cs;
    if (live) dW[o * kin + real] += is_hidden ? v * prev_stats[512 + hid] + dbias * prev_stats[768 + hid] : v;
    if (c == 0) db[o] += dbias;
    if (!is_hidden) return;                              // (whole block: uniform)
    const float w = Wp[(size_t)o * kpad + c];
    double st0 = warp_sum_d(cs * (double)w), st1 = warp_sum_d((double)w * (double)v);
    if ((o & 31) == 0) { r0[o >> 5] = st0; r1[o >> 5] = st1; }
    __syncthreads();
    const int n = hid;
    const float a = prev_stats[512 + n];
    if (o == 0) {
        st0 = 0.0;
        st1 = 0.0;
        for (int k = 0; k < 8; ++k) { st0 += r0[k]; st1 += r1[k]; }
        const float mean = prev_stats[n], invstd = prev_stats[256 + n];
        const float dbt = (float)st0;
        const float dg = invstd * (float)(st1 - (double)mean * st0);
        dgamma_prev[n] += dg;
        dbeta_prev[n] += dbt;
        const float B = (float)rows;
        const float c1 = a * dbt / B, c2 = a * invstd * dg / B;
        coef[n] = a;
        coef[256 + n] = c1;
        coef[512 + n] = c2;
        coef[768 + n] = mean;
        if (B1) {
            // operands of the pair-form data-gradient GEMM (see k_tc_dgrad2_prep): per-column power-of-two scale t
            float t = 1.f;
            if (c2 != 0.f && isfinite(c2)) {
                int e = 0;
                frexpf(fabsf(c2), &e);
                e = e < -100 ? -100 : (e > 100 ? 100 : e);
                t = ldexpf(1.f, -e);
            }
            s_c2t[0] = c2;
            s_c2t[1] = t;
            rtk[n] = 1.f / t;
            rtk[256 + n] = c2 * mean - c1;
        }
    }
    if (!B1) return;                                      // (uniform)
    __syncthreads();
    const float c2 = s_c2t[0], t = s_c2t[1];
    B1[(size_t)n * 256 + o] = __float2bfloat16_rn(w * a * t);
    if (o < 64) Dg[n * 64 + o] = (o == (n & 63)) ? __float2half_rn(-c2 * t) : __float2half_rn(0.f);
}

struct PrepTArgs {
    const float* Wp[8];
    __nv_bfloat16* WT[8];
};
// WT_l[in][out] = bf16(Wp_l[out][hidden col in]) for l = 1..7 (the B operand of the data-gradient GEMM)
__global__ void k_tc_prep_bwd(PrepTArgs a) {
    __shared__ float t[32][33];
    const int l = blockIdx.z + 1;
    const int kpad = mlp_kpad(l), off = l == 4 ? 64 : 0;
    const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = a.Wp[l][(size_t)(o0 + r) * kpad + off + i0 + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        a.WT[l][(size_t)(i0 + r) * 256 + o0 + threadIdx.x] = __float2bfloat16_rn(t[threadIdx.x][r]);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D map over a row-major 16-bit matrix [rows][ld] exposing `cols` columns; box = box_rows x box_cols columns with the
// swizzle that matches a box row (64 columns = 128 B -> SWIZZLE_128B, 32 columns = 64 B -> SWIZZLE_64B)
int make_map(CUtensorMap* m, const void* base, int64_t rows, int cols, int ld, int box_rows, int box_cols = 64) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { pcn_set_error("cuTensorMapEncodeTiled is not available from this driver"); return PCNERF_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { pcn_set_error("cuTensorMapEncodeTiled failed (%d) for [%lld][%d] ld %d", (int)r, (long long)rows, cols, ld); return PCNERF_ERR_CUDA; }
    return 0;
}

int tc_sched_mode() {
    static int m = -1;
    if (m < 0) { const char* e = getenv("PCNERF_TC_SCHED"); m = e ? atoi(e) : 2; if (m < 0 || m > 2) m = 2; }
    return m;
}

// row GEMMs on CTA pairs (k_tc_rowgemm2): pcnerf_tc_set_row_pairs(1) or PCNERF_TC_PAIRS=1; default off (see the kernel)
// bit 0: forward GEMMs, bit 1: data-gradient GEMMs, bit 2: weight-gradient GEMMs (N = 256)
int g_row_pairs = -1;
int tc_pairs_mode() {
    if (g_row_pairs < 0) { const char* e = getenv("PCNERF_TC_PAIRS"); g_row_pairs = e ? (atoi(e) & 7) : TC_PAIRS_DEFAULT; }
    return g_row_pairs;
}
bool tc_pairs_for(int mode) { return (tc_pairs_mode() >> (mode == TC_FWD ? 0 : 1)) & 1; }

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = PCN_SM_COUNT;
    }
    return n;
}

// mode TC_FWD: A fp16, B fp16 -> out fp16 (+bias), out2 bf16 copy;  TC_DGRAD: A bf16, B bf16 -> out bf16 (BN backward
// fused: vec = c0|c1|c2|mean, E = fp16 H)
// `work` = >= TC_ROWGEMM_WORK_BYTES of device memory: [0,4) CTA counter (zeroed here), then the per-CTA partials
#define TC_ROWGEMM_WORK_CORE (256 + 160 * 2 * 256 * 8)
#define TC_ROWGEMM_WORK_BYTES (TC_ROWGEMM_WORK_CORE + 256 * 256 * 2 + 256 * 64 * 2 + 512 * 4)     // + the DGRAD2 operands
// a0_rep: every 64-column block of A0 is consumed a0_rep times (k-blocks 0 .. a0_rep*k0/64 - 1 of B multiply A0)
int launch_rowgemm(int mode, const void* A0, int lda0, int k0, const void* A1, int lda1, int k1, const void* B, int ldb,
                   const float* vec, const __half* E, int64_t rows, void* out, __nv_bfloat16* out2, double* stat0,
                   double* stat1, void* work, int dir, cudaStream_t st, int nostat = 0, int a0_rep = 1, int alg_k = -1) {
    PCN_CHECK_ARG(k0 % 64 == 0 && k1 % 64 == 0 && k0 >= 64 && a0_rep >= 1 && (k0 * a0_rep + k1) <= 384,
                  "tc rowgemm: K must be 64..384 in 64s");
    CUtensorMap mA0, mA1, mB;
    int rc = make_map(&mA0, A0, rows, k0, lda0, 128);
    if (rc) return rc;
    rc = k1 ? make_map(&mA1, A1, rows, k1, lda1, 128) : make_map(&mA1, A0, rows, k0, lda0, 128);
    if (rc) return rc;
    rc = make_map(&mB, B, 256, k0 * a0_rep + k1, ldb, TC_NCTA);
    if (rc) return rc;
    CUtensorMap mO, mO2;
    rc = make_map(&mO, out, rows, 256, 256, 32, 32);
    if (rc) return rc;
    rc = make_map(&mO2, out2 ? (void*)out2 : out, rows, 256, 256, 32, 32);
    if (rc) return rc;
    RowGemmArgs g;
    g.rows = (int)rows; g.a0_blocks = k0 / 64; g.kb0 = a0_rep * k0 / 64; g.kb_total = g.kb0 + k1 / 64;
    const bool pairs = mode == TC_FWD && tc_pairs_for(mode);      // (the data gradient on pairs is launch_dgrad2)
    static int epi8 = -1;     // forward on pairs: 8 epilogue warps (four 32-column chunks each) or 16 (PCNERF_TC_EPI8=0)
    if (epi8 < 0) { const char* e = getenv("PCNERF_TC_EPI8"); epi8 = e ? (atoi(e) != 0) : 1; }
    const bool wide = pairs && mode == TC_FWD && !epi8;
    // 2 KB staging buffers per 8 epilogue warps' worth: the pair form alternates two per warp (its A ring gets the rest:
    // six stages at K = 320 -- measured 17.2 vs 18.0-18.9 ms per step against four stages with four buffers per warp)
    const int nbuf = pairs ? (wide ? 4 : 2) : TC_NBUF;
    {
        // everything that is left of the 227 KB after the resident weights and the staging buffers becomes A ring
        const size_t fixed = 1024 + (size_t)g.kb_total * TC_B_BYTES + 8 * nbuf * TC_STAGE_BYTES + 4096 /* static */;
        int ns = (int)((232448 - fixed) / TC_A_BYTES);
        g.nstage = ns > 8 ? 8 : ns;
    }
    g.out = out; g.out2 = out2; g.vec = vec; g.E = E; g.stat0 = stat0; g.stat1 = stat1;
    g.counter = (unsigned int*)work;            // zero on entry (the caller clears it once; the last CTA re-arms it)
    g.partials = (double*)((char*)work + 256);
    {
        // PCNERF_TC_SCHED: 0 = interleaved row tiles, 1 = serpentine slabs, 2 (default) = serpentine slabs + L2 hints
        const int m = tc_sched_mode();
        g.sched = m == 0 ? 0 : (dir ? 2 : 1);
        g.hint = m == 2 ? 1 : 0;
    }
    g.nostat = nostat;
    g.stage_pairs = 1;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("PCNERF_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
        g.debug = dbg;
    }
    const size_t smem = 1024 + (size_t)g.kb_total * TC_B_BYTES + (size_t)g.nstage * TC_A_BYTES + 8 * nbuf * TC_STAGE_BYTES;
    const int ntiles = (int)pcn_cdiv(rows, 128);
    const int grid = 2 * ntiles < sm_count() ? 2 * ntiles : (sm_count() & ~1);
    // (algorithmic FLOPs: the repeated encoding blocks carry the split / correction weights -- extra tensor work, not
    // extra algorithmic work)
    const double flops = 2.0 * (double)rows * 256.0 * (double)(alg_k >= 0 ? alg_k : k0 + k1);
    if (pairs) {
        // CTA pairs (k_tc_rowgemm2): clusters of two, a unit of work = two row tiles
        const int nunits = (ntiles + 1) / 2;
        const int ncl = nunits < sm_count() / 2 ? nunits : sm_count() / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * ncl);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (mode == TC_FWD) {
            // forward: 8 epilogue warps by default; 16 (one 64-column round each) measured no faster (18.4 vs 17.9 ms per step)
            PcnScope ps(PCN_K_GEMM_FWD, st, flops);
            if (epi8) {
                PCN_CUDA(cudaFuncSetAttribute(k_tc_rowgemm2<TC_FWD, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PCN_CUDA(cudaLaunchKernelEx(&cfg, k_tc_rowgemm2<TC_FWD, 8>, mA0, mA1, mB, mO, mO, g));
            } else {
                cfg.blockDim = dim3(64 + 32 * 16);
                PCN_CUDA(cudaFuncSetAttribute(k_tc_rowgemm2<TC_FWD, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PCN_CUDA(cudaLaunchKernelEx(&cfg, k_tc_rowgemm2<TC_FWD, 16>, mA0, mA1, mB, mO, mO, g));
            }
        }
        PCN_LAUNCH_CHECK();
        return 0;
    }
    if (mode == TC_FWD) {
        PCN_CUDA(cudaFuncSetAttribute(k_tc_rowgemm<TC_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PcnScope ps(PCN_K_GEMM_FWD, st, flops);
        k_tc_rowgemm<TC_FWD><<<grid, TC_THREADS, smem, st>>>(mA0, mA1, mB, mO, mO2, g);
    } else {
        PCN_CUDA(cudaFuncSetAttribute(k_tc_rowgemm<TC_DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PcnScope ps(PCN_K_GEMM_DGRAD, st, flops);
        k_tc_rowgemm<TC_DGRAD><<<grid, TC_THREADS, smem, st>>>(mA0, mA1, mB, mO, mO2, g);
    }
    PCN_LAUNCH_CHECK();
    return 0;
}

int launch_wgrad(const void* DH, const void* X, int ldx, int N, int x_is_bf16, int64_t rows, float* out, int ldo,
                 int col_off, cudaStream_t st) {
    PCN_CHECK_ARG(N == 256 || N == 64, "tc wgrad: N must be 64 or 256");
    CUtensorMap mA, mB;
    int rc = make_map(&mA, DH, rows, 256, 256, 64);
    if (rc) return rc;
    rc = make_map(&mB, X, rows, N, ldx, 64);
    if (rc) return rc;
    WgradArgs g;
    g.rows = (int)rows; g.N = N; g.out = out; g.ldo = ldo; g.col_off = col_off; g.convert_b = x_is_bf16 ? 0 : 1;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("PCNERF_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
        g.debug = dbg;
    }
    const int nkb = (int)pcn_cdiv(rows, 64);
    if (N == 256 && ((tc_pairs_mode() >> 2) & 1)) {
        // CTA pairs (k_tc_wgrad2): one 256 x 256 accumulator per pair, half the split-K atomics
        const int npairs = sm_count() / 2;
        g.kb_per_cta = (int)pcn_cdiv(nkb, npairs);                    // k-blocks per PAIR
        const int ncl = (int)pcn_cdiv(nkb, g.kb_per_cta);
        const size_t smem2 = 1024 + (size_t)TC_WG2_NST * TC_WG2_STAGE + 32 * 8;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * ncl);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = smem2;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        PCN_CUDA(cudaFuncSetAttribute(k_tc_wgrad2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        PcnScope ps(PCN_K_GEMM_WGRAD, st, 2.0 * (double)rows * 256.0 * (double)N);
        PCN_CUDA(cudaLaunchKernelEx(&cfg, k_tc_wgrad2, mA, mB, g));
        PCN_LAUNCH_CHECK();
        return 0;
    }
    g.kb_per_cta = (int)pcn_cdiv(nkb, sm_count());
    const int grid = (int)pcn_cdiv(nkb, g.kb_per_cta);
    const size_t smem = 1024 + 3 * (size_t)TC_WG_STAGE + 16 * 8 + 64;      // 13 barriers + the TMEM slot
    PCN_CUDA(cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PcnScope ps(PCN_K_GEMM_WGRAD, st, 2.0 * (double)rows * 256.0 * (double)N);
    k_tc_wgrad<<<grid, TC_THREADS, smem, st>>>(mA, mB, g);
    PCN_LAUNCH_CHECK();
    return 0;
}

// Operands of the DGRAD2 form for one layer and chunk (grid: 256 blocks = hidden column n of layer l-1, 256 threads = output
// row o of layer l).  DH_{l-1}[r,n] = a_n G[r,n] - c2_n H[r,n] + (c2_n mean_n - c1_n), G = DH_l W_l.  With t_n the power of
// two that brings |c2_n| t_n into [0.5, 1):
//   B1[n][o] = bf16(W_l[o][n] a_n t_n)   (K-major in o: the B operand of the N = 256 MMAs)
//   Dg[n][k] = (k == n mod 64) ? fp16(-c2_n t_n) : 0   (four 64 x 64 diagonal blocks: the B operand of the N = 64 MMAs on H)
//   rtk = 1 / t_n | c2_n mean_n - c1_n   (epilogue: out = D / t + k)
// The per-column power-of-two scale keeps the fp16 diagonal exact to 11 bits whatever the gradient magnitude.
__global__ void __launch_bounds__(256) k_tc_dgrad2_prep(const float* __restrict__ Wp, int kpad, int off,
                                                        const __nv_bfloat16* __restrict__ WT, const float* __restrict__ coef,
                                                        __nv_bfloat16* __restrict__ B1, __half* __restrict__ Dg,
                                                        float* __restrict__ rtk) {
    const int n = blockIdx.x, o = threadIdx.x;
    const float a = coef[n], c1 = coef[256 + n], c2 = coef[512 + n], mean = coef[768 + n];
    float t = 1.f;
    if (c2 != 0.f && isfinite(c2)) {
        int e = 0;
        frexpf(fabsf(c2), &e);
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        t = ldexpf(1.f, -e);
    }
    const float w = Wp ? Wp[(size_t)o * kpad + off + n] : __bfloat162float(WT[(size_t)n * 256 + o]);
    B1[(size_t)n * 256 + o] = __float2bfloat16_rn(w * a * t);
    if (o < 64) Dg[n * 64 + o] = (o == (n & 63)) ? __float2half_rn(-c2 * t) : __float2half_rn(0.f);
    if (o == 0) { rtk[n] = 1.f / t; rtk[256 + n] = c2 * mean - c1; }
}

#define TC_DGRAD2_SCRATCH (256 * 256 * 2 + 256 * 64 * 2 + 512 * 4)     // B1 | Dg | rtk
// data-gradient GEMM + BN backward on CTA pairs, H_{l-1} term on the tensor core (k_tc_rowgemm2<TC_DGRAD2>)
// (Wp == NULL and WT == NULL: the operands in `opscratch` were already written, by k_tc_wgrad_finish)
int launch_dgrad2(const void* DH, const __half* Hprev, const float* Wp, int kpad, int off, const __nv_bfloat16* WT,
                  const float* coef, int64_t rows, void* out, double* colsum, void* work, char* opscratch, int dir,
                  cudaStream_t st) {
    __nv_bfloat16* B1 = (__nv_bfloat16*)opscratch;
    __half* Dg = (__half*)(opscratch + 256 * 256 * 2);
    float* rtk = (float*)(opscratch + 256 * 256 * 2 + 256 * 64 * 2);
    if (Wp || WT)
        PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_tc_dgrad2_prep<<<256, 256, 0, st>>>(Wp, kpad, off, WT, coef, B1, Dg, rtk));
    CUtensorMap mA0, mA1, mB, mO, mD;
    int rc = make_map(&mA0, DH, rows, 256, 256, 128);
    if (rc) return rc;
    rc = make_map(&mA1, Hprev, rows, 256, 256, 128);
    if (rc) return rc;
    rc = make_map(&mB, B1, 256, 256, 256, TC_NCTA);
    if (rc) return rc;
    rc = make_map(&mD, Dg, 256, 64, 64, 32);
    if (rc) return rc;
    rc = make_map(&mO, out, rows, 256, 256, 32, 32);
    if (rc) return rc;
    RowGemmArgs g = {};
    g.rows = (int)rows; g.a0_blocks = 4; g.kb0 = 4; g.kb_total = 8;
    g.stage_pairs = 1;
    const int nbuf = 2;
    {
        const size_t fixed = 1024 + (size_t)5 * TC_B_BYTES + 8 * nbuf * TC_STAGE_BYTES + 4096;
        int ns = (int)((232448 - fixed) / TC_A_BYTES);
        g.nstage = ns > 8 ? 8 : ns;
    }
    g.out = out; g.out2 = nullptr; g.vec = rtk; g.E = nullptr; g.stat0 = colsum; g.stat1 = nullptr;
    g.counter = (unsigned int*)work;
    g.partials = (double*)((char*)work + 256);
    {
        const int m = tc_sched_mode();
        g.sched = m == 0 ? 0 : (dir ? 2 : 1);
        g.hint = m == 2 ? 1 : 0;
    }
    g.nostat = 0;
    g.debug = 0;
    const size_t smem = 1024 + (size_t)5 * TC_B_BYTES + (size_t)g.nstage * TC_A_BYTES + 8 * nbuf * TC_STAGE_BYTES;
    const int ntiles = (int)pcn_cdiv(rows, 128), nunits = (ntiles + 1) / 2;
    const int ncl = nunits < sm_count() / 2 ? nunits : sm_count() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ncl);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PCN_CUDA(cudaFuncSetAttribute(k_tc_rowgemm2<TC_DGRAD2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PcnScope ps(PCN_K_GEMM_DGRAD, st, 2.0 * (double)rows * 256.0 * 256.0);
    PCN_CUDA(cudaLaunchKernelEx(&cfg, k_tc_rowgemm2<TC_DGRAD2, 8>, mA0, mA1, mB, mO, mD, g));
    PCN_LAUNCH_CHECK();
    return 0;
}

}  // namespace

static int g_fused_eval = 2;     // 0 = layered, 1 = fused (one CTA per unit), 2 = fused on CTA pairs (cta_group::2)
extern "C" void pcnerf_tc_set_fused_eval(int on) { g_fused_eval = on < 0 ? 0 : (on > 2 ? 2 : on); }
extern "C" int pcnerf_tc_get_fused_eval(void) { return g_fused_eval; }
extern "C" void pcnerf_tc_set_row_pairs(int on) { g_row_pairs = on & 7; }
extern "C" int pcnerf_tc_get_row_pairs(void) { return tc_pairs_mode(); }

// ---- two BN batches (chunks) in flight: pcnerf_mlp_tc_{forward,backward}_chunks issue consecutive chunks on two internal
// streams ("lanes").  The kernels of a chunk form a serial chain (GEMM -> fold -> GEMM ...) whose fixed costs -- prologue,
// pipeline fill, last-tile drain, statistics reduction, the small kernels and the launch gaps between them, ~10 us of a
// 62 us forward GEMM at 262,144 rows -- are dead time for the whole GPU; with a second, independent chain the CTAs of the
// other chunk's GEMM take over every SM the moment it is released.  What the chains share is ordered explicitly: the
// running-statistics update of BN(l) (forward, k_bn_fold) and the += into the parameter gradients (backward,
// k_out_bwd_finalize / k_tc_wgrad_finish) of chunk c+1 wait for the same kernel of chunk c, so every sum and every
// momentum update happens in the sequential order (results are bit-identical to one lane).
struct ChainSync {
    cudaEvent_t wait[9];         // recorded by the previous chunk (other lane) after its kernel of this slot; may be null
    cudaEvent_t rec[9];          // recorded by this chunk after its kernel of this slot
};
static thread_local const ChainSync* g_chain = nullptr;
static int chain_before(int slot, cudaStream_t st) {
    if (g_chain && g_chain->wait[slot]) PCN_CUDA(cudaStreamWaitEvent(st, g_chain->wait[slot], 0));
    return 0;
}
static int chain_after(int slot, cudaStream_t st) {
    if (g_chain && g_chain->rec[slot]) PCN_CUDA(cudaEventRecord(g_chain->rec[slot], st));
    return 0;
}

// fp32 padded weight copies (same kernel as the fp32 path; defined in mlp_small.cuh)
static void tc_prep_weights(const pcnerf_mlp_params* P, const MlpLayout& L, char* scratch, cudaStream_t st) {
    PrepArgs pa;
    for (int l = 0; l < 8; ++l) { pa.W[l] = P->W[l]; pa.Wp[l] = L.Wp(scratch, l); }
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_prep_weights<<<dim3(64, 8), 256, 0, st>>>(pa));
}

// layout of the 16-bit weight area (MlpLayout::off_tc): Wh[8] fp16 [256][<= 384] | WT[8] bf16 [256][256]
static __half* tc_Wh(const MlpLayout& L, char* scratch, int l) { return (__half*)L.tc(scratch) + (size_t)l * 256 * 384; }
static __nv_bfloat16* tc_WT(const MlpLayout& L, char* scratch, int l) {
    return (__nv_bfloat16*)(L.tc(scratch) + (size_t)8 * 256 * 384 * 2) + (size_t)l * 256 * 256;
}
// chain of the affine maps H_l = T_l x + t_l (k_tc_fold): two ping-pong [256][64] fp32 matrices and [256] vectors, in the
// (otherwise unused on this path) fp32 folded-weight slots of layers 1 and 2
static float* tc_T(const MlpLayout& L, char* scratch, int k) { return L.Wf(scratch, 1) + (size_t)k * 256 * 64; }
static float* tc_t(const MlpLayout& L, char* scratch, int k) { return L.Wf(scratch, 2) + (size_t)k * 256; }

// Linear correction of the fp16 weight rounding in the layered forward (k_tc_fold): 1 = on (default), 0 = off (the round-1
// operand layout; A/B measurements and tests).  PCNERF_TC_CORRECT=0|1 sets the initial value.
static int g_tc_correct = -1;
static int tc_correct_mode() {
    if (g_tc_correct < 0) { const char* e = getenv("PCNERF_TC_CORRECT"); g_tc_correct = e ? (atoi(e) != 0) : 1; }
    return g_tc_correct;
}
extern "C" void pcnerf_tc_set_weight_correction(int on) { g_tc_correct = on ? 1 : 0; }
extern "C" int pcnerf_tc_get_weight_correction(void) { return tc_correct_mode(); }

int mlp_tc_forward(const pcnerf_mlp_params* P, const void* enc, int64_t rows, float* out_p, void* saved, size_t,
                   void* scratch_v, size_t, cudaStream_t st) {
    const MlpLayout L(rows, 1);
    char* scratch = (char*)scratch_v;
    char* sv = (char*)saved;
    const __half* ench = (const __half*)enc;
    // (the fused eval kernel streams the round-1 operand layout: no correction blocks there)
    const int corr = (!P->training && g_fused_eval) ? 0 : tc_correct_mode();
    if (!P->prepared) {
        tc_prep_weights(P, L, scratch, st);
        PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_tc_prep_fwd<<<64, 256, 0, st>>>(L.Wp(scratch, 0), tc_Wh(L, scratch, 0), corr));
    }
    // H_l is stored once, fp16 (10-bit mantissa for the 1e-3 gate): it is the next layer's operand and what the backward
    // pass re-reads (the weight-gradient kernel converts its tiles to bf16 in shared memory)
    // Eval mode (running statistics): every fold is a function of the parameters alone, so the folded fp16 weights of all
    // layers are derived once per pass (first chunk) and the row GEMMs neither wait for nor compute batch statistics.
    const int ns = P->training ? 0 : 1;
    if (!P->training && g_fused_eval) {
        // ... and all nine layers run as ONE kernel with the activations resident in shared memory / TMEM (k_tc_fused_eval)
        if (!P->prepared) {
            for (int l = 0; l < 8; ++l) {
                const bool last = l == 7;
                PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0,
                          k_bn_fold<<<last ? 1 : 256, 256, 0, st>>>(
                              l, 0, rows, nullptr, nullptr, P->gamma[l], P->beta[l], P->running_mean[l], P->running_var[l],
                              P->num_batches_tracked[l], P->momentum, P->eps, L.coef(scratch) /* (mean, invstd, a, s): unused */,
                              last ? P->W[8] : L.Wp(scratch, l + 1), last ? P->b[8] : P->b[l + 1],
                              last ? L.wout_f(scratch) : L.Wf(scratch, l + 1),
                              last ? L.wout_f(scratch) + 256 : L.bf(scratch, l + 1), last ? nullptr : tc_Wh(L, scratch, l + 1)));
            }
        }
        CUtensorMap mE;
        FusedMaps wm;
        int rc = make_map(&mE, ench, rows, 64, 64, 128);
        if (rc) return rc;
        for (int l = 0; l < 8; ++l) {
            rc = make_map(&wm.w[l], tc_Wh(L, scratch, l), 256, mlp_kpad(l), mlp_kpad(l), 128);
            if (rc) return rc;
        }
        if (P->prepared != 1) {                           // 0 or 2: first chunk of a call (another model may have run since)
            PCN_CUDA(cudaMemcpyToSymbolAsync(c_fz, P->b[0], 256 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
            PCN_CUDA(cudaMemcpyToSymbolAsync(c_fz, L.bf(scratch, 1), 7 * 256 * sizeof(float), 256 * sizeof(float),
                                             cudaMemcpyDeviceToDevice, st));
            PCN_CUDA(cudaMemcpyToSymbolAsync(c_fz, L.wout_f(scratch), 257 * sizeof(float), 8 * 256 * sizeof(float),
                                             cudaMemcpyDeviceToDevice, st));
        }
        FusedArgs fa;
        {
            static int ca = -1;
            if (ca < 0) { const char* e = getenv("PCNERF_TC_EVAL_CLUSTER_ARRIVE"); ca = e ? (atoi(e) != 0) : 0; }
            fa.cluster_arrive = ca;
        }
        fa.rows = (int)rows;
        fa.out_p = out_p;
        const size_t smem = 1024 + (size_t)(8 + FZ_STAGES) * FZ_BLK;
        PcnScope ps(PCN_K_GEMM_FWD, st, (double)rows * 982528.0);
        if (g_fused_eval == 2) {
            // CTA pairs (cta_group::2): clusters of two CTAs, four row tiles per unit of work
            const int nunits = (int)pcn_cdiv(pcn_cdiv(rows, 128), 4);
            const int ncl = nunits < sm_count() / 2 ? nunits : sm_count() / 2;
            PCN_CUDA(cudaFuncSetAttribute(k_tc_fused_eval<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * ncl);
            cfg.blockDim = dim3(TC_THREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            PCN_CUDA(cudaLaunchKernelEx(&cfg, k_tc_fused_eval<2>, mE, wm, fa));
        } else {
            const int npairs = (int)pcn_cdiv(pcn_cdiv(rows, 128), 2);
            const int grid = npairs < sm_count() ? npairs : sm_count();
            PCN_CUDA(cudaFuncSetAttribute(k_tc_fused_eval<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_tc_fused_eval<1><<<grid, TC_THREADS, smem, st>>>(mE, wm, fa);
        }
        PCN_LAUNCH_CHECK();
        return 0;
    }
    PCN_CUDA(cudaMemsetAsync(L.dstat(scratch, 0), 0, sizeof(double) * 8 * 512, st));
    PCN_CUDA(cudaMemsetAsync(L.rgwork(scratch), 0, 4, st));
    for (int l = 0; l < 8; ++l) {
        __half* Hout = (__half*)L.Hraw(sv, l);
        const __half* Hin = l > 0 ? (const __half*)L.Hraw(sv, l - 1) : nullptr;
        __nv_bfloat16* Hsave = nullptr;
        double* s0 = L.dstat(scratch, l);
        const float* bias = l == 0 ? P->b[0] : L.bf(scratch, l);
        int rc;
        // operand layout of layer l (k_tc_fold): [encoding blocks | H_{l-1}] x [their weight blocks | fp16(W')]
        const int encb = corr ? (l == 0 || l == 4 ? 2 : 1) : (l == 0 || l == 4 ? 1 : 0);
        const int ldw = encb * 64 + (l == 0 ? 0 : 256);
        const int alg_k = mlp_kpad(l);                    // the Linear's own K: split / correction blocks are not algorithmic work
        if (l == 0) rc = launch_rowgemm(TC_FWD, ench, 64, 64, nullptr, 0, 0, tc_Wh(L, scratch, 0), ldw, bias, nullptr, rows, Hout, Hsave, s0, s0 + 256, L.rgwork(scratch), l & 1, st, ns, encb, alg_k);
        else if (encb) rc = launch_rowgemm(TC_FWD, ench, 64, 64, Hin, 256, 256, tc_Wh(L, scratch, l), ldw, bias, nullptr, rows, Hout, Hsave, s0, s0 + 256, L.rgwork(scratch), l & 1, st, ns, encb, alg_k);
        else rc = launch_rowgemm(TC_FWD, Hin, 256, 256, nullptr, 0, 0, tc_Wh(L, scratch, l), ldw, bias, nullptr, rows, Hout, Hsave, s0, s0 + 256, L.rgwork(scratch), l & 1, st, ns);
        if (rc) return rc;
        const bool last = l == 7;
        if (!P->training && P->prepared) continue;       // folded copies of the first chunk are still in `scratch`
        if (int rc2 = chain_before(l, st)) return rc2;   // (running statistics: after the previous chunk's update)
        if (last) {
            PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0,
                      k_bn_fold<<<1, 256, 0, st>>>(l, P->training, rows, s0, s0 + 256, P->gamma[l], P->beta[l], P->running_mean[l],
                                                   P->running_var[l], P->num_batches_tracked[l], P->momentum, P->eps,
                                                   L.stats(sv, l), P->W[8], P->b[8], L.wout_f(scratch),
                                                   L.wout_f(scratch) + 256, nullptr));
        } else {
            const int nl = l + 1;
            const int encn = corr ? (nl == 4 ? 2 : 1) : (nl == 4 ? 1 : 0);
            PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0,
                      k_tc_fold<<<256, 256, 0, st>>>(l, P->training, rows, s0, s0 + 256, P->gamma[l], P->beta[l],
                                                     P->running_mean[l], P->running_var[l], P->num_batches_tracked[l],
                                                     P->momentum, P->eps, L.stats(sv, l), L.Wp(scratch, nl), P->b[nl],
                                                     L.bf(scratch, nl), tc_Wh(L, scratch, nl), encn * 64 + 256,
                                                     l == 0 ? L.Wp(scratch, 0) : tc_T(L, scratch, l & 1),
                                                     l == 0 ? P->b[0] : tc_t(L, scratch, l & 1), tc_T(L, scratch, nl & 1),
                                                     tc_t(L, scratch, nl & 1), corr));
        }
        if (int rc2 = chain_after(l, st)) return rc2;
    }
    int64_t blocks = pcn_cdiv(rows, 8);
    if (blocks > PCN_SM_COUNT * 16) blocks = PCN_SM_COUNT * 16;
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0,
              k_logit_sigmoid<__half><<<(int)blocks, 256, 0, st>>>((const __half*)L.Hraw(sv, 7), rows, L.wout_f(scratch),
                                                                  L.wout_f(scratch) + 256, out_p));
    PCN_LAUNCH_CHECK();
    return 0;
}

int mlp_tc_backward(const pcnerf_mlp_params* P, const pcnerf_mlp_grads* G, const void* enc, int64_t rows,
                    const float* out_p, const float* grad_p, void* saved, size_t, void* scratch_v, size_t,
                    cudaStream_t st) {
    const MlpLayout L(rows, 1);
    char* scratch = (char*)scratch_v;
    char* sv = (char*)saved;
    const __half* ench = (const __half*)enc;
    PCN_CUDA(cudaMemsetAsync(L.dstat(scratch, 0), 0, sizeof(double) * L.n_dstat, st));
    PCN_CUDA(cudaMemsetAsync(L.rgwork(scratch), 0, 4, st));
    if (!P->prepared) {
        tc_prep_weights(P, L, scratch, st);
        PrepTArgs pa;
        for (int l = 0; l < 8; ++l) { pa.Wp[l] = L.Wp(scratch, l); pa.WT[l] = tc_WT(L, scratch, l); }
        PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_tc_prep_bwd<<<dim3(8, 8, 7), dim3(32, 8), 0, st>>>(pa));
    }
    double* acc_out = L.dstat(scratch, 8);
    float* gvec = L.gvec(scratch);
    float* coef = L.coef(scratch);
    __nv_bfloat16* Gb[2] = {(__nv_bfloat16*)L.Graw(scratch, 0), (__nv_bfloat16*)L.Graw(scratch, 1)};
    const int strips = (int)(pcn_cdiv(rows, STRIP) < 4 * PCN_SM_COUNT ? pcn_cdiv(rows, STRIP) : 4 * PCN_SM_COUNT);
    const __half* H7 = (const __half*)L.Hraw(sv, 7);
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0, k_out_bwd_reduce<__half><<<strips, 256, 0, st>>>(grad_p, out_p, H7, rows, gvec, acc_out));
    if (int rc2 = chain_before(8, st)) return rc2;       // (+= into the parameter gradients: after the previous chunk's)
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0,
              k_out_bwd_finalize<<<1, 256, 0, st>>>(acc_out, rows, P->W[8], L.stats(sv, 7), G->dW[8], G->db[8], G->dgamma[7],
                                                    G->dbeta[7], coef));
    if (int rc2 = chain_after(8, st)) return rc2;
    int cur = 0;
    PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0,
              k_bn_bwd_apply<true, __half, __nv_bfloat16><<<strips, 256, 0, st>>>(gvec, Gb[cur], H7, rows, coef, L.stats(sv, 7),
                                                                                  L.colsum(scratch, 7)));
    float* part = L.partial(scratch);
    // Per layer: weight gradient first (its raw result also yields the BN(l-1) backward coefficients, see
    // k_tc_bn_bwd_coef2), then the data-gradient GEMM whose epilogue applies the BN backward and emits DH_{l-1} directly.
    // tcgen05 kind::f16 needs A and B in the same 16-bit format (bf16 x fp16 is an illegal instruction on sm_100a): the
    // weight-gradient kernel converts the fp16 activation / encoding tiles to bf16 in shared memory.
    const bool d2 = tc_pairs_for(TC_DGRAD);              // data gradient on CTA pairs (k_tc_rowgemm2<TC_DGRAD2>)
    char* ops2 = (char*)L.Wf(scratch, 3);                // its operands (B1 | Dg | rtk), rewritten per layer
    for (int l = 7; l >= 0; --l) {
        const __nv_bfloat16* DH = Gb[cur];
        const int kpad = mlp_kpad(l), off = l == 4 ? 64 : 0;
        const __half* Hprev = l > 0 ? (const __half*)L.Hraw(sv, l - 1) : nullptr;
        // (a scattered in-kernel clear by k_tc_wgrad_finish was measured 3 ms/step slower than this memset node)
        PCN_CUDA(cudaMemsetAsync(part, 0, (size_t)256 * kpad * sizeof(float), st));
        int rc = 0;
        if (l == 0 || l == 4) rc = launch_wgrad(DH, ench, 64, 64, 0, rows, part, kpad, 0, st);
        if (rc) return rc;
        if (l != 0) rc = launch_wgrad(DH, Hprev, 256, 256, 0, rows, part, kpad, off, st);
        if (rc) return rc;
        if (int rc2 = chain_before(l, st)) return rc2;
        PCN_TIMED(PCN_K_MLP_SMALL, st, 0.0,
                  k_tc_wgrad_finish<<<kpad, 256, 0, st>>>(l, part, L.colsum(scratch, l), l == 0 ? nullptr : L.stats(sv, l - 1),
                                                          L.Wp(scratch, l), rows, G->dW[l], G->db[l],
                                                          l == 0 ? nullptr : G->dgamma[l - 1], l == 0 ? nullptr : G->dbeta[l - 1],
                                                          coef, d2 ? (__nv_bfloat16*)ops2 : nullptr,
                                                          d2 ? (__half*)(ops2 + 256 * 256 * 2) : nullptr,
                                                          d2 ? (float*)(ops2 + 256 * 256 * 2 + 256 * 64 * 2) : nullptr));
        if (int rc2 = chain_after(l, st)) return rc2;
        if (l == 0) break;
        if (d2)      // (its operands were written by k_tc_wgrad_finish above)
            rc = launch_dgrad2(DH, Hprev, nullptr, kpad, off, nullptr, coef, rows, Gb[cur ^ 1], L.colsum(scratch, l - 1),
                               L.rgwork(scratch), ops2, 1, st);
        else
            rc = launch_rowgemm(TC_DGRAD, DH, 256, 256, nullptr, 0, 0, tc_WT(L, scratch, l), 256, coef, Hprev, rows, Gb[cur ^ 1],
                                nullptr, L.colsum(scratch, l - 1), nullptr, L.rgwork(scratch), 1, st);
        if (rc) return rc;
        cur ^= 1;
    }
    PCN_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// All chunks of a pass, two in flight (see ChainSync)
// ---------------------------------------------------------------------------------------------------------------
#define TC_MAX_LANES 4
namespace {
struct Lanes {
    cudaStream_t stream[TC_MAX_LANES];
    cudaEvent_t fork, join[TC_MAX_LANES], slot[TC_MAX_LANES][9];
    bool ok = false;
};
int lanes_get(Lanes** out) {
    static Lanes L;
    if (!L.ok) {
        for (int k = 0; k < TC_MAX_LANES; ++k) {
            PCN_CUDA(cudaStreamCreateWithFlags(&L.stream[k], cudaStreamNonBlocking));
            PCN_CUDA(cudaEventCreateWithFlags(&L.join[k], cudaEventDisableTiming));
            for (int i = 0; i < 9; ++i) PCN_CUDA(cudaEventCreateWithFlags(&L.slot[k][i], cudaEventDisableTiming));
        }
        PCN_CUDA(cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming));
        L.ok = true;
    }
    *out = &L;
    return 0;
}
}  // namespace

static int tc_chunks(bool backward, const pcnerf_mlp_params* P, const pcnerf_mlp_grads* G, const void* enc, int64_t rows,
                     int64_t chunk, float* out_p, const float* grad_p, void* const* saved, const size_t* saved_bytes,
                     void* const* scratch, size_t scratch_bytes, int lanes, cudaStream_t st) {
    PCN_CHECK_ARG(P && P->precision == 1 && enc && out_p && saved && saved_bytes && scratch && rows >= 1 && chunk >= 1 &&
                      lanes >= 1 && lanes <= TC_MAX_LANES,
                  "mlp_tc_chunks: bad arguments (precision 1, 1..4 lanes)");
    {
        // the per-chunk checks of pcnerf_mlp_forward / pcnerf_mlp_backward, for every chunk, BEFORE anything is launched
        const int64_t nch = pcn_cdiv(rows, chunk);
        const int64_t big = rows < chunk ? rows : chunk;
        PCN_CHECK_ARG(big < (1ll << 31) / 320, "mlp_tc_chunks: chunk of %lld rows is too large (use a smaller chunk)", (long long)big);
        PCN_CHECK_ARG(scratch_bytes >= MlpLayout(big, 1).scratch_bytes, "mlp_tc_chunks: scratch too small (%zu < %zu)",
                      scratch_bytes, MlpLayout(big, 1).scratch_bytes);
        for (int k = 0; k < lanes; ++k) PCN_CHECK_ARG(scratch[k], "mlp_tc_chunks: null scratch[%d]", k);
        for (int64_t c = 0; c < nch; ++c) {
            const int64_t r = rows - c * chunk < chunk ? rows - c * chunk : chunk;
            if (P->training && r == 1) {                     // what torch's BatchNorm1d raises for a one-row batch
                pcn_set_error("Expected more than 1 value per channel when training, got input size [1, 256]");
                return PCNERF_ERR_ARG;
            }
            PCN_CHECK_ARG(saved[c] && saved_bytes[c] >= MlpLayout(r, 1).saved_bytes,
                          "mlp_tc_chunks: saved buffer of chunk %lld too small (%zu < %zu)", (long long)c, saved_bytes[c],
                          MlpLayout(r, 1).saved_bytes);
        }
    }
    Lanes* L = nullptr;
    if (lanes > 1) {
        if (int rc = lanes_get(&L)) return rc;
        PCN_CUDA(cudaEventRecord(L->fork, st));
        for (int k = 0; k < lanes; ++k) PCN_CUDA(cudaStreamWaitEvent(L->stream[k], L->fork, 0));
    }
    const int64_t nchunks = pcn_cdiv(rows, chunk);
    int rc = 0;
    for (int64_t c = 0; c < nchunks && !rc; ++c) {
        const int k = lanes > 1 ? (int)(c % lanes) : 0;
        const int kprev = lanes > 1 ? (int)((c + lanes - 1) % lanes) : 0;       // lane of chunk c - 1
        const int64_t r0 = c * chunk, r = rows - r0 < chunk ? rows - r0 : chunk;
        pcnerf_mlp_params Pc = *P;
        Pc.prepared = c >= lanes ? 1 : P->prepared;      // every lane derives its weight copies on its first chunk
        ChainSync cs;
        for (int i = 0; i < 9; ++i) {
            cs.rec[i] = lanes > 1 ? L->slot[k][i] : nullptr;
            cs.wait[i] = (lanes > 1 && c > 0) ? L->slot[kprev][i] : nullptr;
        }
        cudaStream_t s = lanes > 1 ? L->stream[k] : st;
        const void* e = (const char*)enc + (size_t)r0 * 64 * 2;
        g_chain = lanes > 1 ? &cs : nullptr;
        rc = backward ? mlp_tc_backward(&Pc, G, e, r, out_p + r0, grad_p + r0, saved[c], 0, scratch[k], 0, s)
                      : mlp_tc_forward(&Pc, e, r, out_p + r0, saved[c], 0, scratch[k], 0, s);
        g_chain = nullptr;
    }
    if (lanes > 1)
        for (int k = 0; k < lanes; ++k) {                 // join even after an error: never leave a capture forked
            cudaEventRecord(L->join[k], L->stream[k]);
            cudaStreamWaitEvent(st, L->join[k], 0);
        }
    return rc;
}

extern "C" int pcnerf_mlp_tc_forward_chunks(const pcnerf_mlp_params* P, const void* enc, int64_t rows, int64_t chunk,
                                            float* out_p, void* const* saved, const size_t* saved_bytes,
                                            void* const* scratch, size_t scratch_bytes, int lanes, void* stream) {
    PCN_CHECK_ARG(P && P->training, "mlp_tc_forward_chunks: training mode only (eval mode has no per-chunk state)");
    return tc_chunks(false, P, nullptr, enc, rows, chunk, out_p, nullptr, saved, saved_bytes, scratch, scratch_bytes, lanes,
                     (cudaStream_t)stream);
}

extern "C" int pcnerf_mlp_tc_backward_chunks(const pcnerf_mlp_params* P, const pcnerf_mlp_grads* G, const void* enc,
                                             int64_t rows, int64_t chunk, const float* out_p, const float* grad_p,
                                             void* const* saved, const size_t* saved_bytes, void* const* scratch,
                                             size_t scratch_bytes, int lanes, void* stream) {
    PCN_CHECK_ARG(G && grad_p, "mlp_tc_backward_chunks: null argument");
    PCN_CHECK_ARG(P && P->training, "mlp_tc_backward_chunks: only the training-mode (batch-statistics) backward exists");
    return tc_chunks(true, P, G, enc, rows, chunk, const_cast<float*>(out_p), grad_p, saved, saved_bytes, scratch,
                     scratch_bytes, lanes, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Building-block entry points (unit-tested against torch.matmul; also usable on their own)
// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t pcnerf_tc_rowgemm_work_bytes(void) { return TC_ROWGEMM_WORK_BYTES; }

extern "C" int pcnerf_tc_rowgemm(int mode, const void* A0, int k0, const void* A1, int k1, const void* B,
                                 const float* vec, const void* E, int64_t rows, void* out, void* out2, double* stats,
                                 void* work, void* stream) {
    PCN_CHECK_ARG(mode == 0 || mode == 1, "tc_rowgemm: mode must be 0 (fp16 forward) or 1 (bf16 data gradient)");
    PCN_CHECK_ARG(A0 && B && out && stats && vec && work && rows >= 1, "tc_rowgemm: null argument");
    PCN_CHECK_ARG(mode == 0 || E, "tc_rowgemm: data-gradient mode needs E");
    PCN_CHECK_ARG(out2 == nullptr, "tc_rowgemm: out2 is reserved and must be NULL (activations are stored once, fp16)");
    cudaStream_t st = (cudaStream_t)stream;
    PCN_CUDA(cudaMemsetAsync(stats, 0, 512 * sizeof(double), st));
    PCN_CUDA(cudaMemsetAsync(work, 0, 4, st));
    if (mode == 1 && tc_pairs_for(TC_DGRAD)) {
        PCN_CHECK_ARG(k0 == 256 && k1 == 0, "tc_rowgemm: the data-gradient form on CTA pairs needs K = 256");
        return launch_dgrad2(A0, (const __half*)E, nullptr, 0, 0, (const __nv_bfloat16*)B, vec, rows, out, stats, work,
                             (char*)work + TC_ROWGEMM_WORK_CORE, mode, st);
    }
    return launch_rowgemm(mode, A0, k0, k0, A1, k1, k1, B, k0 + k1, vec, (const __half*)E, rows, out, nullptr, stats,
                          stats + 256, work, mode, st);
}

extern "C" int pcnerf_tc_wgrad(const void* DH, const void* X, int ldx, int ncols, int x_is_bf16, int64_t rows,
                               float* out, int ldo, int col_off, void* stream) {
    PCN_CHECK_ARG(DH && X && out && rows >= 1, "tc_wgrad: null argument");
    PCN_CHECK_ARG(ldo >= col_off + ncols && (ldo % 4) == 0 && (col_off % 4) == 0, "tc_wgrad: bad output window");

    return launch_wgrad(DH, X, ldx, ncols, x_is_bf16, rows, out, ldo, col_off, (cudaStream_t)stream);
}

extern "C" int pcnerf_tc_last_fault(void) {
    int v = 0;
    cudaMemcpyFromSymbol(&v, g_tc_err, sizeof(int));
    return v;
}
