// Micro-benchmark: how fast can ONE SM fill a shared-memory ring through TMA, by box shape and source?
//   mode 0: 2-D tensor box {64 cols x 128 rows} SWIZZLE_128B out of a row-major [R][256] fp16 matrix (the row GEMM's A tile,
//           the fused kernel's weight stage): 128 segments of 128 B, 512 B apart
//   mode 1: 1-D bulk copy of 16 KB contiguous bytes (what a pre-tiled operand would allow)
//   mode 2: 2-D tensor box {64 x 128} out of a [R][64] matrix (rows contiguous: one 16 KB run, still a tensor copy)
// src 0: every CTA streams the same 1 MB (L2 hits, like the folded weights); src 1: every CTA streams its own slab of a
// 2 GB buffer (HBM, like activations).  A consumer warp releases each stage as soon as it lands (no math).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define STAGE 16384
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}

__global__ void __launch_bounds__(64, 1) k_fill(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int mode, int src,
                                                int nstage, int iters, long long slab_rows, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[32];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t full = su32(bars), empty = su32(bars + 16);
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        int s = 0; uint32_t ph = 0;
        const long long row0 = src ? (long long)blockIdx.x * slab_rows : 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(empty + 8 * s, ph ^ 1);
            mbar_expect(full + 8 * s, STAGE);
            // walk the source in 16 KB steps: mode 0: (kb, row tile) over [R][256]; mode 2: row tile over [R][64]
            const long long step = src ? it : (it & 63);
            if (mode == 1) {
                const uint8_t* g = base + (row0 * 512) + step * STAGE;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(su32(smem + s * STAGE)), "l"(g), "r"(STAGE), "r"(full + 8 * s) : "memory");
            } else {
                int c0, c1;
                if (mode == 0) { c0 = (int)(step & 3) * 64; c1 = (int)(row0 + (step >> 2) * 128); }
                else { c0 = 0; c1 = (int)(row0 * 4 + step * 128); }
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(su32(smem + s * STAGE)), "l"(&tm), "r"(full + 8 * s), "r"(c0), "r"(c1) : "memory");
            }
            if (++s == nstage) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full + 8 * s, ph);
            mbar_arrive(empty + 8 * s);
            if (++s == nstage) { s = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncFn enc = (EncFn)fp;
    const size_t bytes = 2ull << 30;
    uint8_t* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const long long rows256 = bytes / 512;
    for (int src = 0; src < 2; ++src)
        for (int mode = 0; mode < 3; ++mode)
            for (int nstage = 2; nstage <= 12; nstage += (nstage < 8 ? 2 : 4)) {
                CUtensorMap tm;
                cuuint64_t dims[2], strides[1]; cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
                if (mode == 2) { dims[0] = 64; dims[1] = (cuuint64_t)rows256 * 4; strides[0] = 128; }
                else { dims[0] = 256; dims[1] = (cuuint64_t)rows256; strides[0] = 512; }
                enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                const int iters = 800;                                   // 12.5 MB per CTA
                const long long slab_rows = rows256 / 148;
                const size_t smem = 1024 + (size_t)nstage * STAGE;
                cudaFuncSetAttribute(k_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k_fill<<<148, 64, smem>>>(tm, buf, mode, src, nstage, iters, slab_rows, cyc);
                cudaEventRecord(e0);
                k_fill<<<148, 64, smem>>>(tm, buf, mode, src, nstage, iters, slab_rows, cyc);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
                printf("src=%s mode=%d stages=%2d  %7.1f us  %6.2f TB/s chip  %5.1f B/clk/SM  (%s)\n", src ? "hbm" : "l2 ", mode, nstage,
                       ms * 1e3, 148.0 * iters * STAGE / (ms * 1e-3) / 1e12, (double)iters * STAGE / avg, cudaGetErrorString(err));
            }
    return 0;
}
