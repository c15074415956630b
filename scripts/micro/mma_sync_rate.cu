// Micro-benchmark: issue rate of the LEGACY tensor path (mma.sync) on sm_100a, register operands only -- the question behind
// "would a 3 x TF32 mma.sync form of the closed-form engine's second-moment kernel (csrc/affine_rays.cu) beat its packed-FFMA2
// form?".  Every warp keeps ACC independent accumulator tiles and issues MMAs back to back; prints MMAs per clock and SM and the
// equivalent dense TFLOP/s for 4 / 8 / 16 warps per SM.
//   kind 0: mma.sync.m16n8k8  tf32 (1,024 FMA per MMA)     kind 1: mma.sync.m16n8k16 bf16 (2,048 FMA)
//   kind 2: fma.rn.f32x2 (packed FFMA2, 64 FMA per warp instruction) for scale
//   kind 3: tf32 MMAs with a DIFFERENT A quad and B pair per accumulator (no operand reuse between consecutive MMAs)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ACC 8

template <int KIND>
__global__ void k_rate(int iters, float* out, long long* cycles) {
    float c[ACC][4];
#pragma unroll
    for (int t = 0; t < ACC; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) c[t][q] = 0.f;
    uint32_t a[4], b[2];
    a[0] = 0x3f800000u + threadIdx.x; a[1] = 0x3f000000u + threadIdx.x; a[2] = 0x3e800000u; a[3] = 0x3f800000u;
    b[0] = 0x3f800000u; b[1] = 0x3f000000u + blockIdx.x;
    uint32_t av[ACC][4], bv[ACC][2];
#pragma unroll
    for (int t = 0; t < ACC; ++t) {
#pragma unroll
        for (int q = 0; q < 4; ++q) av[t][q] = 0x3f800000u + 64u * t + q + threadIdx.x;
        bv[t][0] = 0x3f000000u + t; bv[t][1] = 0x3e800000u + t + blockIdx.x;
    }
    unsigned long long p[ACC];
#pragma unroll
    for (int t = 0; t < ACC; ++t) p[t] = 0ull;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < ACC; ++t) {
            if (KIND == 0) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            } else if (KIND == 1) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            } else if (KIND == 3) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                             : "r"(av[t][0]), "r"(av[t][1]), "r"(av[t][2]), "r"(av[t][3]), "r"(bv[t][0]), "r"(bv[t][1]));
            } else {
                unsigned long long aa, bb;
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "r"(a[0]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "r"(b[0]), "r"(b[1]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[t]) : "l"(aa), "l"(bb));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < ACC; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3] + (float)(p[t] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
static void run(const char* name, double fma_per_op, int warps) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * warps * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const int iters = 20000;
    k_rate<KIND><<<sms, warps * 32>>>(100, out, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_rate<KIND><<<sms, warps * 32>>>(iters, out, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double ops_sm = (double)iters * ACC * warps;
    printf("%-28s warps/SM %2d: %7.3f ops/clk/SM, %8.1f FMA/clk/SM, %8.1f TFLOP/s (%d SMs, %.3f ms, %lld clk)\n", name, warps,
           ops_sm / (double)h, ops_sm * fma_per_op / (double)h, 2.0 * ops_sm * fma_per_op * sms / (ms * 1e-3) / 1e12, sms, ms, h);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) run<0>("mma.sync m16n8k8 tf32", 1024.0, w);
    for (int w : {4, 8, 16}) run<1>("mma.sync m16n8k16 bf16", 2048.0, w);
    for (int w : {4, 8, 16}) run<2>("fma.rn.f32x2 (FFMA2)", 64.0, w);
    for (int w : {4, 8, 16}) run<3>("mma.sync tf32, distinct operands", 1024.0, w);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
