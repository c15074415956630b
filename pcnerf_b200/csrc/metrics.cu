// K6 (the step immediately downstream of the render path) -- point-cloud metrics.
// nn_correspondance of nof/criteria/pointcloud_metrics.py:5-33 builds an Open3D KDTreeFlann over `verts1` and queries the
// 1-nearest neighbour of every vertex of `verts2` (exact search, float64, squared L2 then np.sqrt).  On the device the
// exact nearest neighbour is found by brute force in float64: one thread per query, reference points streamed through
// shared memory in tiles.  No FMA contraction in the distance (build flag -fmad=false), accumulation order x, y, z.
#include "common.cuh"

#define NN_TILE 1024
#define NN_THREADS 256

__global__ void __launch_bounds__(NN_THREADS) k_nn_bruteforce(const double* __restrict__ ref, int64_t n1,
                                                              const double* __restrict__ qry, int64_t n2,
                                                              int32_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    __shared__ double sx[NN_TILE], sy[NN_TILE], sz[NN_TILE];
    const int64_t q = (int64_t)blockIdx.x * NN_THREADS + threadIdx.x;
    double qx = 0, qy = 0, qz = 0;
    if (q < n2) { qx = qry[3 * q]; qy = qry[3 * q + 1]; qz = qry[3 * q + 2]; }
    double best = INFINITY;
    int32_t bi = -1;
    for (int64_t t0 = 0; t0 < n1; t0 += NN_TILE) {
        const int cnt = (int)min((int64_t)NN_TILE, n1 - t0);
        for (int i = threadIdx.x; i < cnt; i += NN_THREADS) {
            sx[i] = ref[3 * (t0 + i)]; sy[i] = ref[3 * (t0 + i) + 1]; sz[i] = ref[3 * (t0 + i) + 2];
        }
        __syncthreads();
        if (q < n2) {
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) {
                const double dx = qx - sx[i], dy = qy - sy[i], dz = qz - sz[i];
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best) { best = d2; bi = (int32_t)(t0 + i); }      // strict <: the first of equal distances wins
            }
        }
        __syncthreads();
    }
    if (q < n2) {
        out_idx[q] = bi;
        out_dist[q] = sqrt(best);
    }
}

// precision = mean(dist1 < thr), recall = mean(dist2 < thr), sums of dist1 / dist2 -> sums[4] f64 (zeroed by the caller)
__global__ void __launch_bounds__(256) k_dist_stats(const double* __restrict__ d, int64_t n, double thr, double* __restrict__ sums) {
    double s = 0, c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = d[i];
        s += v;
        c += v < thr ? 1.0 : 0.0;
    }
    s = warp_sum_d(s);
    c = warp_sum_d(c);
    if ((threadIdx.x & 31) == 0) { atomicAdd(sums, s); atomicAdd(sums + 1, c); }
}

extern "C" int pcnerf_nn_correspondance(const double* verts1, int64_t n1, const double* verts2, int64_t n2,
                                        int32_t* out_idx, double* out_dist, void* stream) {
    PCN_CHECK_ARG(n1 >= 0 && n2 >= 0 && n1 < (1ll << 31), "nn_correspondance: bad sizes");
    if (n1 == 0 || n2 == 0) return 0;
    PCN_CHECK_ARG(verts1 && verts2 && out_idx && out_dist, "nn_correspondance: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    PcnScope ps(PCN_K_SEARCH, st, (double)(n1 + n2) * 24.0 + (double)n2 * 12.0);
    k_nn_bruteforce<<<(unsigned)pcn_cdiv(n2, NN_THREADS), NN_THREADS, 0, st>>>(verts1, n1, verts2, n2, out_idx, out_dist);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_dist_stats(const double* dist, int64_t n, double threshold, double* sums2, void* stream) {
    PCN_CHECK_ARG(dist && sums2 && n >= 1, "dist_stats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PCN_CUDA(cudaMemsetAsync(sums2, 0, 2 * sizeof(double), st));
    int64_t blocks = pcn_cdiv(n, 256);
    if (blocks > PCN_SM_COUNT * 4) blocks = PCN_SM_COUNT * 4;
    PcnScope ps(PCN_K_SEARCH, st, (double)n * 8.0);
    k_dist_stats<<<(int)blocks, 256, 0, st>>>(dist, n, threshold, sums2);
    PCN_LAUNCH_CHECK();
    return 0;
}
