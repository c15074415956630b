// K0: LiDAR frame -> returns of the block (nof/dataset/ipb2dmapping.py:662-711 KITTI, :319-345 MaiCity:
// no height window, no interest region, but a parent-box test on the transformed points).
//
// One thread per raw point.  The reference filters the sensor-frame float32 points with numpy masks (near-sensor box,
// 120 m range gate, height window), transforms the survivors with the frame pose in float64 (pose entries are float32
// values), keeps those within the interest region of any pose of the run (float32 arithmetic: numpy scalar minus 0-d
// float32 tensor), and derives ray direction and range from the sensor position.  All of it is elementwise: the kernel
// evaluates every stage for every point and writes a keep flag; the ordered compaction is the caller's.
// Built with -fmad=false: every product and sum rounds on its own, like the numpy expressions it replaces.
#include "common.cuh"

struct FrameParams {
    double pose[12];            // rows 0..2 of the 4x4 frame pose (float32 values)
    double pos[3];              // sensor position of the frame = pose[:3, 3] (float32 values)
    float rdx, rdy, rdz;        // near-sensor box: keep |x| >= rdx or |y| >= rdy or |z| >= rdz
    float max_range;            // 120
    float over_height, over_low;
    float interest_x, interest_y;
    int npose;                  // 0: no interest-region test (MaiCity loader)
    int use_box;                // 1: keep only world points inside the closed parent box (MaiCity loader, :334-336)
    double box[6];              // x_min, x_max, y_min, y_max, z_min, z_max
};

__global__ void k_frame_returns(const float* __restrict__ pts, int64_t n, FrameParams fp,
                                const float* __restrict__ pose_xy, uint8_t* __restrict__ keep,
                                double* __restrict__ world, double* __restrict__ dir, double* __restrict__ dist) {
    extern __shared__ float sxy[];
    for (int i = threadIdx.x; i < 2 * fp.npose; i += blockDim.x) sxy[i] = pose_xy[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    bool k = fabsf(x) >= fp.rdx || fabsf(y) >= fp.rdy || fabsf(z) >= fp.rdz;               // :666-668
    const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    k = k && r <= fp.max_range && z <= fp.over_height && z >= fp.over_low;                 // :670-678
    const double X = x, Y = y, Z = z;
    double w[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)                                                            // pose @ [x y z 1]^T  (:683)
        w[a] = ((fp.pose[4 * a] * X + fp.pose[4 * a + 1] * Y) + fp.pose[4 * a + 2] * Z) + fp.pose[4 * a + 3];
    // interest region (:690-699): |x - pose_k.x| <= interest_x and |y - pose_k.y| <= interest_y for some pose k
    const float xf = (float)w[0], yf = (float)w[1];
    bool near = fp.npose == 0;
    for (int p = 0; p < fp.npose && !near; ++p)
        near = !(fabsf(__fsub_rn(xf, sxy[2 * p])) > fp.interest_x || fabsf(__fsub_rn(yf, sxy[2 * p + 1])) > fp.interest_y);
    k = k && near;
    if (fp.use_box)
        k = k && w[0] >= fp.box[0] && w[1] >= fp.box[2] && w[2] >= fp.box[4] && w[0] <= fp.box[1] && w[1] <= fp.box[3] && w[2] <= fp.box[5];
    const double vx = w[0] - fp.pos[0], vy = w[1] - fp.pos[1], vz = w[2] - fp.pos[2];       // :706-709
    const double d = sqrt((vx * vx + vy * vy) + vz * vz);
    keep[i] = k ? 1 : 0;
    world[3 * i] = w[0]; world[3 * i + 1] = w[1]; world[3 * i + 2] = w[2];
    dir[3 * i] = vx / d; dir[3 * i + 1] = vy / d; dir[3 * i + 2] = vz / d;
    dist[i] = d;
}

extern "C" int pcnerf_frame_returns(const float* pts, int64_t n, const double* h_pose16, const float* pose_xy, int npose,
                                    float rdx, float rdy, float rdz, float max_range, float over_height, float over_low,
                                    float interest_x, float interest_y, const double* h_box6, const double* h_pos3,
                                    uint8_t* keep, double* world, double* dir, double* dist, void* stream) {
    PCN_CHECK_ARG(n >= 0 && h_pose16 && npose >= 0 && npose <= 8192, "frame_returns: bad arguments (at most 8192 poses)");
    if (n == 0) return 0;
    PCN_CHECK_ARG(pts && keep && world && dir && dist && (npose == 0 || pose_xy), "frame_returns: null argument");
    FrameParams fp;
    for (int i = 0; i < 12; ++i) fp.pose[i] = h_pose16[i];
    for (int a = 0; a < 3; ++a) fp.pos[a] = h_pos3 ? h_pos3[a] : h_pose16[4 * a + 3];
    fp.rdx = rdx; fp.rdy = rdy; fp.rdz = rdz; fp.max_range = max_range; fp.over_height = over_height; fp.over_low = over_low;
    fp.interest_x = interest_x; fp.interest_y = interest_y; fp.npose = npose;
    fp.use_box = h_box6 ? 1 : 0;
    for (int i = 0; i < 6; ++i) fp.box[i] = h_box6 ? h_box6[i] : 0.0;
    const size_t smem = (size_t)2 * npose * sizeof(float);
    if (smem > 48 * 1024) PCN_CUDA(cudaFuncSetAttribute(k_frame_returns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PcnScope ps(PCN_K_AABB, (cudaStream_t)stream, (double)n * (12.0 + 1.0 + 56.0));
    k_frame_returns<<<(int)pcn_cdiv(n, 256), 256, smem, (cudaStream_t)stream>>>(pts, n, fp, pose_xy, keep, world, dir, dist);
    PCN_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-parent scenes (BASELINE.json configs[4]; README.md:46: a large scene is a collection of parent NeRF blocks, each
// with its own networks and child boxes).  One thread per return: index of the FIRST parent block whose closed box contains
// the point (-1: none), and the ray from the sensor position of the return's frame (direction, range: the expressions of
// eval_kitti_render.py:706-709).  The parent boxes (P x 6 doubles: min xyz, max xyz) sit in shared memory.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_route_points(const double* __restrict__ pts, int64_t n, const double* __restrict__ boxes, int P,
                               const double* __restrict__ origins, const int32_t* __restrict__ frame_id,
                               int32_t* __restrict__ which, double* __restrict__ origin_out, double* __restrict__ dir,
                               double* __restrict__ dist) {
    extern __shared__ double sb[];
    for (int i = threadIdx.x; i < 6 * P; i += blockDim.x) sb[i] = boxes[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    int w = -1;
    for (int b = 0; b < P; ++b) {
        const double* q = sb + 6 * b;
        if (x >= q[0] && y >= q[1] && z >= q[2] && x <= q[3] && y <= q[4] && z <= q[5]) { w = b; break; }
    }
    which[i] = w;
    if (origins) {
        const double* o = origins + 3 * (frame_id ? frame_id[i] : 0);
        const double vx = x - o[0], vy = y - o[1], vz = z - o[2];
        const double d = sqrt((vx * vx + vy * vy) + vz * vz);
        origin_out[3 * i] = o[0]; origin_out[3 * i + 1] = o[1]; origin_out[3 * i + 2] = o[2];
        dir[3 * i] = vx / d; dir[3 * i + 1] = vy / d; dir[3 * i + 2] = vz / d;
        dist[i] = d;
    }
}

extern "C" int pcnerf_route_points(const double* pts, int64_t n, const double* boxes, int P, const double* origins,
                                   const int32_t* frame_id, int32_t* which, double* origin_out, double* dir, double* dist,
                                   void* stream) {
    PCN_CHECK_ARG(n >= 0 && P >= 1 && P <= 1024 && boxes && (n == 0 || (pts && which)), "route_points: bad arguments (at most 1024 parent blocks)");
    PCN_CHECK_ARG(!origins || (origin_out && dir && dist), "route_points: origins given but no ray outputs");
    if (n == 0) return 0;
    const size_t smem = (size_t)6 * P * sizeof(double);
    PcnScope ps(PCN_K_AABB, (cudaStream_t)stream, (double)n * (24.0 + 4.0 + (origins ? 84.0 : 0.0)));
    k_route_points<<<(int)pcn_cdiv(n, 256), 256, smem, (cudaStream_t)stream>>>(pts, n, boxes, P, origins, frame_id, which,
                                                                               origin_out, dir, dist);
    PCN_LAUNCH_CHECK();
    return 0;
}
