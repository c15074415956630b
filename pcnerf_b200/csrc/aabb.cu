// K1: AABB stage in fp64 (compiled with -fmad=false: the reference is numpy / python-scalar arithmetic with
// no fused multiply-add, and child selection / segment bounds must be bit-exact).
//
// Mapping: one thread per ray (a warp = a 32-ray tile); the child boxes are staged once per CTA in shared
// memory (persistent CTAs loop over ray tiles), every lane reads the same box -> broadcast, no bank conflicts.
#include "common.cuh"
#include <limits.h>

#define AABB_THREADS 128

// ---------------------------------------------------------------------------------------------------------------
// leaf predicates
// ---------------------------------------------------------------------------------------------------------------

// compute_far_bound, nof/dataset/ipb2dmapping.py:36-77
__device__ __forceinline__ double far_bound(const double o[3], const double d[3], const double pl[6]) {
    double t = INFINITY;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if (d[ax] != 0.0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double tt = (pl[2 * ax + j] - o[ax]) / d[ax];
                if (tt < 0) tt = INFINITY;
                t = fmin(t, tt);   // np.min over the six values (NaN cannot occur: d != 0)
            }
        }
    }
    return t;
}

// shared body of compute_far_bound0406/0606/0429 (ipb2dmapping.py:82-145, eval_kitti_render.py:170-196).
// Returns the number of valid hits; dist[] holds them in the reference's append order.
__device__ __forceinline__ int plane_hits(const double p[3], const double d[3], const double bmin[3],
                                          const double bmax[3], double dist[6]) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const double plane = s == 0 ? bmin[i] : bmax[i];
            if (d[i] * (plane - p[i]) > 0) {
                const double distance = (plane - p[i]) / d[i];
                int count = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k == i) continue;
                    const double pe = p[k] + distance * d[k];
                    if (pe >= bmin[k] && pe <= bmax[k]) count++;
                }
                if (count >= 2) dist[n++] = distance;
            }
        }
    }
    return n;
}

// eval_kitti_render.py:198-211: exactly two hits
__device__ __forceinline__ bool child_0429(const double p[3], const double d[3], const double* bmin,
                                           const double* bmax, double& nr, double& fr) {
    double dist[6];
    const int n = plane_hits(p, d, bmin, bmax, dist);
    if (n != 2) { nr = 0; fr = 0; return false; }
    nr = dist[0]; fr = dist[1];
    if (nr > fr) { double t = nr; nr = fr; fr = t; }
    return true;
}

// ipb2dmapping.py:147-172
__device__ __forceinline__ bool child_0606(const double p[3], const double d[3], const double* bmin,
                                           const double* bmax, double& nr, double& fr) {
    double dist[6];
    const int n = plane_hits(p, d, bmin, bmax, dist);
    if (n == 0) { nr = 0; fr = 0; return false; }
    nr = dist[0]; fr = dist[0];
    for (int i = 1; i < n; ++i) { if (nr > dist[i]) nr = dist[i]; if (fr < dist[i]) fr = dist[i]; }
    return true;
}

// ipb2dmapping.py:109-114 (IndexError in the reference with < 2 hits -> NaN here)
__device__ __forceinline__ bool child_0406(const double p[3], const double d[3], const double* bmin,
                                           const double* bmax, double& nr, double& fr) {
    double dist[6];
    const int n = plane_hits(p, d, bmin, bmax, dist);
    if (n < 2) { nr = NAN; fr = NAN; return false; }
    nr = dist[0]; fr = dist[1];
    if (nr > fr) { double t = nr; nr = fr; fr = t; }
    return true;
}

// distance_to_ray, eval_kitti_render.py:237-244 (one centre)
__device__ __forceinline__ double dist_to_ray(const double o[3], const double d[3], const double c[3]) {
    const double v0 = c[0] - o[0], v1 = c[1] - o[1], v2 = c[2] - o[2];
    const double dist = sqrt((v0 * v0 + v1 * v1) + v2 * v2);
    const double cosang = ((v0 * d[0] + v1 * d[1]) + v2 * d[2]) / dist;
    const double sinang = sqrt(1 - cosang * cosang);
    return dist * sinang;
}

__device__ __forceinline__ void load3(const double* src, int64_t i, double out[3]) {
    out[0] = src[3 * i]; out[1] = src[3 * i + 1]; out[2] = src[3 * i + 2];
}

// ---------------------------------------------------------------------------------------------------------------
// simple per-ray kernels
// ---------------------------------------------------------------------------------------------------------------

struct Six { double v[6]; };

__global__ void k_far_bound(const double* __restrict__ ro, const double* __restrict__ rd, int64_t n, Six pl,
                            double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double o[3], d[3];
        load3(ro, i, o); load3(rd, i, d);
        out[i] = far_bound(o, d, pl.v);
    }
}

// ray_aabb_distances, eval_kitti_render.py:213-235; box = {min xyz, max xyz}
__device__ __forceinline__ double slab_far(const double o[3], const double d[3], const double* b) {
    double tmin = -INFINITY, tmax = INFINITY;
    bool nan_min = false, nan_max = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double t1 = (b[a] - o[a]) / d[a];
        const double t2 = (b[3 + a] - o[a]) / d[a];
        // np.minimum / np.maximum / np.max / np.min propagate NaN
        double lo = fmin(t1, t2), hi = fmax(t1, t2);
        if (isnan(t1) || isnan(t2)) { lo = NAN; hi = NAN; }
        if (isnan(lo)) nan_min = true;
        if (isnan(hi)) nan_max = true;
        tmin = fmax(tmin, lo);
        tmax = fmin(tmax, hi);
    }
    if (nan_min) tmin = NAN;
    if (nan_max) tmax = NAN;
    return (tmax >= tmin) ? tmax : INFINITY;
}

__global__ void k_slab(const double* __restrict__ ro, const double* __restrict__ rd, int64_t n, Six box,
                       double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double o[3], d[3];
        load3(ro, i, o); load3(rd, i, d);
        out[i] = slab_far(o, d, box.v);
    }
}

// Stage `count` doubles from global to shared (whole CTA), or return the global pointer when they do not fit.
__device__ __forceinline__ const double* stage(const double* g, double* s, int count, bool use_smem) {
    if (!use_smem) return g;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s[i] = g[i];
    return s;
}

template <int VARIANT>
__global__ void k_child_pairs(const double* __restrict__ ro, const double* __restrict__ rd, int64_t n,
                              const double* __restrict__ boxes, int K, int use_smem, uint8_t* __restrict__ oflag,
                              double* __restrict__ onear, double* __restrict__ ofar) {
    extern __shared__ double sm[];
    const double* bx = stage(boxes, sm, K * 6, use_smem);
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double o[3], d[3];
        load3(ro, i, o); load3(rd, i, d);
        for (int k = 0; k < K; ++k) {
            double nr, fr;
            bool f;
            if (VARIANT == 429) f = child_0429(o, d, bx + 6 * k, bx + 6 * k + 3, nr, fr);
            else if (VARIANT == 606) f = child_0606(o, d, bx + 6 * k, bx + 6 * k + 3, nr, fr);
            else f = child_0406(o, d, bx + 6 * k, bx + 6 * k + 3, nr, fr);
            oflag[i * K + k] = f;
            onear[i * K + k] = nr;
            ofar[i * K + k] = fr;
        }
    }
}

__global__ void k_dist_to_ray(const double* __restrict__ ro, const double* __restrict__ rd, int64_t n,
                              const double* __restrict__ centres, int K, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double o[3], d[3];
        load3(ro, i, o); load3(rd, i, d);
        for (int k = 0; k < K; ++k) out[i * K + k] = dist_to_ray(o, d, centres + 3 * k);
    }
}

// find_aabb_box, ipb2dmapping.py:174-197: the containing box with the smallest centre distance, provided fewer than
// `knn` centres are strictly closer (== "first containing box among the knn nearest centres").
__device__ __forceinline__ int find_box(const double q[3], const double* centres, const double* boxes, int K, int knn) {
    double best = INFINITY;
    int bi = -1;
    for (int k = 0; k < K; ++k) {
        const double* b = boxes + 6 * k;
        if (q[0] >= b[0] && q[1] >= b[1] && q[2] >= b[2] && q[0] <= b[3] && q[1] <= b[4] && q[2] <= b[5]) {
            const double* c = centres + 3 * k;
            const double t0 = q[0] - c[0], t1 = q[1] - c[1], t2 = q[2] - c[2];
            const double r = (t0 * t0 + t1 * t1) + t2 * t2;
            if (r < best) { best = r; bi = k; }
        }
    }
    if (bi < 0) return -1;
    int rank = 0;
    for (int k = 0; k < K; ++k) {
        const double* c = centres + 3 * k;
        const double t0 = q[0] - c[0], t1 = q[1] - c[1], t2 = q[2] - c[2];
        const double r = (t0 * t0 + t1 * t1) + t2 * t2;
        rank += (r < best);
    }
    return rank < knn ? bi : -1;
}

__global__ void k_find_box(const double* __restrict__ centres, const double* __restrict__ boxes, int K,
                           const double* __restrict__ pts, int64_t nq, int knn, int use_smem,
                           int32_t* __restrict__ out) {
    extern __shared__ double sm[];
    const double* bx = stage(boxes, sm, K * 6, use_smem);
    const double* cx = stage(centres, sm + (use_smem ? K * 6 : 0), K * 3, use_smem);
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        double q[3];
        load3(pts, i, q);
        out[i] = find_box(q, cx, bx, K, knn);
    }
}

// ipb2dmapping.py:367-397 (406) / :736-768 (606) + the record layout of :447-452
template <int VARIANT>
__global__ void k_pack_train(const double* __restrict__ ro, const double* __restrict__ rd,
                             const double* __restrict__ dist, const double* __restrict__ pts, int64_t n,
                             const double* __restrict__ centres, const double* __restrict__ boxes,
                             const double* __restrict__ bigger, int K, Six parent, double se, int knn, int use_smem,
                             float* __restrict__ orays, uint8_t* __restrict__ okeep) {
    extern __shared__ double sm[];
    const double* bx = stage(boxes, sm, K * 6, use_smem);
    const double* cx = stage(centres, sm + (use_smem ? K * 6 : 0), K * 3, use_smem);
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double o[3], d[3], q[3];
        load3(ro, i, o); load3(rd, i, d); load3(pts, i, q);
        float* r = orays + 15 * i;
        const int idx = find_box(q, cx, bx, K, knn);
        bool keep = idx >= 0;
        double nr = 0, fr = 0;
        if (keep) {
            const double* b = bigger + 6 * idx;
            if (VARIANT == 406) child_0406(o, d, b, b + 3, nr, fr);
            else keep = child_0606(o, d, b, b + 3, nr, fr);
        }
        okeep[i] = keep;
        if (!keep) {
            for (int c = 0; c < 15; ++c) r[c] = 0.f;
            continue;
        }
        nr = nr - se;
        fr = fr + se;
        double fp = far_bound(o, d, parent.v);
        if (fp < fr) fp = fr;
        r[0] = (float)o[0]; r[1] = (float)o[1]; r[2] = (float)o[2];
        r[3] = (float)d[0]; r[4] = (float)d[1]; r[5] = (float)d[2];
        r[6] = 0.f; r[7] = (float)fp; r[8] = 3.f; r[9] = (float)(idx + 1);
        r[10] = (float)nr; r[11] = (float)fr; r[12] = (float)(dist[i] - se);
        r[13] = (float)fr;                     // ipb2dmapping.py:443 packs far_bound again, not far_bound_point
        r[14] = (float)dist[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// candidate groups (eval_kitti_render.py:353-461 / :681-803)
// ---------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ bool prefilter_pass(const double o[3], const double d[3], const double* b, double thr) {
    double c[3] = {(b[0] + b[3]) / 2, (b[1] + b[4]) / 2, (b[2] + b[5]) / 2};
    return dist_to_ray(o, d, c) <= thr;      // NaN (cos^2 > 1 by rounding) compares false, like numpy
}

// Which boxes a ray has to look at.  The prefilter keeps a box when its CENTRE lies within `thr` of the ray's (infinite) line
// (eval_kitti_render.py:237-244,367-369), so any superset of those boxes gives the same result -- the prefilter itself is
// still evaluated, exactly, on every box handed out.
//   AllBoxes: 0 .. K-1 (the reference's scan; boxes staged in shared memory when they fit).
//   GridBoxes: a uniform x/y grid over the box centres (cell size h, cell -> boxes in CSR form, built once per scene by the
//   caller).  The ray's projection is walked column by column along its major axis; per column the cells whose y (x) range
//   meets the projected line widened by thr sqrt(1 + m^2) are visited.  At the shipped scenes' K (15,333 / 5,729 boxes) this
//   cuts the boxes per ray by an order of magnitude.  Boxes arrive in cell order, not index order: the candidate records carry
//   their box index and the sort key of the group builder is (near, index), which is what a stable argsort of the
//   index-ordered scan (:440) produces.
struct BoxGrid {
    double x0, y0, h;
    int nx, ny;
    const int32_t* cell_start;      // [nx * ny + 1]
    const int32_t* cell_boxes;      // [K] box indices, cell by cell
};

struct AllBoxes {
    int K;
    template <class F>
    __device__ __forceinline__ void operator()(const double*, const double*, double, F f) const {
        for (int k = 0; k < K; ++k) f(k);
    }
};

struct GridBoxes {
    BoxGrid g;
    template <class F>
    __device__ void operator()(const double o[3], const double d[3], double thr, F f) const {
        const double dx = d[0], dy = d[1];
        if (!(isfinite(dx) && isfinite(dy) && isfinite(o[0]) && isfinite(o[1]))) return;    // the prefilter is false for NaN
        const double R = thr + 1e-6;
        auto cell = [&](int ix, int iy) {
            const int c = iy * g.nx + ix;
            for (int q = g.cell_start[c]; q < g.cell_start[c + 1]; ++q) f(g.cell_boxes[q]);
        };
        const double nrm = sqrt(dx * dx + dy * dy);
        if (nrm < 1e-9) {                          // (nearly) vertical ray: the cells within R of its foot point
            int i0 = (int)floor((o[0] - R - g.x0) / g.h), i1 = (int)floor((o[0] + R - g.x0) / g.h);
            int j0 = (int)floor((o[1] - R - g.y0) / g.h), j1 = (int)floor((o[1] + R - g.y0) / g.h);
            i0 = i0 < 0 ? 0 : i0; j0 = j0 < 0 ? 0 : j0;
            i1 = i1 >= g.nx ? g.nx - 1 : i1; j1 = j1 >= g.ny ? g.ny - 1 : j1;
            for (int j = j0; j <= j1; ++j)
                for (int i = i0; i <= i1; ++i) cell(i, j);
            return;
        }
        const bool xmaj = fabs(dx) >= fabs(dy);
        const double m = xmaj ? dy / dx : dx / dy;                 // |m| <= 1
        const double W = R * sqrt(1.0 + m * m) + 1e-9;
        const double oa = xmaj ? o[0] : o[1], ob = xmaj ? o[1] : o[0];
        const double a_org = xmaj ? g.x0 : g.y0, b_org = xmaj ? g.y0 : g.x0;
        const int na = xmaj ? g.nx : g.ny, nb = xmaj ? g.ny : g.nx;
        for (int i = 0; i < na; ++i) {
            const double a0 = a_org + i * g.h, a1 = a0 + g.h;
            const double v0 = ob + (a0 - oa) * m, v1 = ob + (a1 - oa) * m;
            const double lo = fmin(v0, v1) - W, hi = fmax(v0, v1) + W;
            if (hi < b_org || lo > b_org + nb * g.h) continue;
            int j0 = (int)floor((lo - b_org) / g.h), j1 = (int)floor((hi - b_org) / g.h);
            j0 = j0 < 0 ? 0 : j0;
            j1 = j1 >= nb ? nb - 1 : j1;
            for (int j = j0; j <= j1; ++j) {
                if (xmaj) cell(i, j); else cell(j, i);
            }
        }
    }
};

template <class Boxes, class Emit>
__device__ int enum_candidates(const double o[3], const double d[3], const double* boxes, const double* larger, Boxes each,
                               int method, double grow, double thr, double parent_far, Emit emit) {
    int n = 0, nfilt = 0;
    bool done = false;
    each(o, d, thr, [&](int k) {
        if (done || !prefilter_pass(o, d, boxes + 6 * k, thr)) return;
        nfilt++;
        double nr, fr;
        if (child_0429(o, d, larger + 6 * k, larger + 6 * k + 3, nr, fr)) {
            if (method == 1) { emit(n, 0.0, parent_far, k); n = 1; done = true; return; }
            emit(n, nr, fr, k);
            n++;
        }
    });
    if (done) return 1;
    if (n > 0 || nfilt == 0) return n;
    // grow-until-hit fallback (:399-437): extend_iter accumulates, boxes grow by the running total each round.
    int tstar = INT_MAX;
    each(o, d, thr, [&](int k) {
        if (!prefilter_pass(o, d, boxes + 6 * k, thr)) return;
        double lo[3] = {larger[6 * k], larger[6 * k + 1], larger[6 * k + 2]};
        double hi[3] = {larger[6 * k + 3], larger[6 * k + 4], larger[6 * k + 5]};
        double ext = 0;
        for (int t = 1; t < tstar; ++t) {
            if (ext > 0.5) break;
            ext = ext + grow;
            for (int a = 0; a < 3; ++a) { lo[a] = lo[a] - ext; hi[a] = hi[a] + ext; }
            double nr, fr;
            if (child_0429(o, d, lo, hi, nr, fr)) { tstar = t; break; }
        }
    });
    if (tstar == INT_MAX) return 0;
    each(o, d, thr, [&](int k) {
        if (done || !prefilter_pass(o, d, boxes + 6 * k, thr)) return;
        double lo[3] = {larger[6 * k], larger[6 * k + 1], larger[6 * k + 2]};
        double hi[3] = {larger[6 * k + 3], larger[6 * k + 4], larger[6 * k + 5]};
        double ext = 0;
        for (int t = 1; t <= tstar; ++t) {
            ext = ext + grow;
            for (int a = 0; a < 3; ++a) { lo[a] = lo[a] - ext; hi[a] = hi[a] + ext; }
        }
        double nr, fr;
        if (child_0429(o, d, lo, hi, nr, fr)) {
            if (method == 1) { emit(n, 0.0, parent_far, k); n = 1; done = true; return; }
            emit(n, nr, fr, k);
            n++;
        }
    });
    return done ? 1 : n;
}

template <class Boxes>
__global__ void k_groups_count(const double* __restrict__ ro, const double* __restrict__ rd, int64_t n,
                               const double* __restrict__ boxes, const double* __restrict__ larger, int K, Boxes each, Six pbox,
                               int method, double grow, double thr, int use_smem, int32_t* __restrict__ ocount,
                               double* __restrict__ opfar) {
    extern __shared__ double sm[];
    const double* bx = stage(boxes, sm, K * 6, use_smem);
    const double* lx = stage(larger, sm + (use_smem ? K * 6 : 0), K * 6, use_smem);
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double o[3], d[3];
        load3(ro, i, o); load3(rd, i, d);
        const double pf = slab_far(o, d, pbox.v);
        opfar[i] = pf;
        ocount[i] = enum_candidates(o, d, bx, lx, each, method, grow, thr, pf, [](int, double, double, int) {});
    }
}

// scratch: 3 doubles per candidate row (near, far, box index)
template <class Boxes>
__global__ void k_groups_fill(const double* __restrict__ ro, const double* __restrict__ rd,
                              const double* __restrict__ dist, int64_t n, const double* __restrict__ boxes,
                              const double* __restrict__ larger, int K, Boxes each, int method, double grow, double thr,
                              int use_smem, const int32_t* __restrict__ count, const int64_t* __restrict__ offset,
                              const double* __restrict__ pfar, double* __restrict__ scratch,
                              float* __restrict__ orays, float* __restrict__ oranges, int64_t* __restrict__ oother) {
    extern __shared__ double sm[];
    const double* bx = stage(boxes, sm, K * 6, use_smem);
    const double* lx = stage(larger, sm + (use_smem ? K * 6 : 0), K * 6, use_smem);
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int cnt = count[i];
        if (cnt == 0) continue;
        double o[3], d[3];
        load3(ro, i, o); load3(rd, i, d);
        const int64_t base = offset[i];
        double* sc = scratch + 3 * base;
        enum_candidates(o, d, bx, lx, each, method, grow, thr, pfar[i],
                        [&](int j, double nr, double fr, int k) { sc[3 * j] = nr; sc[3 * j + 1] = fr; sc[3 * j + 2] = (double)k; });
        // np.argsort(near) (:440) on the index-ordered scan = insertion sort on (near, box index)
        for (int a = 1; a < cnt; ++a) {
            const double kn = sc[3 * a], kf = sc[3 * a + 1], kk = sc[3 * a + 2];
            int b = a - 1;
            while (b >= 0 && (sc[3 * b] > kn || (sc[3 * b] == kn && sc[3 * b + 2] > kk))) {
                sc[3 * b + 3] = sc[3 * b]; sc[3 * b + 4] = sc[3 * b + 1]; sc[3 * b + 5] = sc[3 * b + 2];
                --b;
            }
            sc[3 * b + 3] = kn; sc[3 * b + 4] = kf; sc[3 * b + 5] = kk;
        }
        for (int j = 0; j < cnt; ++j) {
            float* r = orays + 13 * (base + j);
            r[0] = (float)o[0]; r[1] = (float)o[1]; r[2] = (float)o[2];
            r[3] = (float)d[0]; r[4] = (float)d[1]; r[5] = (float)d[2];
            r[6] = (float)sc[3 * j]; r[7] = (float)sc[3 * j + 1]; r[8] = 3.f;
            r[9] = 0.f; r[10] = (float)pfar[i]; r[11] = (float)(j + 1);
            r[12] = j == 0 ? (float)(cnt - 1) : -1.f;
            oranges[base + j] = (float)dist[i];
            oother[base + j] = j == 0 ? (cnt - 1) : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------

static int grid_for(int64_t n) {
    int64_t g = pcn_cdiv(n, AABB_THREADS);
    int64_t cap = (int64_t)PCN_SM_COUNT * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static const size_t kMaxBoxSmem = 200 * 1024;

template <class Kern>
static int prep_smem(Kern kern, size_t bytes, int* use_smem, size_t* smem) {
    *use_smem = bytes <= kMaxBoxSmem;
    *smem = *use_smem ? bytes : 0;
    if (*smem > 48 * 1024) PCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
    return 0;
}

extern "C" int pcnerf_aabb_far_bound(const double* ray_o, const double* ray_d, int64_t n, const double* h_parent,
                                     double* out_t, void* stream) {
    PCN_CHECK_ARG(n >= 0 && h_parent, "aabb_far_bound: bad arguments");
    if (n == 0) return 0;
    Six pl;
    for (int i = 0; i < 6; ++i) pl.v[i] = h_parent[i];
    PcnScope ps(PCN_K_AABB, (cudaStream_t)stream, (double)(n * 56.0));
    k_far_bound<<<grid_for(n), AABB_THREADS, 0, (cudaStream_t)stream>>>(ray_o, ray_d, n, pl, out_t);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_aabb_slab(const double* ray_o, const double* ray_d, int64_t n, const double* h_min3,
                                const double* h_max3, double* out_t, void* stream) {
    PCN_CHECK_ARG(n >= 0 && h_min3 && h_max3, "aabb_slab: bad arguments");
    if (n == 0) return 0;
    Six b;
    for (int i = 0; i < 3; ++i) { b.v[i] = h_min3[i]; b.v[3 + i] = h_max3[i]; }
    PcnScope ps(PCN_K_AABB, (cudaStream_t)stream, (double)(n * 56.0));
    k_slab<<<grid_for(n), AABB_THREADS, 0, (cudaStream_t)stream>>>(ray_o, ray_d, n, b, out_t);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_aabb_child_pairs(int variant, const double* ray_o, const double* ray_d, int64_t n,
                                       const double* boxes, int K, uint8_t* out_flag, double* out_near,
                                       double* out_far, void* stream) {
    PCN_CHECK_ARG(n >= 0 && K >= 0, "aabb_child_pairs: bad sizes");
    PCN_CHECK_ARG(variant == 406 || variant == 606 || variant == 429, "aabb_child_pairs: variant must be 406|606|429");
    if (n == 0 || K == 0) return 0;
    int use_smem; size_t smem;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_PAIRS(V)                                                                                   \
    {                                                                                                     \
        int rc = prep_smem(k_child_pairs<V>, (size_t)K * 48, &use_smem, &smem);                           \
        if (rc) return rc;                                                                                \
        k_child_pairs<V><<<grid_for(n), AABB_THREADS, smem, st>>>(ray_o, ray_d, n, boxes, K, use_smem,    \
                                                                   out_flag, out_near, out_far);          \
    }
    PcnScope ps(PCN_K_AABB, st, (double)n * (48.0 + 17.0 * K) + K * 48.0);
    if (variant == 429) LAUNCH_PAIRS(429) else if (variant == 606) LAUNCH_PAIRS(606) else LAUNCH_PAIRS(406)
#undef LAUNCH_PAIRS
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_aabb_dist_to_ray(const double* ray_o, const double* ray_d, int64_t n, const double* centres,
                                       int K, double* out, void* stream) {
    PCN_CHECK_ARG(n >= 0 && K >= 0, "aabb_dist_to_ray: bad sizes");
    if (n == 0 || K == 0) return 0;
    PcnScope ps(PCN_K_AABB, (cudaStream_t)stream, (double)(n * (48.0 + 8.0 * K)));
    k_dist_to_ray<<<grid_for(n), AABB_THREADS, 0, (cudaStream_t)stream>>>(ray_o, ray_d, n, centres, K, out);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_aabb_find_box(const double* centres, const double* boxes, int K, const double* points,
                                    int64_t q, int knn, int32_t* out_idx, void* stream) {
    PCN_CHECK_ARG(q >= 0 && K >= 1 && knn >= 1, "aabb_find_box: bad sizes");
    PCN_CHECK_ARG(knn <= K, "aabb_find_box: k must be less than or equal to the number of training points");
    if (q == 0) return 0;
    int use_smem; size_t smem;
    int rc = prep_smem(k_find_box, (size_t)K * 72, &use_smem, &smem);
    if (rc) return rc;
    PcnScope ps(PCN_K_AABB, (cudaStream_t)stream, (double)(q * 28.0 + K * 72.0));
    k_find_box<<<grid_for(q), AABB_THREADS, smem, (cudaStream_t)stream>>>(centres, boxes, K, points, q, knn, use_smem,
                                                                           out_idx);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_aabb_pack_train(int variant, const double* ray_o, const double* ray_d, const double* dist,
                                      const double* points, int64_t n, const double* centres, const double* boxes,
                                      const double* boxes_bigger, int K, const double* h_parent,
                                      double surface_expand, int knn, float* out_rays, uint8_t* out_keep,
                                      void* stream) {
    PCN_CHECK_ARG(n >= 0 && K >= 1 && h_parent, "aabb_pack_train: bad arguments");
    PCN_CHECK_ARG(variant == 406 || variant == 606, "aabb_pack_train: variant must be 406 (MaiCity) or 606 (KITTI)");
    PCN_CHECK_ARG(knn >= 1 && knn <= K, "aabb_pack_train: k must be less than or equal to the number of training points");
    if (n == 0) return 0;
    Six pl;
    for (int i = 0; i < 6; ++i) pl.v[i] = h_parent[i];
    int use_smem; size_t smem;
    cudaStream_t st = (cudaStream_t)stream;
    PcnScope ps(PCN_K_AABB, st, (double)n * (80.0 + 61.0) + K * 120.0);
    if (variant == 406) {
        int rc = prep_smem(k_pack_train<406>, (size_t)K * 72, &use_smem, &smem);
        if (rc) return rc;
        k_pack_train<406><<<grid_for(n), AABB_THREADS, smem, st>>>(ray_o, ray_d, dist, points, n, centres, boxes,
                                                                    boxes_bigger, K, pl, surface_expand, knn, use_smem,
                                                                    out_rays, out_keep);
    } else {
        int rc = prep_smem(k_pack_train<606>, (size_t)K * 72, &use_smem, &smem);
        if (rc) return rc;
        k_pack_train<606><<<grid_for(n), AABB_THREADS, smem, st>>>(ray_o, ray_d, dist, points, n, centres, boxes,
                                                                    boxes_bigger, K, pl, surface_expand, knn, use_smem,
                                                                    out_rays, out_keep);
    }
    PCN_LAUNCH_CHECK();
    return 0;
}

static bool grid_from(const double* h_grid5, const int32_t* cell_start, const int32_t* cell_boxes, BoxGrid* g) {
    if (!h_grid5 || !cell_start || !cell_boxes) return false;
    g->x0 = h_grid5[0]; g->y0 = h_grid5[1]; g->h = h_grid5[2];
    g->nx = (int)h_grid5[3]; g->ny = (int)h_grid5[4];
    g->cell_start = cell_start; g->cell_boxes = cell_boxes;
    return g->h > 0 && g->nx >= 1 && g->ny >= 1;
}

extern "C" int pcnerf_aabb_groups_count(const double* ray_o, const double* ray_d, int64_t n, const double* boxes,
                                        const double* boxes_larger, int K, const double* h_pmin3,
                                        const double* h_pmax3, int method, double grow_step, double prefilter,
                                        const double* h_grid5, const int32_t* cell_start, const int32_t* cell_boxes,
                                        int32_t* out_count, double* out_parent_far, void* stream) {
    PCN_CHECK_ARG(n >= 0 && K >= 0 && h_pmin3 && h_pmax3, "aabb_groups_count: bad arguments");
    PCN_CHECK_ARG(method == 1 || method == 2, "aabb_groups_count: depth_inference_method must be 1 or 2");
    if (n == 0) return 0;
    Six b;
    for (int i = 0; i < 3; ++i) { b.v[i] = h_pmin3[i]; b.v[3 + i] = h_pmax3[i]; }
    PcnScope ps(PCN_K_AABB, (cudaStream_t)stream, (double)(n * 60.0 + K * 96.0));
    BoxGrid g;
    if (grid_from(h_grid5, cell_start, cell_boxes, &g)) {
        GridBoxes each = {g};
        k_groups_count<GridBoxes><<<grid_for(n), AABB_THREADS, 0, (cudaStream_t)stream>>>(
            ray_o, ray_d, n, boxes, boxes_larger, K, each, b, method, grow_step, prefilter, 0, out_count, out_parent_far);
    } else {
        int use_smem; size_t smem;
        int rc = prep_smem(k_groups_count<AllBoxes>, (size_t)K * 96, &use_smem, &smem);
        if (rc) return rc;
        AllBoxes each = {K};
        k_groups_count<AllBoxes><<<grid_for(n), AABB_THREADS, smem, (cudaStream_t)stream>>>(
            ray_o, ray_d, n, boxes, boxes_larger, K, each, b, method, grow_step, prefilter, use_smem, out_count, out_parent_far);
    }
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_aabb_groups_fill(const double* ray_o, const double* ray_d, const double* dist, int64_t n,
                                       const double* boxes, const double* boxes_larger, int K, int method,
                                       double grow_step, double prefilter, const double* h_grid5,
                                       const int32_t* cell_start, const int32_t* cell_boxes, const int32_t* count,
                                       const int64_t* offset, const double* parent_far, double* scratch,
                                       float* out_rays13, float* out_ranges, int64_t* out_other, void* stream) {
    PCN_CHECK_ARG(n >= 0 && K >= 0, "aabb_groups_fill: bad sizes");
    PCN_CHECK_ARG(method == 1 || method == 2, "aabb_groups_fill: depth_inference_method must be 1 or 2");
    if (n == 0) return 0;
    PcnScope ps(PCN_K_AABB, (cudaStream_t)stream, (double)(n * 230.0 + K * 96.0));
    BoxGrid g;
    if (grid_from(h_grid5, cell_start, cell_boxes, &g)) {
        GridBoxes each = {g};
        k_groups_fill<GridBoxes><<<grid_for(n), AABB_THREADS, 0, (cudaStream_t)stream>>>(
            ray_o, ray_d, dist, n, boxes, boxes_larger, K, each, method, grow_step, prefilter, 0, count, offset, parent_far,
            scratch, out_rays13, out_ranges, out_other);
    } else {
        int use_smem; size_t smem;
        int rc = prep_smem(k_groups_fill<AllBoxes>, (size_t)K * 96, &use_smem, &smem);
        if (rc) return rc;
        AllBoxes each = {K};
        k_groups_fill<AllBoxes><<<grid_for(n), AABB_THREADS, smem, (cudaStream_t)stream>>>(
            ray_o, ray_d, dist, n, boxes, boxes_larger, K, each, method, grow_step, prefilter, use_smem, count, offset,
            parent_far, scratch, out_rays13, out_ranges, out_other);
    }
    PCN_LAUNCH_CHECK();
    return 0;
}
