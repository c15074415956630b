/*
 * pcnerf_b200 -- C ABI of the B200-native PC-NeRF ray-rendering hot path (libpcnerf_b200.so).
 *
 * The reference (biter0088/pc-nerf) has no FFI: its hot path is plain Python/PyTorch.  Each entry point below
 * replaces the eager-torch / numpy body of the reference function cited next to it (file:line relative to the
 * reference tree) and is what a binding for that function would call (see INTEGRATION.md for the ctypes stubs).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `h_` (host, read during the call);
 *   - tensors are dense row-major; `ld` parameters are row strides in elements;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point synchronises the device or
 *     allocates device memory.  Process-wide state is limited to: the instrumentation counters, the eval-engine switch
 *     (pcnerf_tc_set_fused_eval), the constant bank holding the epilogue constants of the fused eval MLP (re-loaded at the
 *     first chunk of every call) and the two internal streams of pcnerf_mlp_tc_*_chunks (forked from / joined to
 *     `stream`).  The reference drives the path from one host thread (the Lightning main loop); the precision-1 MLP entry
 *     points assume the same: one host thread, one stream at a time per process.  Everything else is re-entrant per stream;
 *   - return value: 0 on success, negative on error; pcnerf_last_error() gives the message (thread local).
 *   - fp32 "renderer" arithmetic follows the reference's operation order where masks/indices depend on it
 *     (no FMA contraction in z placement and thresholds); the AABB stage is fp64 like the reference's numpy.
 */
#ifndef PCNERF_B200_H
#define PCNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCNERF_OK 0
#define PCNERF_ERR_ARG (-1)
#define PCNERF_ERR_CUDA (-2)
#define PCNERF_ERR_UNSUPPORTED (-3)

#define PCNERF_ENC_DIM 63      /* 3 + 3*10*2, nof/networks/models.py:27-41 */
#define PCNERF_ENC_LD 64       /* encodings are stored padded to 64 columns (col 63 == 0) */
#define PCNERF_HID 256         /* feature_size */
#define PCNERF_NLIN 9          /* 8 hidden Linear + occ_out */
#define PCNERF_NBN 8

const char* pcnerf_last_error(void);
int pcnerf_version(void);

/* Instrumentation (no reference counterpart; used by bench.py for `gpu_launches` and the roofline figures).
 * pcnerf_launch_count: kernels launched by this library since the last reset (always counted).
 * pcnerf_prof_enable(1): from now on every launch is bracketed by CUDA events on its own stream, grouped in
 * pcnerf_prof_classes() kernel classes; pcnerf_prof_read synchronises those events and returns the summed device
 * time (ms), launches and algorithmic work (FLOPs for the MLP GEMM classes, bytes for the others) of one class. */
long long pcnerf_launch_count(int reset);
void pcnerf_prof_enable(int on);
int pcnerf_prof_classes(void);
const char* pcnerf_prof_name(int id);
int pcnerf_prof_read(int id, double* out_ms, long long* out_launches, double* out_work);

/* ------------------------------------------------------------------------------------------------------------
 * K1  AABB stage (fp64).  Replaces the numpy / python-scalar leaf functions and the per-ray loops around them.
 * ---------------------------------------------------------------------------------------------------------- */

/* compute_far_bound, nof/dataset/ipb2dmapping.py:36-77.  h_parent = {x_max,x_min,y_max,y_min,z_max,z_min}.
 * out_t[i] = min positive plane parameter, +inf where the reference returns None. */
int pcnerf_aabb_far_bound(const double* ray_o, const double* ray_d, int64_t n, const double* h_parent,
                          double* out_t, void* stream);

/* ray_aabb_distances, eval_kitti_render.py:213-235 (slab test; +inf when tmax < tmin). */
int pcnerf_aabb_slab(const double* ray_o, const double* ray_d, int64_t n, const double* h_min3,
                     const double* h_max3, double* out_t, void* stream);

/* compute_far_bound0406 / 0606 / 0429 for every (ray, box) pair: ipb2dmapping.py:82-114, :119-172,
 * eval_kitti_render.py:170-211.  variant = 406 | 606 | 429.  boxes (K,6) = min xyz, max xyz.
 * out_flag (n,K) u8, out_near/out_far (n,K) f64 (NaN for 0406 with fewer than two hits). */
int pcnerf_aabb_child_pairs(int variant, const double* ray_o, const double* ray_d, int64_t n, const double* boxes,
                            int K, uint8_t* out_flag, double* out_near, double* out_far, void* stream);

/* distance_to_ray, eval_kitti_render.py:237-244, for every (ray, centre) pair -> out (n,K) f64. */
int pcnerf_aabb_dist_to_ray(const double* ray_o, const double* ray_d, int64_t n, const double* centres, int K,
                            double* out, void* stream);

/* find_aabb_box, ipb2dmapping.py:174-197, batched: first containing box among the knn nearest centres.
 * out_idx[q] = box index or -1. */
int pcnerf_aabb_find_box(const double* centres, const double* boxes, int K, const double* points, int64_t q,
                         int knn, int32_t* out_idx, void* stream);

/* Training ray packing: loop body of ipb2dmapping.py:367-397 (variant 406, MaiCity) / :736-768 (variant 606, KITTI)
 * + the 15-column record of :447-452.  out_rays (n,15) f32, out_keep (n) u8 (rows with keep==0 are dropped by the
 * caller, order preserved). */
int pcnerf_aabb_pack_train(int variant, const double* ray_o, const double* ray_d, const double* dist,
                           const double* points, int64_t n, const double* centres, const double* boxes,
                           const double* boxes_bigger, int K, const double* h_parent, double surface_expand,
                           int knn, float* out_rays, uint8_t* out_keep, void* stream);

/* Candidate-group builder: loop body of eval_kitti_render.py:353-461 (grow_step 0.005) / :681-803 (0.05).
 * Pass 1 counts candidate rows per ray (0 = ray dropped) and emits the parent far bound;
 * pass 2 (after an exclusive prefix sum of the counts done by the caller) fills the (N',13) rows, ranges and the
 * `other_interest_sub_nerf_number` side array, candidates sorted by ascending near (ties: ascending box index = the stable
 * argsort of :440 on the index-ordered scan).
 * Optional uniform grid over the box CENTRES (all three NULL = scan every box, as the reference does): h_grid5 (HOST) =
 * {x0, y0, cell size h, nx, ny}; cell_start (nx*ny + 1) int32 and cell_boxes (K) int32 = box indices cell by cell (cell =
 * iy*nx + ix of the box centre).  The prefilter keeps boxes whose centre lies within `prefilter` of the ray's line, so only
 * the cells along the projected line are visited (an order of magnitude fewer boxes per ray at the shipped scenes' K =
 * 15,333 / 5,729); results are identical to the full scan. */
int pcnerf_aabb_groups_count(const double* ray_o, const double* ray_d, int64_t n, const double* boxes,
                             const double* boxes_larger, int K, const double* h_pmin3, const double* h_pmax3,
                             int method, double grow_step, double prefilter, const double* h_grid5,
                             const int32_t* cell_start, const int32_t* cell_boxes, int32_t* out_count,
                             double* out_parent_far, void* stream);
int pcnerf_aabb_groups_fill(const double* ray_o, const double* ray_d, const double* dist, int64_t n,
                            const double* boxes, const double* boxes_larger, int K, int method, double grow_step,
                            double prefilter, const double* h_grid5, const int32_t* cell_start,
                            const int32_t* cell_boxes, const int32_t* count, const int64_t* offset,
                            const double* parent_far, double* scratch /* (N',3) f64 */, float* out_rays13,
                            float* out_ranges, int64_t* out_other, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K2  sample placement + positional encoding (fp32).
 * ---------------------------------------------------------------------------------------------------------- */

/* nof/render.py:429-458 (+ :565-570 disparity, :622-628 view variant) fused with Embedding.forward
 * (nof/networks/models.py:27-41).
 *   rays (n, ld) f32; near/far read from columns near_col/far_col; child near/far from cnear_col/cfar_col.
 *   steps_a (n_a) = torch.linspace(0,1,n_a) for the parent segment, steps_b (n_b) for the child segment
 *   (n_b == 0 -> uniform sampling).  S = n_a + n_b.
 *   U (n,S) = pre-drawn uniform numbers or NULL (perturb == 0).
 *   out_z (n,S); out_enc (n*S, 64) f32 or NULL; out_enc_f16 (n*S,64) fp16 or NULL (the tensor-core MLP's operand). */
int pcnerf_sample_encode_coarse(const float* rays, int ld, int64_t n, int near_col, int far_col, int cnear_col,
                                int cfar_col, const float* steps_a, int n_a, const float* steps_b, int n_b,
                                int use_disp, float perturb, const float* U, float* out_z, float* out_enc,
                                void* out_enc_f16, void* stream);

/* sample_pdf + merge (nof/render.py:371-412, :463-468) fused with the encoding of the merged samples.
 *   z (n,S), w (n,S) coarse depths / weights; u: (Ni) shared (u_ld == 0, det) or (n,Ni) (u_ld == Ni).
 *   out_z (n, S+Ni) ascending. */
int pcnerf_sample_encode_fine(const float* rays, int ld, int64_t n, const float* z, const float* w, int S,
                              const float* u, int u_ld, int Ni, float* out_z, float* out_enc, void* out_enc_f16,
                              void* stream);

/* sample_pdf alone (nof/render.py:371-412): bins (n,nb), weights (n,nb-1), u as above -> out (n,Ni), unsorted. */
int pcnerf_sample_pdf(const float* bins, const float* weights, int64_t n, int nb, const float* u, int u_ld, int Ni,
                      float* out, void* stream);

/* Embedding.forward alone (nof/networks/models.py:27-41): x (b,3) -> out (b, out_ld) f32, 63 columns written
 * (+ zero padding up to out_ld). */
int pcnerf_embed(const float* x, int64_t b, float* out, int out_ld, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K3  occupancy MLP (NOF / NOF_coarse / NOF_fine / NOF_plusfine, nof/networks/models.py:44-359).
 * As constructed by the reference every activation is LeakyReLU(negative_slope=1.0) == identity
 * (models.py:152,172), so the network is Linear->BN x8 -> Linear -> Sigmoid; BN(l) is folded into Linear(l+1).
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct pcnerf_mlp_params {
    const float* W[PCNERF_NLIN];        /* Linear weights (out,in): 63|256|256|256|319|256|256|256 -> 256, 256 -> 1 */
    const float* b[PCNERF_NLIN];
    const float* gamma[PCNERF_NBN];
    const float* beta[PCNERF_NBN];
    float* running_mean[PCNERF_NBN];    /* updated in place in training mode (momentum, unbiased var) */
    float* running_var[PCNERF_NBN];
    int64_t* num_batches_tracked[PCNERF_NBN];
    float momentum;                     /* 0.1 */
    float eps;                          /* 1e-5 */
    int training;                       /* 1: batch statistics of this chunk; 0: running statistics */
    int precision;                      /* 0: fp32 CUDA-core GEMM (1e-5 gate); 1: tcgen05 GEMM, fp16 operands forward /
                                           bf16 gradients, fp32 accumulation (1e-3 gate) */
    int prepared;                       /* 1: `scratch` still holds the padded / transposed weight copies written by the
                                           previous call with these same weights (next chunk of the same pass): skip
                                           re-deriving them.  0 is always safe.  2 (eval mode, precision 1 only): the
                                           copies in `scratch` are valid but this is the first chunk of a new call (the
                                           per-call constants of the fused eval kernel are re-loaded). */
} pcnerf_mlp_params;

typedef struct pcnerf_mlp_grads {       /* accumulated (+=) by pcnerf_mlp_backward */
    float* dW[PCNERF_NLIN];
    float* db[PCNERF_NLIN];
    float* dgamma[PCNERF_NBN];
    float* dbeta[PCNERF_NBN];
} pcnerf_mlp_grads;

/* Bytes of the activation store for one chunk of `rows` samples (saved between forward and backward) and of
 * the scratch area (reusable across chunks). */
size_t pcnerf_mlp_saved_bytes(int64_t rows, int precision);
size_t pcnerf_mlp_scratch_bytes(int64_t rows, int precision);

/* One BN batch (= one `chunk` of nof/render.py:47-49).  enc (rows,64) f32 (precision 0) or fp16 (precision 1).
 * out_p (rows) = sigmoid(logit).  `saved` receives the pre-BN activations and the batch statistics; it is not touched
 * (and may be NULL) for precision 1 with training == 0 while pcnerf_tc_get_fused_eval() is 1. */
int pcnerf_mlp_forward(const pcnerf_mlp_params* h_params, const void* enc, int64_t rows, float* out_p,
                       void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes, void* stream);

/* Backward of the same chunk.  grad_p (rows) = dL/dp.  Needs out_p and `saved` of the forward call.
 * Destroys `saved`. */
int pcnerf_mlp_backward(const pcnerf_mlp_params* h_params, const pcnerf_mlp_grads* h_grads, const void* enc,
                        int64_t rows, const float* out_p, const float* grad_p, void* saved, size_t saved_bytes,
                        void* scratch, size_t scratch_bytes, void* stream);

/* K3' closed form ("affine" mode): for one BatchNorm batch the logit of the network AS THE REFERENCE BUILDS IT (identity
 * activations, models.py:152,172) is exactly alpha . x + c, with (alpha, c) a function of the parameters and of the
 * batch's first two moments.  These are the three data-sized kernels; the parameter-sized algebra lives in the host
 * mirror.  enc (rows,64) f32; chunks are consecutive ranges of `chunk` rows (nchunk = ceil(rows/chunk)).
 * pcnerf_affine_moments: out_part [nchunk][P][65][64] f64, P = pcnerf_affine_parts(): partial sums over the chunk's rows
 *   of (x-s)(x-s)^T (rows 0..63) and (x-s) (row 64), s = the chunk's first row.
 * pcnerf_affine_apply: out_p[r] = sigmoid(alpha[chunk(r)] . x_r + c[chunk(r)]); alpha (nchunk,64), c (nchunk) f32.
 * pcnerf_affine_grad: out_part [nchunk][P][65] f64: partial sums of g_r x_r (0..63) and g_r (64), g = grad_p * p (1-p). */
int pcnerf_affine_parts(void);
int pcnerf_affine_moments(const float* enc, int64_t rows, int64_t chunk, double* out_part, void* stream);
int pcnerf_affine_apply(const float* enc, int64_t rows, int64_t chunk, const float* alpha, const float* c, float* out_p,
                        void* stream);
int pcnerf_affine_grad(const float* enc, const float* p, const float* grad_p, int64_t rows, int64_t chunk,
                       double* out_part, void* stream);

/* K3' closed form on RAYS, training mode (csrc/affine_rays.cu): the whole precision-2 MLP of one pass -- batch moments,
 * the parameter-sized float64 algebra (forward and hand-derived backward, repo kernels) and the per-row dot + sigmoid -- from
 * (rays, z) directly: row r = (ray r / S, depth z[r]) and its encoding embed(o + d z) (models.py:27-41, nof/render.py:458)
 * is re-derived inside the kernels that need it, so no (rows,64) encoding tensor exists.  Replaces, for a training pass,
 * NOF.forward over every chunk (nof/render.py:47-49 -> models.py:183-203) and its autograd backward.
 * rays (n_rays, ld) f32 with the origin in columns 0..2 and the direction in 3..5; z (n_rays, S) f32; chunks are
 * consecutive ranges of `chunk` rows of the (n_rays S) samples, one BatchNorm batch each; running statistics are updated
 * once per chunk in chunk order.  `work` (pcnerf_affine_work_bytes(nchunk) bytes) carries the forward's intermediate
 * matrices to the backward call of the same pass.  h_params->training must be 1 (PCNERF_ERR_UNSUPPORTED otherwise).
 * pcnerf_affine_backward_rays ACCUMULATES (+=) into h_grads like pcnerf_mlp_backward. */
size_t pcnerf_affine_work_bytes(int64_t nchunk);
int pcnerf_affine_forward_rays(const pcnerf_mlp_params* h_params, const float* rays, int ld, int64_t n_rays, const float* z,
                               int S, int64_t chunk, float* out_p, void* work, size_t work_bytes, void* stream);
int pcnerf_affine_backward_rays(const pcnerf_mlp_params* h_params, const pcnerf_mlp_grads* h_grads, const float* rays,
                                int ld, int64_t n_rays, const float* z, int S, int64_t chunk, const float* out_p,
                                const float* grad_p, void* work, size_t work_bytes, void* stream);

/* K3' closed form on RAYS, eval mode (model.eval(): BatchNorm1d on its running statistics, models.py:183-203 as called from
 * nof/render.py:47-49 by the depth-inference path render.py:614-699): the logit is alpha . (x, 1) with alpha a function of the
 * parameters and running statistics alone.
 * pcnerf_affine_eval_alpha: alpha (64 floats, device; element 63 = the constant term) from h_params in float64 repo kernels;
 *   work: pcnerf_affine_work_bytes(1) bytes of scratch.  Callers cache alpha until a parameter or running statistic changes.
 * pcnerf_affine_apply_rays: out_p[r] = sigmoid(alpha . (embed(o + d z[r]), 1)) for the n_rays * S (ray, depth) rows
 *   (rays (n_rays, ld) f32: origin in columns 0..2, direction in 3..5; z (n_rays, S) f32); no encoding tensor is read. */
int pcnerf_affine_eval_alpha(const pcnerf_mlp_params* h_params, float* alpha, void* work, size_t work_bytes, void* stream);
int pcnerf_affine_apply_rays(const float* rays, int ld, int64_t n_rays, const float* z, int S, const float* alpha,
                             float* out_p, void* stream);

/* Building blocks of the precision-1 path (TMA + tcgen05 + TMEM), exported for unit tests and reuse.
 * pcnerf_tc_rowgemm: C[rows,256] = [A0 | A1][rows, k0+k1] * B[256, k0+k1]^T, k0, k1 multiples of 64, k0 + k1 <= 320.
 *   mode 0 (forward): A, B fp16; vec = bias[256]; out = fp16(C + bias); out2 reserved (NULL);
 *     stats (2,256) f64 = column sums of out and out^2 (the fp16 values; zeroed by the call).
 *   mode 1 (data gradient with the BatchNorm backward fused): A, B bf16; E (rows,256) fp16; vec = c0|c1|c2|mean [4][256];
 *     out = bf16(c0*C - c1 - (E - mean)*c2); stats[0] = column sums of out.
 * pcnerf_tc_wgrad: out[256, ldo] window [col_off, col_off+ncols) += DH[rows,256]^T (bf16) * X[rows, 0:ncols] (bf16, or
 *   fp16 converted to bf16 tile by tile in shared memory; row stride ldx); ncols 64 or 256; accumulated with atomics
 *   (zero `out` first).
 * pcnerf_tc_last_fault: non-zero if a tensor-core kernel aborted on a pipeline time-out (diagnostic).
 * work: pcnerf_tc_rowgemm_work_bytes() of device scratch (CTA counter + per-CTA statistic partials). */
size_t pcnerf_tc_rowgemm_work_bytes(void);
int pcnerf_tc_rowgemm(int mode, const void* A0, int k0, const void* A1, int k1, const void* B, const float* vec,
                      const void* E, int64_t rows, void* out, void* out2, double* stats, void* work, void* stream);
int pcnerf_tc_wgrad(const void* DH, const void* X, int ldx, int ncols, int x_is_bf16, int64_t rows, float* out, int ldo,
                    int col_off, void* stream);
int pcnerf_tc_last_fault(void);

/* All chunks (BN batches) of one pass of the precision-1 MLP in training mode, `lanes` (1 .. 4) chunks in flight on
 * internal streams forked from / joined to `stream` (capturable in a CUDA graph).  Chunk c covers rows [c*chunk, ...) of
 * enc (rows,64) fp16; h_saved[c] holds h_saved_bytes[c] >= pcnerf_mlp_saved_bytes(rows of chunk c, 1) bytes; h_scratch[k],
 * k < lanes, holds scratch_bytes >= pcnerf_mlp_scratch_bytes(min(chunk, rows), 1) bytes (h_*: host arrays of device
 * pointers / sizes).  Every chunk is validated like a pcnerf_mlp_forward call BEFORE anything is launched: a training-mode
 * chunk of exactly one row fails with torch's "Expected more than 1 value per channel" message (ValueError in the host
 * mirror), oversized chunks and undersized buffers with PCNERF_ERR_ARG.  Results are bit-identical to calling
 * pcnerf_mlp_forward / pcnerf_mlp_backward chunk after chunk (nof/render.py:47-49): running statistics and gradient sums
 * are updated in chunk order. */
int pcnerf_mlp_tc_forward_chunks(const pcnerf_mlp_params* h_params, const void* enc, int64_t rows, int64_t chunk,
                                 float* out_p, void* const* h_saved, const size_t* h_saved_bytes, void* const* h_scratch,
                                 size_t scratch_bytes, int lanes, void* stream);
int pcnerf_mlp_tc_backward_chunks(const pcnerf_mlp_params* h_params, const pcnerf_mlp_grads* h_grads, const void* enc,
                                  int64_t rows, int64_t chunk, const float* out_p, const float* grad_p,
                                  void* const* h_saved, const size_t* h_saved_bytes, void* const* h_scratch,
                                  size_t scratch_bytes, int lanes, void* stream);

/* Eval-mode (running-statistics) forward of the precision-1 MLP (replaces the per-chunk loop of nof/render.py:21-24 for
 * model.eval()).  Process-wide switch:
 *   2 (default) = all nine layers in one persistent kernel, activations resident in shared memory / TMEM, CTA pairs
 *       (clusters of two, tcgen05.mma.cta_group::2 with M = 256: each SM streams half of every weight block);
 *   1 = the same kernel with one CTA per unit of work (cta_group::1);
 *   0 = the layered row GEMMs (one kernel per Linear, activations through HBM). */
void pcnerf_tc_set_fused_eval(int on);
int pcnerf_tc_get_fused_eval(void);

/* Training-mode GEMMs of the precision-1 MLP (forward, data gradient, weight gradient of every Linear; also
 * pcnerf_tc_rowgemm / pcnerf_tc_wgrad).  Process-wide switch, a bit mask: bit 0 = forward, bit 1 = data gradient, bit 2 =
 * weight gradient (N = 256) on clusters of two CTAs (tcgen05.mma.cta_group::2, M = 256):
 *   forward / data gradient: the pair splits the ROWS of a two-tile unit, every A tile is loaded once, each SM keeps half of
 *     the weight rows resident; the data-gradient form (DGRAD2) also takes the BN-backward term -c2 (.) H_{l-1} on the tensor
 *     core (a block-diagonal fp16 operand) with `a` folded into the bf16 weight operand per chunk;
 *   weight gradient: one 256 x 256 accumulator per pair, half the split-K atomics.
 * A clear bit = the single-CTA kernels of round 1 (column-split row GEMMs with N = 128 MMAs, k_tc_wgrad).  Same results up
 * to fp32 summation order (and, for the data gradient, the rounding of a W instead of W to bf16).  Default 7; initial value:
 * environment variable PCNERF_TC_PAIRS. */
void pcnerf_tc_set_row_pairs(int on);
int pcnerf_tc_get_row_pairs(void);

/* Layered precision-1 forward (training mode, and eval mode with the fused kernel switched off): linear correction of
 * the fp16 rounding of the folded weights W' = W_{l+1} diag(a_l) (nof/networks/models.py:183-203 evaluated layer by
 * layer).  As the reference builds the network every activation is the identity (models.py:152,172), so the input of
 * every layer is an affine function of the 64-d encoding x and the term the rounding loses, (W' - fp16(W')) H_l, is
 * restored exactly by one extra K = 64 block [C | fp16(W')] x [x | H_l], C = (W' - fp16(W')) T_l (csrc/mlp_tc.cu,
 * k_tc_fold); the layer-0 and skip-connection weights enter as hi + lo fp16 pairs.  1 (default) = on, 0 = the plain
 * fp16(W') operands.  Initial value: environment variable PCNERF_TC_CORRECT. */
void pcnerf_tc_set_weight_correction(int on);
int pcnerf_tc_get_weight_correction(void);

/* ------------------------------------------------------------------------------------------------------------
 * K4  compositing + losses (nof/render.py:51-61, :75-161, :13-36, :166-226; train_kitti.py:145-146).
 * ---------------------------------------------------------------------------------------------------------- */

#define PCNERF_COMP_CHILD_LOSS 1        /* masks + child free / depth losses (use_child_nerf_loss == 1) */
#define PCNERF_COMP_OPACITY 2           /* opacity regulariser partial sums (render.py:224) */
#define PCNERF_COMP_RANGE_LOSS 4        /* scene-level range loss term sum_r SmoothL1(10 depth_r, 10 rays[r, range_col]) -> sums[3]
                                           (train_kitti.py:145-146 with loss_type smoothl1), fused into the same pass */

/* p, z (n,P).  rays (n,ld): child near/far in columns cnear_col/cfar_col, range reading in range_col.
 * noise (n,P) or NULL is added as noise*noise_std before normalisation.
 * Outputs: w (n,P); depth (n); per_ray (n,8) f32 = {free_r, dhat_r, sl1_child_r, C_r, lo0, hi0, lo2, hi2};
 * sums (5) f64 = {sum free_r, sum sl1_child_r, sum opacity terms, sum SmoothL1(10 depth_r, 10 range_r), arrival counter of
 * the finaliser} (zeroed by the call); out3 (3) f32 or NULL: the loss scalars of pcnerf_composite_losses, written by the
 * last block of the forward kernel to finish (no separate launch).
 * P = 64 / 128 / 192 / 384 with 16-byte aligned p / z / w / noise / per_ray take the register-resident kernels; any other P
 * or alignment takes the generic kernels. Both
 * forms evaluate the same formulas; products and sums are associated differently (ulp-level differences). */
int pcnerf_composite_fwd(const float* p, const float* z, const float* rays, int ld, int64_t n, int P,
                         int cnear_col, int cfar_col, int range_col, const float* noise, float noise_std,
                         float epsilon, int flags, float* w, float* depth, float* per_ray, double* sums,
                         float* out3, void* stream);

/* child_free_loss = sums[0]/n; child_depth_loss = (1/n)*0.1*(sums[1]/n)  (render.py:121,155); range term =
 * sums[3]/n = SmoothL1Loss(mean)(10 depth, 10 gt) (train_kitti.py:145-146 before the 0.1 * lambda_loss factor) -> out3 (3) f32 */
int pcnerf_composite_losses(const double* sums, int64_t n, float* out3, void* stream);

/* Backward.  Upstream gradients (any may be NULL): g_depth (n) for depth; g_free / g_dloss device scalars for the two
 * losses of pcnerf_composite_losses (n_total = number of rays their means are taken over); g_free_r / g_sl1_r (n) for
 * per_ray[:,0] / per_ray[:,2] (used by the use_child_nerf_divide == 1 segmented variant, render.py:106-119,135-152);
 * g_range device scalar for the range term of pcnerf_composite_losses (flags & PCNERF_COMP_RANGE_LOSS; needs `depth`
 * (n), the forward's output).  out grad_p (n,P) = dL/dp. */
int pcnerf_composite_bwd(const float* p, const float* z, const float* w, const float* rays, int ld, int64_t n,
                         int P, int range_col, float noise_std, float epsilon, int flags, const float* per_ray,
                         const float* g_depth, const float* g_free, const float* g_dloss, const float* g_free_r,
                         const float* g_sl1_r, int64_t n_total, const float* depth, const float* g_range,
                         float* grad_p, void* stream);

/* Masked-mean losses / metrics on (n) vectors of rendered depths: nof/criteria/loss.py:7-50 (NOFSmoothL1Loss, NOFMSELoss,
 * NOFL1Loss behind `nof_loss`; the scene-level range loss of train_kitti.py:145-146) and nof/criteria/metrics.py:5-21.
 * kind: 0 SmoothL1 (beta 1), 1 MSE, 2 L1, 3 abs_error, 4 acc_thres (percent of |pred - target| < 0.2).  mask (n) uint8 or
 * NULL (all elements).  acc2 (2) f64 scratch: sum of the elementwise terms, number of selected elements (kept for the
 * backward).  out (1) f32 = the mean (NaN for an empty selection, like torch).  Backward (kinds 0..2): g_pred / g_target
 * (n) = g_out[0] * d(mean)/d(pred | target); either may be NULL. */
int pcnerf_masked_loss_fwd(int kind, const float* pred, const float* target, const uint8_t* mask, int64_t n, double* acc2,
                           float* out, void* stream);
int pcnerf_masked_loss_bwd(int kind, const float* pred, const float* target, const uint8_t* mask, int64_t n,
                           const double* acc2, const float* g_out, float* g_pred, float* g_target, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K5  two-step depth-inference search (nof/render.py:229-368, :674-684).
 * ---------------------------------------------------------------------------------------------------------- */

/* Per candidate row: normalised weights, strict child mask with expand-until-non-empty (:252-263), fp64 Gaussian
 * smoothing sigma=5 'reflect' + first-max argmax (:303-308), peak-in-child (:309-313), in-child weight sum
 * (:314-315), depth by method 1 (:342-343) or 2 (:345-348), opacity partial sum -> sums[0] (f64, zeroed). */
int pcnerf_search_rows(const float* p, const float* z, const float* rays, int ld, int64_t n, int P,
                       int cnear_col, int cfar_col, float epsilon, int method, float* w, float* depth,
                       uint8_t* peak_in_child, float* wsum_child, double* sums, void* stream);

/* The same per-row search with the samples evaluated ONCE PER PHYSICAL RAY: all candidate rows of a group share origin,
 * direction and the parent segment (:619-628; eval_kitti_render.py:379-390), so z, the occupancies and the weights of a
 * group's rows are identical -- the reference recomputes them for every row.  p_ray / z_ray / w_ray (G,P) are indexed by
 * physical ray, row r of `rays` (n_rows, ld) belongs to ray row_ray[r] (non-decreasing); only the child interval
 * (cnear_col, cfar_col) and therefore masks, peak test, in-child sum and depth are per row.  w_ray may be NULL. */
int pcnerf_search_rows_grouped(const float* p_ray, const float* z_ray, const float* rays, int ld, int64_t n_rows, int P,
                               int cnear_col, int cfar_col, float epsilon, int method, const int32_t* row_ray,
                               float* w_ray, float* depth, uint8_t* peak_in_child, float* wsum_child, double* sums,
                               void* stream);

/* Group winner selection (:317-340).  other (n) i64: head = followers, followers = 0.  out_flag (n) u8.
 * n_rendered (device scalar, may be NULL): rows at or beyond *n_rendered never win (pcnerf_eval_rows_rendered). */
int pcnerf_search_select(const int64_t* other, const uint8_t* peak_in_child, const float* wsum_child, int64_t n,
                         const int64_t* n_rendered, uint8_t* out_flag, void* stream);

/* Group structure of the candidate rows, on the device: head_flag[i] = 1 where row i starts a group (the sequential walk
 * of :317-340 / eval_kitti_render.py:449-450: a head carries its follower count, followers carry 0).
 * pcnerf_group_uniform: *mismatch = number of rows whose ray (columns 0..5, pnear_col, pfar_col) differs bit-wise from the
 * ray of their group's head row head_row[row_ray[i]] (0 for rows built by the reference's frame builders). */
int pcnerf_group_heads(const int64_t* other, int64_t n, uint8_t* head_flag, void* stream);
int pcnerf_group_uniform(const float* rays, int ld, int64_t n, const int32_t* row_ray, const int64_t* head_row,
                         int pnear_col, int pfar_col, int* mismatch, void* stream);

/* The group-aligned batch walk of eval_kitti_render.py:979-1005 / :1111-1136 (tag column: followers carry -1) on the
 * device: *out_n = number of leading rows the reference's loop hands to the renderer -- n, or n - 1 when a batch ends one
 * row before the end (`if i == n - 1: break`).  Batch boundaries change no value in eval mode (running statistics). */
int pcnerf_eval_rows_rendered(const float* rays, int ld, int tag_col, int64_t n, int64_t batch_size_set, int64_t* out_n,
                              void* stream);

/* points = o + depth*d (:674-684). */
int pcnerf_points(const float* rays, int ld, int64_t n, const float* depth, float* out_xyz, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K6  point-cloud metrics (nof/criteria/pointcloud_metrics.py:5-49, nof/criteria/metrics.py:24-32).
 * ---------------------------------------------------------------------------------------------------------- */

/* nn_correspondance (pointcloud_metrics.py:5-33): for each vertex of verts2 (n2,3) f64 the exact nearest vertex of verts1
 * (n1,3) f64 -> out_idx (n2) i32, out_dist (n2) f64 = sqrt of the squared L2 distance.  n1 == 0 or n2 == 0: no output. */
int pcnerf_nn_correspondance(const double* verts1, int64_t n1, const double* verts2, int64_t n2, int32_t* out_idx,
                             double* out_dist, void* stream);
/* sums2 = {sum dist, count(dist < threshold)} (the ingredients of eval_pts, pointcloud_metrics.py:39-49). */
int pcnerf_dist_stats(const double* dist, int64_t n, double threshold, double* sums2, void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * K0  LiDAR frame -> returns of the block.  Replaces the numpy / python filtering of
 * nof/dataset/ipb2dmapping.py:662-711: near-sensor box |x|>=rdx or |y|>=rdy or |z|>=rdz, range <= max_range, height window
 * (float32, sensor frame), pose transform in float64 (h_pose16: HOST 4x4 row-major, float32 values), interest region
 * around any of the `npose` run poses (pose_xy (npose,2) float32, device; npose = 0: no such test), optional closed
 * parent-box test on the transformed points (h_box6: HOST x_min,x_max,y_min,y_max,z_min,z_max or NULL; the MaiCity loader,
 * :334-336, which passes +-inf for the height window), ray direction / range from the sensor position h_pos3 (HOST,
 * NULL = pose[:3,3]; the MaiCity loader transforms with the float32 pose but measures from the float64 position).
 * pts (n,3) float32.  Outputs for EVERY point (the caller compacts in order with `keep`): keep (n) uint8, world (n,3),
 * dir (n,3), dist (n) float64.  The outputs feed pcnerf_aabb_pack_train unchanged.
 * ------------------------------------------------------------------------------------------------------------- */
int pcnerf_frame_returns(const float* pts, int64_t n, const double* h_pose16, const float* pose_xy, int npose, float rdx,
                         float rdy, float rdz, float max_range, float over_height, float over_low, float interest_x,
                         float interest_y, const double* h_box6, const double* h_pos3, uint8_t* keep, double* world,
                         double* dir, double* dist, void* stream);

/* Multi-parent scenes (BASELINE.json configs[4]; the reference renders ONE parent block per run, README.md:46 describes a
 * large scene as a collection of them): which[i] = index of the first of the P parent boxes (P x 6 doubles, min xyz then
 * max xyz, closed) that contains return i, -1 if none.  origins (F,3) != NULL: also the ray of every return from the sensor
 * position of its frame frame_id[i] (NULL: frame 0) -- origin_out (n,3), dir (n,3), dist (n), the expressions of
 * eval_kitti_render.py:706-709 -- in the layout pcnerf_aabb_groups_count / _fill take.  All float64, P <= 1024. */
int pcnerf_route_points(const double* pts, int64_t n, const double* boxes, int P, const double* origins,
                        const int32_t* frame_id, int32_t* which, double* origin_out, double* dir, double* dist,
                        void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * Optimizer step on one flat buffer: torch.optim.Adam as configured by nof/nof_utils.py:162-173
 * (amsgrad off, L2 weight decay) over n contiguous fp32 parameters, their gradients (the buffer the gradient all-reduce
 * runs on; grad_scale = 1/world folds the averaging in) and the two moment vectors.  step (int64), lr (float) and coef2
 * (2 floats of scratch) live on the device: the call is replayable from a CUDA graph.
 * ------------------------------------------------------------------------------------------------------------- */
int pcnerf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t* step,
                     const float* lr, float* coef2, float beta1, float beta2, float eps, float weight_decay,
                     float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCNERF_B200_H */
