// Memory layout of the MLP's saved-activation store and scratch area (shared by mlp.cu and mlp_tc.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#define MLP_MAX_SPLITS 64

__host__ __device__ inline int mlp_kin(int l) { return l == 0 ? 63 : (l == 4 ? 319 : 256); }    // Linear in_features
__host__ __device__ inline int mlp_kpad(int l) { return l == 0 ? 64 : (l == 4 ? 320 : 256); }   // padded to 64

static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

// saved  : H[0..7] (rows x 256, fp32 (precision 0) or fp16 (precision 1))  |  stats[8][4][256] fp32 = {mean, invstd, a = gamma*invstd, s = beta - mean*a}
// scratch: Wp[8] | Wf[8] | bf[8][256] | wout_f[256]+bout_f | coef[3][256] | gvec[rows] | G[2][rows x 256] |
//          partial[MAX_SPLITS][256][320] | dstat (fp64) : 16 x 512 stat slots + colsum[8][256]
struct MlpLayout {
    int64_t rows;
    int precision;
    size_t esz;                 // activation element size
    size_t h_bytes;             // one activation matrix
    size_t off_stats, saved_bytes;
    size_t off_wp[8], off_wf[8], off_bf, off_wout, off_coef, off_gvec, off_g[2], off_partial, off_dstat, off_tc;
    size_t off_hf[2], off_encb, off_rgwork;
    size_t n_dstat, scratch_bytes;

    MlpLayout(int64_t rows_, int precision_) : rows(rows_), precision(precision_) {
        esz = precision == 1 ? 2 : 4;
        h_bytes = al256((size_t)rows * 256 * esz);
        off_stats = 8 * h_bytes;
        saved_bytes = off_stats + al256(8 * 4 * 256 * sizeof(float));
        size_t o = 0;
        for (int l = 0; l < 8; ++l) { off_wp[l] = o; o += al256((size_t)256 * mlp_kpad(l) * 4); }
        for (int l = 0; l < 8; ++l) { off_wf[l] = o; o += al256((size_t)256 * mlp_kpad(l) * 4); }
        off_bf = o; o += al256(8 * 256 * 4);
        off_wout = o; o += al256(512 * 4);
        off_coef = o; o += al256(4 * 256 * 4);
        // rows-independent regions first, so that their offsets are the same for every chunk of a pass (the weight
        // copies written for the first chunk are reused by the following, possibly shorter, ones: params.prepared)
        off_partial = o; o += al256((size_t)MLP_MAX_SPLITS * 256 * 320 * 4);
        n_dstat = 16 * 512 + 8 * 256;
        off_dstat = o; o += al256(n_dstat * sizeof(double));
        off_tc = o;              // 16-bit copies of the weights for the tensor-core path
        o += al256((size_t)8 * 256 * 384 * 2 + (size_t)8 * 256 * 256 * 2 + 4096);
        off_hf[0] = off_hf[1] = off_encb = o;      // (unused since the weight-gradient kernel converts in shared memory)
        off_rgwork = o;                            // row-GEMM CTA counter + per-CTA statistic partials (precision 1)
        if (precision == 1) o += al256(256 + 160 * 2 * 256 * 8);
        off_gvec = o; o += al256((size_t)rows * 4);
        for (int i = 0; i < 2; ++i) { off_g[i] = o; o += al256((size_t)rows * 256 * esz); }
        scratch_bytes = o;
    }
    float* H(char* sv, int l) const { return (float*)(sv + (size_t)l * h_bytes); }
    void* Hraw(char* sv, int l) const { return (void*)(sv + (size_t)l * h_bytes); }
    float* stats(char* sv, int l) const { return (float*)(sv + off_stats) + (size_t)l * 4 * 256; }
    float* Wp(char* sc, int l) const { return (float*)(sc + off_wp[l]); }
    float* Wf(char* sc, int l) const { return (float*)(sc + off_wf[l]); }
    float* bf(char* sc, int l) const { return (float*)(sc + off_bf) + (size_t)l * 256; }
    float* wout_f(char* sc) const { return (float*)(sc + off_wout); }
    float* coef(char* sc) const { return (float*)(sc + off_coef); }
    float* gvec(char* sc) const { return (float*)(sc + off_gvec); }
    float* G(char* sc, int i) const { return (float*)(sc + off_g[i]); }
    void* Graw(char* sc, int i) const { return (void*)(sc + off_g[i]); }
    float* partial(char* sc) const { return (float*)(sc + off_partial); }
    double* dstat(char* sc, int slot) const { return (double*)(sc + off_dstat) + (size_t)slot * 512; }
    double* colsum(char* sc, int l) const { return (double*)(sc + off_dstat) + 16 * 512 + (size_t)l * 256; }
    char* tc(char* sc) const { return sc + off_tc; }
    void* hf(char* sc, int i) const { return (void*)(sc + off_hf[i]); }
    void* rgwork(char* sc) const { return (void*)(sc + off_rgwork); }
    void* encb(char* sc) const { return (void*)(sc + off_encb); }
};
