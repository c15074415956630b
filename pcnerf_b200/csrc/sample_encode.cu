// K2 / K2': sample placement along each segment, hierarchical resampling and the 63-d positional encoding, fused.
//
// Mapping: one warp per ray.  The ray's depths live in shared memory; z placement uses explicit
// round-to-nearest mul/add (no FMA contraction) in the reference's operation order so that every later
// mask decision sees bit-identical z.  The encoding is produced 32 samples at a time, one sample per lane (see
// encode_ray_t), and written as coalesced 512-byte runs -- the (rows,64) layout (256 B fp32 / 128 B fp16 per row) is the
// MLP's A operand, column 63 is zero padding.
#include "common.cuh"
#include "encode.cuh"         // ENC_BIG, enc_phase (shared with the closed-form kernels of affine.cu)
#include <cuda_fp16.h>

#define SE_MAX_SMEM (200 * 1024)

__host__ __device__ __forceinline__ int al4(int v) { return (v + 3) & ~3; }
// floats of per-warp row staging: 32 rows x 64 columns, fp32 (256 B rows) if fp32 rows are written, else fp16
static inline int stage_floats(const void* out_enc, const void* out_f16) { return out_enc ? 2048 : (out_f16 ? 1024 : 0); }

__device__ __forceinline__ float lerp_rn(float a, float b, float s) {
    // near * (1 - s) + far * s, nof/render.py:432
    return __fadd_rn(__fmul_rn(a, __fsub_rn(1.f, s)), __fmul_rn(b, s));
}

// ---- positional encoding of one ray (models.py:27-41), 32 samples at a time, one sample per lane.
//
// The argument of every sin/cos is 2^k * x with x an fp32 value, so the range reduction can be done EXACTLY once per
// coordinate: t = x / (2 pi) in double (relative error 2^-53), q = round(t * 2^41) as a 64-bit integer, and the
// fractional part of 2^k * t is a bit field of q -- no per-frequency Cody-Waite reduction.  Two back ends:
//   * fp32 rows (precision 0 / 2, the 1e-5 gates): the 32-bit fraction goes through sincospif (<= 2 ulp);
//   * fp16 rows (precision 1): the top 23 fraction bits are spliced into a float and fed to MUFU.SIN / MUFU.COS
//     (abs error ~5e-7 on a reduced argument, 1000 x below the fp16 rounding of the stored value).
// Each lane builds its whole 64-column row in registers; rows are transposed through a swizzled shared-memory tile so
// that the global stores are 512 contiguous bytes per instruction (the 32 rows of a round are contiguous in HBM).
template <bool FAST>
__device__ __forceinline__ void enc_coord(float x, float* __restrict__ e, int c) {
    // e: this lane's row; writes columns 3 + 6k + c (sin) and 6 + 6k + c (cos), k = 0..9
    if (fabsf(x) < ENC_BIG) {
        const long long q = enc_phase(x);
        const uint32_t lo = (uint32_t)q, hi = (uint32_t)((unsigned long long)q >> 32);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            float sn, cs;
            if (FAST) {
                // bits [40-k : 18-k] of q = top 23 bits of frac(2^k t); flipping the top one adds half a turn, so that
                // f - 1.5 = frac - 1/2 (mod 1) - ... lands in [-0.5, 0.5) with sin/cos of the ORIGINAL angle
                const uint32_t w = __funnelshift_r(lo, hi, 18 - k);
                const float f = __uint_as_float((w & 0x7FFFFFu) ^ 0x3FC00000u);
                const float ang = fmaf(f - 1.5f, 6.2831853071795865f, 3.7450703e-7f);   // + half an lsb (2^-24 turns)
                sn = __sinf(ang);
                cs = __cosf(ang);
            } else {
                const int fr = (int)__funnelshift_r(lo, hi, 9 - k);                       // frac(2^k t) in 2^-32 turns, signed
                sincospif((float)fr * 4.656612873077393e-10f, &sn, &cs);                  // * 2^-31: [-1, 1) half-turns
            }
            e[3 + 6 * k + c] = sn;
            e[6 + 6 * k + c] = cs;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            float sn, cs;
            sincosf((float)(1 << k) * x, &sn, &cs);
            e[3 + 6 * k + c] = sn;
            e[6 + 6 * k + c] = cs;
        }
    }
}

__device__ __forceinline__ uint32_t enc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Encode P samples of one ray (depths in shared `zs`) into rows [row0, row0+P).  `stage`: 32 x (F32 ? 256 : 128) bytes.
template <bool F32, bool F16>
__device__ __forceinline__ void encode_ray_t(const float o0, const float o1, const float o2, const float d0,
                                             const float d1, const float d2, const float* zs, int P, int64_t row0,
                                             float* __restrict__ out_enc, __half* __restrict__ out_bf,
                                             float* stage, int lane) {
    const uint32_t sbase = enc_smem_u32(stage);
    for (int j0 = 0; j0 < P; j0 += 32) {
        const int j = j0 + lane;
        const float z = zs[j < P ? j : P - 1];
        float e[64];
        // rays_o + rays_d * z, nof/render.py:458 (mul then add, no FMA)
        e[0] = __fadd_rn(o0, __fmul_rn(d0, z));
        e[1] = __fadd_rn(o1, __fmul_rn(d1, z));
        e[2] = __fadd_rn(o2, __fmul_rn(d2, z));
        e[63] = 0.f;
        enc_coord<!F32>(e[0], e, 0);
        enc_coord<!F32>(e[1], e, 1);
        enc_coord<!F32>(e[2], e, 2);
        const int nrows = min(32, P - j0);
        if (F32) {
            // chunk c (16 B) of lane's 256-byte row at lane*256 + ((c & 8) | ((c ^ lane) & 7)) * 16: conflict-free both ways
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const uint32_t a = sbase + (uint32_t)(lane * 256 + (((c & 8) | ((c ^ lane) & 7)) << 4));
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(e[4 * c]), "f"(e[4 * c + 1]),
                             "f"(e[4 * c + 2]), "f"(e[4 * c + 3]) : "memory");
            }
            __syncwarp();
            float4* dst = reinterpret_cast<float4*>(out_enc + (row0 + j0) * 64);
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int idx = it * 32 + lane, row = idx >> 4, c = idx & 15;
                if (row < nrows) {
                    float4 v;
                    const uint32_t a = sbase + (uint32_t)(row * 256 + (((c & 8) | ((c ^ row) & 7)) << 4));
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
                    __stcs(dst + idx, v);
                }
            }
            __syncwarp();
        }
        if (F16) {
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const __half2 h = __floats2half2_rn(e[2 * i], e[2 * i + 1]);
                pk[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t a = sbase + (uint32_t)(lane * 128 + (((c ^ lane) & 7) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                             "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
            }
            __syncwarp();
            uint4* dst = reinterpret_cast<uint4*>(out_bf + (row0 + j0) * 64);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int idx = it * 32 + lane, row = idx >> 3, c = idx & 7;
                if (row < nrows) {
                    uint4 v;
                    const uint32_t a = sbase + (uint32_t)(row * 128 + (((c ^ row) & 7) << 4));
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
                    dst[idx] = v;
                }
            }
            __syncwarp();
        }
    }
}

// ---- fp16 rows only (precision 1, the tensor-core path): the same values as encode_ray_t<false, true>, produced octave
// by octave and packed into the 32 half2 registers of the row as they appear (column order x0 x1 x2 | s_k0 s_k1 s_k2
// c_k0 c_k1 c_k2 ...), instead of building the 64-float row first.  ~60 live registers instead of 128: four CTAs per SM
// instead of two (ncu on the 128-register form: MUFU pipe 45 % busy, issue slots 52 % busy, 4 warps per scheduler --
// neither pipe hidden behind the other).  Lanes with an out-of-range / non-finite coordinate (|x| >= ENC_BIG: never in
// a real scene) rewrite their own staged row from a scalar loop afterwards.
__device__ __forceinline__ void sincos_q(uint32_t lo, uint32_t hi, int k, float& sn, float& cs) {
    const uint32_t w = __funnelshift_r(lo, hi, 18 - k);
    const float f = __uint_as_float((w & 0x7FFFFFu) ^ 0x3FC00000u);
    const float ang = fmaf(f - 1.5f, 6.2831853071795865f, 3.7450703e-7f);
    sn = __sinf(ang);
    cs = __cosf(ang);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __noinline__ void encode_row_slow_f16(float x0, float x1, float x2, uint32_t row_base, int lane) {
    auto put = [&](int col, float v) {
        const uint32_t a = row_base + (uint32_t)(((((col >> 3) ^ lane) & 7) << 4) + ((col & 7) << 1));
        const unsigned short h = __half_as_ushort(__float2half_rn(v));
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(h) : "memory");
    };
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        const float x = c == 0 ? x0 : (c == 1 ? x1 : x2);
        const bool fast = fabsf(x) < ENC_BIG;
        const long long q = enc_phase(fast ? x : 0.f);
        const uint32_t lo = (uint32_t)q, hi = (uint32_t)((unsigned long long)q >> 32);
#pragma unroll 1
        for (int k = 0; k < 10; ++k) {
            float sn, cs;
            if (fast) sincos_q(lo, hi, k, sn, cs);
            else sincosf((float)(1 << k) * x, &sn, &cs);
            put(3 + 6 * k + c, sn);
            put(6 + 6 * k + c, cs);
        }
    }
}
__device__ __forceinline__ void encode_ray_f16(const float o0, const float o1, const float o2, const float d0,
                                               const float d1, const float d2, const float* zs, int P, int64_t row0,
                                               __half* __restrict__ out_bf, float* stage, int lane) {
    const uint32_t sbase = enc_smem_u32(stage);
    for (int j0 = 0; j0 < P; j0 += 32) {
        const int j = j0 + lane;
        const float z = zs[j < P ? j : P - 1];
        // rays_o + rays_d * z, nof/render.py:458 (mul then add, no FMA)
        const float x0 = __fadd_rn(o0, __fmul_rn(d0, z));
        const float x1 = __fadd_rn(o1, __fmul_rn(d1, z));
        const float x2 = __fadd_rn(o2, __fmul_rn(d2, z));
        const bool big = !(fabsf(x0) < ENC_BIG && fabsf(x1) < ENC_BIG && fabsf(x2) < ENC_BIG);
        const long long q0 = enc_phase(x0), q1 = enc_phase(x1), q2 = enc_phase(x2);
        const uint32_t l0 = (uint32_t)q0, h0 = (uint32_t)((unsigned long long)q0 >> 32);
        const uint32_t l1 = (uint32_t)q1, h1 = (uint32_t)((unsigned long long)q1 >> 32);
        const uint32_t l2 = (uint32_t)q2, h2 = (uint32_t)((unsigned long long)q2 >> 32);
        uint32_t pk[32];
        pk[0] = pack_h2(x0, x1);
        float carry = x2;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            float s0, c0, s1, c1, s2, c2;
            sincos_q(l0, h0, k, s0, c0);
            sincos_q(l1, h1, k, s1, c1);
            sincos_q(l2, h2, k, s2, c2);
            pk[1 + 3 * k] = pack_h2(carry, s0);
            pk[2 + 3 * k] = pack_h2(s1, s2);
            pk[3 + 3 * k] = pack_h2(c0, c1);
            carry = c2;
        }
        pk[31] = pack_h2(carry, 0.f);
        const int nrows = min(32, P - j0);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint32_t a = sbase + (uint32_t)(lane * 128 + (((c ^ lane) & 7) << 4));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                         "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
        }
        if (big) encode_row_slow_f16(x0, x1, x2, sbase + (uint32_t)(lane * 128), lane);
        __syncwarp();
        uint4* dst = reinterpret_cast<uint4*>(out_bf + (row0 + j0) * 64);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = it * 32 + lane, row = idx >> 3, c = idx & 7;
            if (row < nrows) {
                uint4 v;
                const uint32_t a = sbase + (uint32_t)(row * 128 + (((c ^ row) & 7) << 4));
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
                dst[idx] = v;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void encode_ray(const float o0, const float o1, const float o2, const float d0,
                                           const float d1, const float d2, const float* zs, int P, int64_t row0,
                                           float* __restrict__ out_enc, __half* __restrict__ out_bf,
                                           float* stage, int lane) {
    if (out_enc && out_bf) encode_ray_t<true, true>(o0, o1, o2, d0, d1, d2, zs, P, row0, out_enc, out_bf, stage, lane);
    else if (out_enc) encode_ray_t<true, false>(o0, o1, o2, d0, d1, d2, zs, P, row0, out_enc, out_bf, stage, lane);
    else encode_ray_t<false, true>(o0, o1, o2, d0, d1, d2, zs, P, row0, out_enc, out_bf, stage, lane);
}

// out[i + #{b (<|<=) a[i]}] = a[i] for every element of the ascending list a: one half of a stable merge by rank.
// Each lane runs RB binary searches at once with a fixed step count (bit_length(nb) probes close any interval of [0, nb]):
// the probes of one search form a dependent chain of shared-memory loads (~45 cycles per step), and ncu had put half of
// the resampling kernel's stall samples on the three one-search-at-a-time `while (lo < hi)` loops of this file.
#define RB 4
template <bool LE>
__device__ __forceinline__ void rank_scatter(const float* a, int na, const float* b, int nb, float* out, int lane) {
    const int steps = 32 - __clz(nb);
    for (int base = 0; base < na; base += 32 * RB) {
        float v[RB];
        int lo[RB], hi[RB];
#pragma unroll
        for (int e = 0; e < RB; ++e) {
            const int i = base + 32 * e + lane;
            v[e] = a[i < na ? i : na - 1];
            lo[e] = 0;
            hi[e] = i < na ? nb : 0;
        }
        for (int s = 0; s < steps; ++s) {
#pragma unroll
            for (int e = 0; e < RB; ++e) {
                const bool act = lo[e] < hi[e];
                const int m = (lo[e] + hi[e]) >> 1;
                const float bv = b[act ? m : 0];
                const bool right = LE ? (bv <= v[e]) : (bv < v[e]);
                if (act) { if (right) lo[e] = m + 1; else hi[e] = m; }
            }
        }
#pragma unroll
        for (int e = 0; e < RB; ++e) {
            const int i = base + 32 * e + lane;
            if (i < na) out[i + lo[e]] = v[e];
        }
    }
}

// stable merge of two ascending lists a (na) and b (nb) into out (na+nb); ties: a first (torch.sort of cat([a,b]))
__device__ __forceinline__ void rank_merge(const float* a, int na, const float* b, int nb, float* out, int lane) {
    rank_scatter<false>(a, na, b, nb, out, lane);         // a[i] lands after the b's that are <  it
    rank_scatter<true>(b, nb, a, na, out, lane);          // b[j] lands after the a's that are <= it
}

__device__ __forceinline__ void bitonic_sort_smem(float* s, int npad, int lane) {
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < npad; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float a = s[i], b = s[ixj];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc) { s[i] = b; s[ixj] = a; }
                }
            }
            __syncwarp();
        }
    }
}

// The same network with the EPL = npad/32 elements [lane*EPL, lane*EPL + EPL) of every lane in registers: compare-exchanges
// at distance j < EPL stay inside the lane, the others are one shfl_xor per element (Ni = 128 / 256 -> 4 / 8 registers).
template <int EPL>
__device__ __forceinline__ void bitonic_sort_regs(float* s, int lane) {
    float v[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) v[i] = s[lane * EPL + i];
#pragma unroll
    for (int k = 2; k <= EPL * 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= EPL) {
                const int lj = j / EPL;
                const bool up = ((lane * EPL) & k) == 0, lower = (lane & lj) == 0;
#pragma unroll
                for (int i = 0; i < EPL; ++i) {
                    const float o = __shfl_xor_sync(FULL_MASK, v[i], lj);
                    v[i] = (lower == up) ? fminf(v[i], o) : fmaxf(v[i], o);
                }
            } else {
#pragma unroll
                for (int i = 0; i < EPL; ++i) {
                    if ((i & j) == 0) {
                        const bool up = (((lane * EPL) + i) & k) == 0;
                        const float a = v[i], b = v[i | j];
                        const float lo = fminf(a, b), hi = fmaxf(a, b);
                        v[i] = up ? lo : hi;
                        v[i | j] = up ? hi : lo;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < EPL; ++i) s[lane * EPL + i] = v[i];
    __syncwarp();
}

__device__ __forceinline__ void bitonic_sort(float* s, int npad, int lane) {
    switch (npad) {
        case 32: bitonic_sort_regs<1>(s, lane); break;
        case 64: bitonic_sort_regs<2>(s, lane); break;
        case 128: bitonic_sort_regs<4>(s, lane); break;
        case 256: bitonic_sort_regs<8>(s, lane); break;
        case 512: bitonic_sort_regs<16>(s, lane); break;
        default: bitonic_sort_smem(s, npad, lane);
    }
}

// sample_pdf body (render.py:371-410) for one ray: `wts` are the nb-1 raw weights, bins/cdf are nb-long shared arrays.
__device__ __forceinline__ void inverse_cdf(const float* bins, float* cdf, const float* __restrict__ wts, int nb,
                                            const float* __restrict__ u, int Ni, int NiPad, float* zs, int lane) {
    const int nw = nb - 1;
    // torch.sum (:374) has no portable summation order (vector width dependent); the correctly rounded sum is used.
    double part = 0.0;
    for (int i = lane; i < nw; i += 32) part += (double)__fadd_rn(wts[i], 1e-5f);     // weights + 1e-5 (:373)
    const float total = (float)warp_sum_d(part);
    // cdf = [0, cumsum(pdf)] (:375-376): torch's CPU cumsum accumulates float inputs in double and rounds every
    // prefix to float -- reproduced with a double warp scan in rounds of 32 (bit-exact against the fixtures).
    double carry = 0.0;
    if (lane == 0) cdf[0] = 0.f;
    for (int base = 0; base < nw; base += 32) {
        const int i = base + lane;
        double v = i < nw ? (double)__fdiv_rn(__fadd_rn(wts[i], 1e-5f), total) : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(FULL_MASK, v, o);
            if (lane >= o) v += t;
        }
        v += carry;
        if (i < nw) cdf[i + 1] = (float)v;
        carry = __shfl_sync(FULL_MASK, v, 31);
    }
    __syncwarp();
    // searchsorted(cdf, u, right=True) (:380), RB samples per lane at a time (see rank_scatter)
    const int steps = 32 - __clz(nb);
    for (int base = 0; base < NiPad; base += 32 * RB) {
        float uu[RB];
        int lo[RB], hi[RB];
#pragma unroll
        for (int e = 0; e < RB; ++e) {
            const int j = base + 32 * e + lane;
            uu[e] = j < Ni ? u[j] : 0.f;
            lo[e] = 0;
            hi[e] = j < Ni ? nb : 0;
        }
        for (int s = 0; s < steps; ++s) {
#pragma unroll
            for (int e = 0; e < RB; ++e) {
                const bool act = lo[e] < hi[e];
                const int m = (lo[e] + hi[e]) >> 1;
                const bool right = cdf[act ? m : 0] <= uu[e];
                if (act) { if (right) lo[e] = m + 1; else hi[e] = m; }
            }
        }
#pragma unroll
        for (int e = 0; e < RB; ++e) {
            const int j = base + 32 * e + lane;
            if (j >= NiPad) continue;
            float out = INFINITY;
            if (j < Ni) {
                const int below = max(lo[e] - 1, 0), above = min(lo[e], nb - 1);
                const float cb = cdf[below], bb = bins[below];
                float denom = __fsub_rn(cdf[above], cb);
                if (denom < 1e-5f) denom = 1.f;
                const float t = __fdiv_rn(__fsub_rn(uu[e], cb), denom);
                out = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(bins[above], bb)));
            }
            zs[j] = out;
        }
    }
    __syncwarp();
}

__global__ void k_sample_pdf(const float* __restrict__ bins_g, const float* __restrict__ wts, int64_t n, int nb,
                             const float* __restrict__ u, int u_ld, int Ni, float* __restrict__ out) {
    extern __shared__ float smf[];
    const int lane = threadIdx.x & 31, wib = warp_in_block(), wpb = blockDim.x >> 5;
    float* bins = smf + (size_t)wib * (2 * nb + Ni);
    float* cdf = bins + nb;
    float* zs = cdf + nb;
    for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < n; r += (int64_t)gridDim.x * wpb) {
        for (int i = lane; i < nb; i += 32) bins[i] = bins_g[r * nb + i];
        __syncwarp();
        inverse_cdf(bins, cdf, wts + r * (nb - 1), nb, u + (u_ld ? r * (int64_t)u_ld : 0), Ni, Ni, zs, lane);
        for (int j = lane; j < Ni; j += 32) out[r * Ni + j] = zs[j];
        __syncwarp();
    }
}

// LIGHT: fp16 rows only (out_enc == nullptr, out_bf != nullptr) -> encode_ray_f16, four CTAs per SM
template <bool LIGHT>
__global__ void __launch_bounds__(256, LIGHT ? 4 : 2) k_sample_encode_coarse(const float* __restrict__ rays, int ld, int64_t n, int near_col, int far_col,
                                       int cnear_col, int cfar_col, const float* __restrict__ steps_a, int n_a,
                                       const float* __restrict__ steps_b, int n_b, int use_disp, float perturb,
                                       const float* __restrict__ U, float* __restrict__ out_z,
                                       float* __restrict__ out_enc, __half* __restrict__ out_bf, int stage_f) {
    extern __shared__ __align__(16) float smf[];
    const int S = n_a + n_b;
    const int lane = threadIdx.x & 31, wib = warp_in_block(), wpb = blockDim.x >> 5;
    float* za = smf + (size_t)wib * (al4(2 * S) + stage_f);
    float* tmp = za + S;
    float* stage = za + al4(2 * S);
    for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < n; r += (int64_t)gridDim.x * wpb) {
        const float* ray = rays + r * ld;
        const float near = ray[near_col], far = ray[far_col];
        // ---- z placement (render.py:430-442, :565-570)
        const bool asc_a = !(far < near);
        for (int i = lane; i < n_a; i += 32) {
            const float s = steps_a[i];
            float v;
            if (use_disp) {
                const float inv = __fadd_rn(__fmul_rn(__fdiv_rn(1.f, near), __fsub_rn(1.f, s)),
                                            __fmul_rn(__fdiv_rn(1.f, far), s));
                v = __fdiv_rn(1.f, inv);
            } else {
                v = lerp_rn(near, far, s);
            }
            tmp[(asc_a || n_b == 0) ? i : n_a - 1 - i] = v;
        }
        if (n_b > 0) {
            const float cn = ray[cnear_col], cf = ray[cfar_col];
            const bool asc_b = !(cf < cn);
            for (int i = lane; i < n_b; i += 32) tmp[n_a + (asc_b ? i : n_b - 1 - i)] = lerp_rn(cn, cf, steps_b[i]);
        }
        __syncwarp();
        if (n_b > 0) rank_merge(tmp, n_a, tmp + n_a, n_b, za, lane);
        else for (int i = lane; i < S; i += 32) za[i] = tmp[i];
        __syncwarp();
        // ---- stratified jitter (render.py:449-454)
        if (perturb > 0.f) {
            for (int i = lane; i < S; i += 32) {
                const float zi = za[i];
                const float lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(za[i - 1], zi)) : zi;
                const float upper = i < S - 1 ? __fmul_rn(0.5f, __fadd_rn(zi, za[i + 1])) : zi;
                const float pr = __fmul_rn(perturb, U[r * S + i]);
                tmp[i] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), pr));
            }
            __syncwarp();
            for (int i = lane; i < S; i += 32) za[i] = tmp[i];
            __syncwarp();
        }
        for (int i = lane; i < S; i += 32) out_z[r * S + i] = za[i];
        if (LIGHT) encode_ray_f16(ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], za, S, r * S, out_bf, stage, lane);
        else if (out_enc || out_bf)
            encode_ray(ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], za, S, r * S, out_enc, out_bf, stage, lane);
        __syncwarp();
    }
}

template <bool LIGHT>
__global__ void __launch_bounds__(256, LIGHT ? 4 : 2) k_sample_encode_fine(const float* __restrict__ rays, int ld, int64_t n, const float* __restrict__ z,
                                     const float* __restrict__ w, int S, const float* __restrict__ u, int u_ld, int Ni,
                                     int NiPad, float* __restrict__ out_z, float* __restrict__ out_enc,
                                     __half* __restrict__ out_bf, int stage_f) {
    extern __shared__ __align__(16) float smf[];
    const int F = S + Ni;
    const int lane = threadIdx.x & 31, wib = warp_in_block(), wpb = blockDim.x >> 5;
    const size_t arrays = al4(S + 2 * (S - 1) + NiPad + F);
    float* zc = smf + wib * (arrays + stage_f);
    float* bins = zc + S;
    float* cdf = bins + (S - 1);
    float* zs = cdf + (S - 1);
    float* zo = zs + NiPad;
    float* stage = zc + arrays;
    const int nb = S - 1;                      // len(bins) == len(cdf)
    for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < n; r += (int64_t)gridDim.x * wpb) {
        for (int i = lane; i < S; i += 32) zc[i] = z[r * S + i];
        __syncwarp();
        // bins = .5 * (z[1:] + z[:-1])  (render.py:463);  weights = w[1:-1] (render.py:464)
        for (int i = lane; i < nb; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(zc[i + 1], zc[i]));
        inverse_cdf(bins, cdf, w + r * S + 1, nb, u + (u_ld ? r * (int64_t)u_ld : 0), Ni, NiPad, zs, lane);
        // torch.sort(cat([z, z_samples])) (render.py:467): sort the new samples only if they are not already
        // ascending, then merge with the (ascending) coarse depths.
        int unsorted = 0;
        for (int j = lane; j + 1 < Ni; j += 32) unsorted |= (zs[j] > zs[j + 1]);
        if (__any_sync(FULL_MASK, unsorted)) bitonic_sort(zs, NiPad, lane);
        rank_merge(zc, S, zs, Ni, zo, lane);
        __syncwarp();
        for (int i = lane; i < F; i += 32) out_z[r * F + i] = zo[i];
        const float* ray = rays + r * ld;
        if (LIGHT) encode_ray_f16(ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], zo, F, r * F, out_bf, stage, lane);
        else if (out_enc || out_bf)
            encode_ray(ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], zo, F, r * F, out_enc, out_bf, stage, lane);
        __syncwarp();
    }
}

// Embedding.forward alone (models.py:27-41): one warp per point.
__global__ void k_embed(const float* __restrict__ x, int64_t b, float* __restrict__ out, int out_ld) {
    __shared__ float stage_all[8 * 64];
    const int lane = threadIdx.x & 31, wib = warp_in_block(), wpb = blockDim.x >> 5;
    float* st = stage_all + wib * 64;
    const int k = lane / 3, c = lane - 3 * k;
    const float freq = (float)(1 << (k < 10 ? k : 0));
    for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < b; r += (int64_t)gridDim.x * wpb) {
        const float x0 = x[3 * r], x1 = x[3 * r + 1], x2 = x[3 * r + 2];
        if (lane < 30) {
            const float xv = c == 0 ? x0 : (c == 1 ? x1 : x2);
            float s, cs;
            sincosf(freq * xv, &s, &cs);
            st[3 + 6 * k + c] = s;
            st[6 + 6 * k + c] = cs;
        } else if (lane == 30) {
            st[0] = x0; st[1] = x1; st[2] = x2; st[63] = 0.f;
        }
        __syncwarp();
        for (int col = lane; col < out_ld; col += 32) out[r * out_ld + col] = col < 63 ? st[col] : 0.f;
        __syncwarp();
    }
}

static int pick_warps(size_t per_warp_bytes, const char* what, int* wpb) {
    int w = (int)(SE_MAX_SMEM / per_warp_bytes);
    if (w < 1) {
        pcn_set_error("%s: %zu bytes of shared memory per ray exceed the %d-byte budget", what, per_warp_bytes,
                      SE_MAX_SMEM);
        return PCNERF_ERR_UNSUPPORTED;
    }
    *wpb = w > 8 ? 8 : w;
    return 0;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

extern "C" int pcnerf_sample_encode_coarse(const float* rays, int ld, int64_t n, int near_col, int far_col,
                                           int cnear_col, int cfar_col, const float* steps_a, int n_a,
                                           const float* steps_b, int n_b, int use_disp, float perturb, const float* U,
                                           float* out_z, float* out_enc, void* out_enc_f16, void* stream) {
    PCN_CHECK_ARG(n >= 0 && ld >= 8 && n_a >= 1 && n_b >= 0, "sample_encode_coarse: bad sizes");
    PCN_CHECK_ARG(steps_a && (n_b == 0 || steps_b), "sample_encode_coarse: linspace tables missing");
    PCN_CHECK_ARG(!(perturb > 0.f) || U, "sample_encode_coarse: perturb > 0 needs pre-drawn U");
    PCN_CHECK_ARG(near_col < ld && far_col < ld && (n_b == 0 || (cnear_col < ld && cfar_col < ld)),
                  "sample_encode_coarse: column index out of range");
    if (n == 0) return 0;
    const int S = n_a + n_b;
    const int stage_f = stage_floats(out_enc, out_enc_f16);
    const size_t per_warp = (size_t)(al4(2 * S) + stage_f) * sizeof(float);
    int wpb;
    int rc = pick_warps(per_warp, "sample_encode_coarse", &wpb);
    if (rc) return rc;
    const size_t smem = per_warp * wpb;
    const bool light = !out_enc && out_enc_f16;
    if (smem > 48 * 1024)
        PCN_CUDA(cudaFuncSetAttribute(light ? k_sample_encode_coarse<true> : k_sample_encode_coarse<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = pcn_cdiv(n, wpb);
    const int64_t cap = (int64_t)PCN_SM_COUNT * 16;
    if (grid > cap) grid = cap;
    PcnScope ps(PCN_K_SAMPLE_ENCODE, (cudaStream_t)stream,
                (double)n * (60.0 + S * (4.0 + (out_enc ? 256.0 : 0.0) + (out_enc_f16 ? 128.0 : 0.0))));
    (light ? k_sample_encode_coarse<true> : k_sample_encode_coarse<false>)<<<(int)grid, wpb * 32, smem, (cudaStream_t)stream>>>(
        rays, ld, n, near_col, far_col, cnear_col, cfar_col, steps_a, n_a, steps_b, n_b, use_disp, perturb, U, out_z,
        out_enc, (__half*)out_enc_f16, stage_f);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_sample_encode_fine(const float* rays, int ld, int64_t n, const float* z, const float* w, int S,
                                         const float* u, int u_ld, int Ni, float* out_z, float* out_enc,
                                         void* out_enc_f16, void* stream) {
    PCN_CHECK_ARG(n >= 0 && ld >= 6 && S >= 3 && Ni >= 1, "sample_encode_fine: bad sizes (need S >= 3, Ni >= 1)");
    PCN_CHECK_ARG(u && (u_ld == 0 || u_ld == Ni), "sample_encode_fine: u must be (Ni) with u_ld 0 or (n,Ni) with u_ld Ni");
    if (n == 0) return 0;
    const int NiPad = next_pow2(Ni);
    const int stage_f = stage_floats(out_enc, out_enc_f16);
    const size_t per_warp = ((size_t)al4(S + 2 * (S - 1) + NiPad + (S + Ni)) + stage_f) * sizeof(float);
    int wpb;
    int rc = pick_warps(per_warp, "sample_encode_fine", &wpb);
    if (rc) return rc;
    const size_t smem = per_warp * wpb;
    const bool light = !out_enc && out_enc_f16;
    if (smem > 48 * 1024)
        PCN_CUDA(cudaFuncSetAttribute(light ? k_sample_encode_fine<true> : k_sample_encode_fine<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = pcn_cdiv(n, wpb);
    const int64_t cap = (int64_t)PCN_SM_COUNT * 16;
    if (grid > cap) grid = cap;
    PcnScope ps(PCN_K_SAMPLE_ENCODE, (cudaStream_t)stream,
                (double)n * (60.0 + 8.0 * S + (S + Ni) * (4.0 + (out_enc ? 256.0 : 0.0) + (out_enc_f16 ? 128.0 : 0.0))));
    (light ? k_sample_encode_fine<true> : k_sample_encode_fine<false>)<<<(int)grid, wpb * 32, smem, (cudaStream_t)stream>>>(
        rays, ld, n, z, w, S, u, u_ld, Ni, NiPad, out_z, out_enc, (__half*)out_enc_f16, stage_f);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_sample_pdf(const float* bins, const float* weights, int64_t n, int nb, const float* u, int u_ld,
                                 int Ni, float* out, void* stream) {
    PCN_CHECK_ARG(n >= 0 && nb >= 2 && Ni >= 1, "sample_pdf: bad sizes");
    PCN_CHECK_ARG(u && (u_ld == 0 || u_ld == Ni), "sample_pdf: u must be (Ni) with u_ld 0 or (n,Ni) with u_ld Ni");
    if (n == 0) return 0;
    const size_t per_warp = ((size_t)2 * nb + Ni) * sizeof(float);
    int wpb;
    int rc = pick_warps(per_warp, "sample_pdf", &wpb);
    if (rc) return rc;
    const size_t smem = per_warp * wpb;
    if (smem > 48 * 1024)
        PCN_CUDA(cudaFuncSetAttribute(k_sample_pdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = pcn_cdiv(n, wpb);
    const int64_t cap = (int64_t)PCN_SM_COUNT * 16;
    if (grid > cap) grid = cap;
    PcnScope ps(PCN_K_SAMPLE_ENCODE, (cudaStream_t)stream, (double)n * (8.0 * nb + 4.0 * Ni));
    k_sample_pdf<<<(int)grid, wpb * 32, smem, (cudaStream_t)stream>>>(bins, weights, n, nb, u, u_ld, Ni, out);
    PCN_LAUNCH_CHECK();
    return 0;
}

extern "C" int pcnerf_embed(const float* x, int64_t b, float* out, int out_ld, void* stream) {
    PCN_CHECK_ARG(b >= 0 && out_ld >= 63, "embed: out_ld must be >= 63");
    if (b == 0) return 0;
    PCN_CHECK_ARG(x && out, "embed: null x / out");
    int64_t grid = pcn_cdiv(b, 8);
    const int64_t cap = (int64_t)PCN_SM_COUNT * 16;
    if (grid > cap) grid = cap;
    PcnScope ps(PCN_K_SAMPLE_ENCODE, (cudaStream_t)stream, (double)b * (12.0 + 4.0 * out_ld));
    k_embed<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(x, b, out, out_ld);
    PCN_LAUNCH_CHECK();
    return 0;
}
