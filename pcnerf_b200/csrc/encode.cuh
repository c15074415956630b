// Positional encoding of one sample position (models.py:27-41), shared by K2 (sample_encode.cu, which materialises the
// (rows,64) encoding tensor) and by the closed-form kernels (affine.cu), which re-derive a row's encoding from its ray and
// depth instead of reading it back from HBM.
//
// The argument of every sin/cos is 2^k * x with x an fp32 value, so the range reduction is done EXACTLY once per
// coordinate: t = x / (2 pi) in double (relative error 2^-53), q = round(t * 2^41) as a 64-bit integer, and the
// fractional part of 2^k * t is a bit field of q -- no per-frequency Cody-Waite reduction.
#pragma once
#include <cstdint>

#define ENC_BIG 8.0e6f          // |x| beyond this (or NaN/Inf): plain sincosf, whose result is what torch computes

__device__ __forceinline__ long long enc_phase(float x) {
    return __double2ll_rn((double)x * (0.15915494309189535 * 2199023255552.0));        // x / (2 pi) * 2^41
}

// fp32 values of columns 3 + 6k + C (sin) and 6 + 6k + C (cos), k = 0..9, of coordinate C -- the values
// enc_coord<false> (sample_encode.cu) writes -- handed to f(column, value); after unrolling every column is a
// compile-time constant, so a visitor that indexes a register array with it keeps the array in registers.
template <int C, class F>
__device__ __forceinline__ void enc_visit_coord(float x, F&& f) {
    if (fabsf(x) < ENC_BIG) {
        const long long q = enc_phase(x);
        const uint32_t lo = (uint32_t)q, hi = (uint32_t)((unsigned long long)q >> 32);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            float sn, cs;
            const int fr = (int)__funnelshift_r(lo, hi, 9 - k);                       // frac(2^k t) in 2^-32 turns, signed
            sincospif((float)fr * 4.656612873077393e-10f, &sn, &cs);                  // * 2^-31: [-1, 1) half-turns
            f(3 + 6 * k + C, sn);
            f(6 + 6 * k + C, cs);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            float sn, cs;
            sincosf((float)(1 << k) * x, &sn, &cs);
            f(3 + 6 * k + C, sn);
            f(6 + 6 * k + C, cs);
        }
    }
}

// all 63 columns of the row of position (x0, x1, x2)
template <class F>
__device__ __forceinline__ void enc_visit(float x0, float x1, float x2, F&& f) {
    f(0, x0);
    f(1, x1);
    f(2, x2);
    enc_visit_coord<0>(x0, f);
    enc_visit_coord<1>(x1, f);
    enc_visit_coord<2>(x2, f);
}

// ---- the same 63 values for kernels that RE-DERIVE a row instead of reading it (affine_rays.cu): sin / cos of the exact
// phase by quadrant + degree-7 / degree-8 polynomials on [-pi/4, pi/4] (Cephes sinf / cosf coefficients) instead of
// sincospif: ~24 instead of ~32 instructions per pair, max abs error 1.2e-7 against the exact angle (sincospif on the
// fp32-rounded phase: 9.4e-8 + its own 2 ulp) -- scripts-level study in DESIGN.md section 4.4.  Not bit-identical to
// enc_coord<false>: the closed-form engine is self-consistent (moments, apply and gradient all use this form).
__device__ __forceinline__ void sincos_turn(int fr, float& sn, float& cs) {
    // fr: angle in 2^-32 turns (signed).  q = nearest quarter turn, y = the rest in radians, |y| <= pi/4
    const uint32_t t = (uint32_t)fr + 0x20000000u;
    const uint32_t q = t >> 30;
    const float y = (float)((int)(t & 0x3FFFFFFFu) - 0x20000000) * 1.4629180792671596e-9f;      // 2 pi / 2^32
    const float z = y * y;
    const float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    const float s = fmaf(y * z, sp, y);
    const float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    const float c = fmaf(z * z, cp, fmaf(-0.5f, z, 1.f));
    const float a = (q & 1u) ? c : s, b = (q & 1u) ? s : c;          // sin(q pi/2 + y), cos(q pi/2 + y) up to sign
    sn = __uint_as_float(__float_as_uint(a) ^ ((q & 2u) << 30));
    cs = __uint_as_float(__float_as_uint(b) ^ (((q + 1u) & 2u) << 30));
}

// |x| >= ENC_BIG, NaN, Inf (never in a real scene): sc[2 k], sc[2 k + 1] = sincosf(2^k x), what torch computes.  OUT OF LINE:
// inlined, the thirty Payne-Hanek bodies made every kernel that re-derives encodings 3,700 - 5,400 SASS instructions long
// (60 - 85 KB), and half of the stall samples of the moment kernels' producer warps were instruction-fetch stalls.
static __device__ __noinline__ void enc_coord_slow(float x, float* sc) {
    for (int k = 0; k < 10; ++k) sincosf((float)(1 << k) * x, sc + 2 * k, sc + 2 * k + 1);
}

template <int C, class F>
__device__ __forceinline__ void enc_visit_coord_poly(float x, F&& f) {
    if (fabsf(x) < ENC_BIG) {
        const long long q = enc_phase(x);
        const uint32_t lo = (uint32_t)q, hi = (uint32_t)((unsigned long long)q >> 32);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            float sn, cs;
            sincos_turn((int)__funnelshift_r(lo, hi, 9 - k), sn, cs);
            f(3 + 6 * k + C, sn);
            f(6 + 6 * k + C, cs);
        }
    } else {
        float sc[20];
        enc_coord_slow(x, sc);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            f(3 + 6 * k + C, sc[2 * k]);
            f(6 + 6 * k + C, sc[2 * k + 1]);
        }
    }
}

template <class F>
__device__ __forceinline__ void enc_visit_poly(float x0, float x1, float x2, F&& f) {
    f(0, x0);
    f(1, x1);
    f(2, x2);
    enc_visit_coord_poly<0>(x0, f);
    enc_visit_coord_poly<1>(x1, f);
    enc_visit_coord_poly<2>(x2, f);
}
