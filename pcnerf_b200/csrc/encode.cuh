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
