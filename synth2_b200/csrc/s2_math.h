// s2_math.h — the three transcendentals of the render path (2^x, e^x, sin/cos), evaluated in binary64 and
// rounded once to binary32.
//
// The reference calls sleef `pow(2, x)` (process.rs:244), libm `powf` (process.rs:227), `expf`
// (filters.rs:21) and `sinf`/`cosf` (dsp_filters.rs:107-109); none is reproducible bit-for-bit on a GPU.
// A binary64 evaluation with ~1e-16 relative error, rounded once, is the correctly rounded binary32
// value except when the exact result lies within ~1e-9 ulp of a rounding boundary — which is also what
// glibc's float functions return in practice, so the CPU oracle (glibc) and this code agree on all but
// ~1e-8 of inputs (tools/check_math.cpp measures it).  Plain Taylor/Horner kernels after an exact range
// reduction: ~15 (exp) / ~30 (sincos) DFMA instead of the general-purpose libdevice routines, because
// the modulated-cutoff segment of every note evaluates them per frame.
//
// Host-and-device so the same source is checked on the CPU.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define S2_HD __host__ __device__ __forceinline__
#else
#define S2_HD static inline
#endif

// e^t for |t| <= 0.35 (Taylor, degree 14: truncation < 1e-19)
S2_HD double s2_exp_kernel(double t) {
    double p = 1.0 / 87178291200.0;                       // 1/14!
    p = fma(p, t, 1.0 / 6227020800.0);
    p = fma(p, t, 1.0 / 479001600.0);
    p = fma(p, t, 1.0 / 39916800.0);
    p = fma(p, t, 1.0 / 3628800.0);
    p = fma(p, t, 1.0 / 362880.0);
    p = fma(p, t, 1.0 / 40320.0);
    p = fma(p, t, 1.0 / 5040.0);
    p = fma(p, t, 1.0 / 720.0);
    p = fma(p, t, 1.0 / 120.0);
    p = fma(p, t, 1.0 / 24.0);
    p = fma(p, t, 1.0 / 6.0);
    p = fma(p, t, 0.5);
    p = fma(p, t, 1.0);
    p = fma(p, t, 1.0);
    return p;
}

// r * 2^k for a normal double r in [0.5, 2] and |k| < 1000, by exponent arithmetic
S2_HD double s2_scale2(double r, int k) {
    int64_t bits;
    memcpy(&bits, &r, 8);
    bits += (int64_t)k << 52;
    memcpy(&r, &bits, 8);
    return r;
}

// 2^x, x binary32.  Exact reduction x = k + f, |f| <= 0.5.
S2_HD float s2_exp2f(float x) {
    if (!(x > -150.0f)) return x != x ? x : 0.0f;
    if (x > 128.0f) return INFINITY;
    const double xd = (double)x;
    const double kd = rint(xd);
    const double f = xd - kd;                             // exact
    const double r = s2_exp_kernel(f * 0.693147180559945309417232121458);
    return (float)s2_scale2(r, (int)kd);                  // one rounding (subnormal results round here too)
}

// e^x, x binary32.  Cody-Waite reduction x = k*ln2 + r, |r| <= 0.35.
S2_HD float s2_expf(float x) {
    if (!(x > -104.0f)) return x != x ? x : 0.0f;
    if (x > 89.0f) return INFINITY;
    const double xd = (double)x;
    const double kd = rint(xd * 1.44269504088896340735992468100);
    const double r = fma(-kd, 1.90821492927058770002e-10, fma(-kd, 6.93147180369123816490e-01, xd));
    return (float)s2_scale2(s2_exp_kernel(r), (int)kd);
}

// sin and cos of x (binary32, |x| < ~1e5) — one reduction, two Taylor kernels on |r| <= pi/4.
S2_HD void s2_sincosf(float x, float* s, float* c) {
    const double xd = (double)x;
    const double qd = rint(xd * 0.636619772367581343075535053490);          // x * 2/pi
    // pi/2 = pio2_1 + pio2_1t (+ 2e-21): pio2_1 has 33 significant bits, so qd * pio2_1 is exact
    double r = fma(-qd, 1.57079632673412561417e+00, xd);
    r = fma(-qd, 6.07710050650619224932e-11, r);
    const double z = r * r;
    double ps = -1.0 / 355687428096000.0;                 // -1/17!
    ps = fma(ps, z, 1.0 / 1307674368000.0);               //  1/15!
    ps = fma(ps, z, -1.0 / 6227020800.0);
    ps = fma(ps, z, 1.0 / 39916800.0);
    ps = fma(ps, z, -1.0 / 362880.0);
    ps = fma(ps, z, 1.0 / 5040.0);
    ps = fma(ps, z, -1.0 / 120.0);
    ps = fma(ps, z, 1.0 / 6.0);
    const double sr = fma(-ps * z, r, r);                 // r - r^3 * (1/6 - ...)
    double pc = 1.0 / 20922789888000.0;                   //  1/16!
    pc = fma(pc, z, -1.0 / 87178291200.0);
    pc = fma(pc, z, 1.0 / 479001600.0);
    pc = fma(pc, z, -1.0 / 3628800.0);
    pc = fma(pc, z, 1.0 / 40320.0);
    pc = fma(pc, z, -1.0 / 720.0);
    pc = fma(pc, z, 1.0 / 24.0);
    pc = fma(pc, z, -0.5);
    const double cr = fma(pc, z, 1.0);
    const int q = (int)qd & 3;
    const double sv = (q & 1) ? cr : sr;
    const double cv = (q & 1) ? sr : cr;
    *s = (float)((q & 2) ? -sv : sv);
    *c = (float)(((q + 1) & 2) ? -cv : cv);
}
