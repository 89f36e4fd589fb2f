// s2_math.h — the transcendentals of the render path (2^x, e^x, sin/cos, tan), evaluated in binary64 and
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

// Constants.  On the device they live in constant memory: a binary64 literal cannot be an immediate operand,
// so every `fma(p, t, 1.0 / 5040.0)` otherwise costs two MOVs to build the constant — 100 of the ~340
// instructions per frame of the moving-cutoff chunk were such moves.  DFMA takes a constant-bank operand directly.
enum {
    S2K_INVFACT = 0,          // 1/k!, k = 0..17
    S2K_LN2 = 18, S2K_LOG2E, S2K_LN2_HI, S2K_LN2_LO, S2K_TWO_OVER_PI, S2K_PIO2_1, S2K_PIO2_1T,
    S2K_COUNT
};
#define S2_MATH_TABLE                                                                                              \
    {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0,    \
     1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0, 1.0 / 87178291200.0,                  \
     1.0 / 1307674368000.0, 1.0 / 20922789888000.0, 1.0 / 355687428096000.0,                                         \
     0.693147180559945309417232121458, 1.44269504088896340735992468100, 6.93147180369123816490e-01,                  \
     1.90821492927058770002e-10, 0.636619772367581343075535053490, 1.57079632673412561417e+00,                       \
     6.07710050650619224932e-11}
#ifdef __CUDA_ARCH__
static __constant__ double s2_math_k[S2K_COUNT] = S2_MATH_TABLE;
#else
static const double s2_math_k[S2K_COUNT] = S2_MATH_TABLE;
#endif
#define S2K(i) s2_math_k[i]
#define S2F(k) s2_math_k[S2K_INVFACT + (k)]       /* 1/k! */

// rint(x) and (int)rint(x) for |x| < 2^31 without the conversion unit: adding 1.5 * 2^52 leaves the rounded integer
// (round-to-nearest-even, the same as rint) in the low mantissa bits.  FRND.F64 and F2I.F64 run on the XU pipe
// at 1/8 rate (tools/ubench/dfma_rate.cu), and that pipe is what bounds the moving-cutoff chunk.
S2_HD double s2_rint_int(double x, int* k) {
    const double magic = 6755399441055744.0;
    const double t = x + magic;
    int64_t bits;
    memcpy(&bits, &t, 8);
    *k = (int)(uint32_t)(uint64_t)bits;
    return t - magic;
}

// e^t for |t| <= 0.35 (Taylor, degree 14: truncation < 1e-19)
S2_HD double s2_exp_kernel(double t) {
    double p = S2F(14);
    p = fma(p, t, S2F(13));
    p = fma(p, t, S2F(12));
    p = fma(p, t, S2F(11));
    p = fma(p, t, S2F(10));
    p = fma(p, t, S2F(9));
    p = fma(p, t, S2F(8));
    p = fma(p, t, S2F(7));
    p = fma(p, t, S2F(6));
    p = fma(p, t, S2F(5));
    p = fma(p, t, S2F(4));
    p = fma(p, t, S2F(3));
    p = fma(p, t, S2F(2));
    p = fma(p, t, S2F(1));
    p = fma(p, t, S2F(0));
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
    int k;
    const double kd = s2_rint_int(xd, &k);
    const double f = xd - kd;                             // exact
    const double r = s2_exp_kernel(f * S2K(S2K_LN2));
    return (float)s2_scale2(r, k);                        // one rounding (subnormal results round here too)
}

// 2^x in binary64 for a binary32 x in (-150, 128] (the centre of a moving-cutoff window, s2_cutoff.h)
S2_HD double s2_exp2_d(float x) {
    const double xd = (double)x;
    int k;
    const double kd = s2_rint_int(xd, &k);
    return s2_scale2(s2_exp_kernel((xd - kd) * S2K(S2K_LN2)), k);
}

// e^x, x binary32.  Cody-Waite reduction x = k*ln2 + r, |r| <= 0.35.
S2_HD float s2_expf(float x) {
    if (!(x > -104.0f)) return x != x ? x : 0.0f;
    if (x > 89.0f) return INFINITY;
    const double xd = (double)x;
    int k;
    const double kd = s2_rint_int(xd * S2K(S2K_LOG2E), &k);
    const double r = fma(-kd, S2K(S2K_LN2_LO), fma(-kd, S2K(S2K_LN2_HI), xd));
    return (float)s2_scale2(s2_exp_kernel(r), k);
}

// sin and cos of x (|x| < ~1e5) in binary64 — one reduction, two Taylor kernels on |r| <= pi/4.
S2_HD void s2_sincos_d(double xd, double* s, double* c) {
    int qi;
    const double qd = s2_rint_int(xd * S2K(S2K_TWO_OVER_PI), &qi);          // x * 2/pi
    // pi/2 = pio2_1 + pio2_1t (+ 2e-21): pio2_1 has 33 significant bits, so qd * pio2_1 is exact
    double r = fma(-qd, S2K(S2K_PIO2_1), xd);
    r = fma(-qd, S2K(S2K_PIO2_1T), r);
    const double z = r * r;
    double ps = -S2F(17);
    ps = fma(ps, z, S2F(15));
    ps = fma(ps, z, -S2F(13));
    ps = fma(ps, z, S2F(11));
    ps = fma(ps, z, -S2F(9));
    ps = fma(ps, z, S2F(7));
    ps = fma(ps, z, -S2F(5));
    ps = fma(ps, z, S2F(3));
    const double sr = fma(-ps * z, r, r);                 // r - r^3 * (1/6 - ...)
    double pc = S2F(16);
    pc = fma(pc, z, -S2F(14));
    pc = fma(pc, z, S2F(12));
    pc = fma(pc, z, -S2F(10));
    pc = fma(pc, z, S2F(8));
    pc = fma(pc, z, -S2F(6));
    pc = fma(pc, z, S2F(4));
    pc = fma(pc, z, -S2F(2));
    const double cr = fma(pc, z, S2F(0));
    const int q = qi & 3;
    const double sv = (q & 1) ? cr : sr;
    const double cv = (q & 1) ? sr : cr;
    *s = (q & 2) ? -sv : sv;
    *c = ((q + 1) & 2) ? -cv : cv;
}

// sin and cos of a binary32 x, each rounded once
S2_HD void s2_sincosf(float x, float* s, float* c) {
    double sd, cd;
    s2_sincos_d((double)x, &sd, &cd);
    *s = (float)sd;
    *c = (float)cd;
}

// tan of a binary32 x: sin / cos in binary64 (relative error ~3e-16), rounded once
S2_HD float s2_tanf(float x) {
    double sd, cd;
    s2_sincos_d((double)x, &sd, &cd);
    return (float)(sd / cd);
}
