// s2_device.cuh — device-side building blocks shared by the render kernels (s2_kernels.cu: one warp per
// voice group; s2_kernel_pc.cu: producer/consumer warp pair per voice group).
//
// Arithmetic contract: every operation below that feeds a discrete decision (phase, table index,
// envelope stage, noise hash) is the reference's binary32 operation, spelled with __f*_rn intrinsics so
// nothing is contracted or reassociated (the files are also built with -fmad=false).
// Reference line numbers are relative to /root/reference/components/s2_lib/src/.
#pragma once

#include "s2_internal.h"
#include "s2_math.h"

// The biquad's products and sums are rounded separately, as the source reads (dsp_filters.rs:116-128).  A
// fused (FFMA) form was tried and rejected: a resonant low-cutoff biquad in f32 direct form amplifies
// per-frame rounding differences by ~1/(1-r) (~10^3 at 100 Hz, damping 0.2), enough to pass the 1e-4 bar.
// (The fast phase step was also tried as a separate pass ahead of everything else: slower, row d of
// profiles/r1_notes.md.)

// Frames per straight-line trip of the time-packed loop; 16 and 32 measured no faster than 8.
#ifndef S2_TRIP
#define S2_TRIP 8
#endif

namespace s2 {


static __device__ const uint32_t d_sin_bits[1024] = {
#include "sin_table_bits.inc"
};

// ------------------------------------------------------------------------------------------
// Per-lane voice description, decoded once per launch.

struct EnvP {
    float A, AD, S, Rs, E;   // stage boundaries in samples: attack end, decay end, release start/end
    float D, R;              // decay / release lengths in samples (scalar tail path)
    float relf;              // release_offset as f32 (None -> u32::MAX as f32), before the max()
    float sA, sD, sR;        // hoisted slopes rise/run of the x16 envelope (old/simdtest.rs:247-251)
};

struct OscC {                // everything derived from the period; hoisting a division whose
    float P, d;              // operands do not change is exact
    float slope;             // -2 / P            (saw,      try3/oscillators.rs:99-119)
    float half;              // P / 2             (square,   try3/oscillators.rs:60-80)
    float ts1, ts2;          // -2 / half, 2 / half (triangle, try3/oscillators.rs:148-183)
    uint32_t fo_bits;        // frequency these were derived from
};

struct FiltC {               // one-pole: c0 = k, c1 = 1 - k.   biquad: c0 = 2*alpha, c1 = 2*beta, c2 = 2*gamma
    float c0, c1, c2;
    uint32_t fl_bits;
};

struct FiltS { float x1, x2, y1, y2; };   // one-pole keeps `last` in y1

struct Lane {
    uint32_t kind, rot;      // rot = seed.rotate_left(5)  (try3/hashnoise.rs:53-55)
    float pitch, gain, namt, lpf, damp, amt_osc, amt_lpf;
    EnvP amp, mod;
};

// units.rs:44-53
__device__ __forceinline__ float ms_as_samples(float ms, float sr) {
    return __fmul_rn(sr, __fdiv_rn(ms, 1000.0f));
}

__device__ __forceinline__ void make_env(EnvP& e, float a_ms, float d_ms, float s, float r_ms,
                                         uint32_t release, float sr) {
    e.A = ms_as_samples(a_ms, sr);
    e.D = ms_as_samples(d_ms, sr);
    e.R = ms_as_samples(r_ms, sr);
    e.S = s;
    e.AD = __fadd_rn(e.A, e.D);
    e.relf = __uint2float_rn(release);             // unwrap_or(u32::MAX) as f32 (simdtest.rs:283)
    e.Rs = fmaxf(e.relf, e.AD);                    // simd_max (simdtest.rs:285)
    e.E = __fadd_rn(e.Rs, e.R);
    e.sA = __fdiv_rn(1.0f, e.A);
    e.sD = __fdiv_rn(__fsub_rn(s, 1.0f), e.D);
    e.sR = __fdiv_rn(-s, e.R);
}

// old/simdtest.rs:287-291: the mask chain, as a stage index
template <class E>
__device__ __forceinline__ int env_stage(const E& e, float x) {
    return x < e.A ? 0 : (x < e.AD ? 1 : (x < e.Rs ? 2 : (x < e.E ? 3 : 4)));
}

// old/simdtest.rs:270-330 for one lane; line = (rise/run)*x + y0, never fused (:247-261)
template <class E>
__device__ __forceinline__ float env_x16(const E& e, float x) {
    switch (env_stage(e, x)) {
    case 0: return __fadd_rn(__fmul_rn(e.sA, x), 0.0f);
    case 1: return __fadd_rn(__fmul_rn(e.sD, __fsub_rn(x, e.A)), 1.0f);
    case 2: return e.S;
    case 3: return __fadd_rn(__fmul_rn(e.sR, __fsub_rn(x, e.Rs)), e.S);
    default: return 0.0f;
    }
}

// The envelope segment a frame offset lies in, as the line g = es * (x - ex0) + ey0 that reproduces that
// stage's formula bit-for-bit (attack: (1/A)*x + 0; decay: ((S-1)/D)*(x-A) + 1; sustain: 0*x + S; release:
// ((-S)/R)*(x-Rs) + S; end: 0*x + 0), valid for frame offsets [.., nend).  Offsets below 2^24 only
// (x = (f32)n exact; nend = first integer whose f32 image reaches the stage boundary).
struct SegEnv { float es, nex0, ey0; uint32_t nend; };

template <class E>
__device__ __forceinline__ SegEnv seg_env(const E& e, uint32_t n) {
    const float x = __uint2float_rn(n);
    const int st = env_stage(e, x);
    SegEnv s;
    s.es = st == 0 ? e.sA : st == 1 ? e.sD : st == 3 ? e.sR : 0.0f;
    s.nex0 = st == 1 ? -e.A : st == 3 ? -e.Rs : -0.0f;             // x + (-0.0) == x
    s.ey0 = st == 1 ? 1.0f : (st == 2 || st == 3) ? e.S : 0.0f;
    const float b = st == 0 ? e.A : st == 1 ? e.AD : st == 2 ? e.Rs : st == 3 ? e.E : 4.0e9f;
    s.nend = min(__float2uint_ru(b), 1u << 24);
    return s;
}
__device__ __forceinline__ float seg_eval(const SegEnv& s, float x) {
    return __fadd_rn(__fmul_rn(s.es, __fadd_rn(x, s.nex0)), s.ey0);
}

// fmod(t, 1) for t in [0, 2) (the fast paths: 0 <= phase < 1, 0 <= 1/P < 1): t - 1 is exact there, and as
// unsigned integers the bits of a negative t - 1 exceed those of any t < 1 while a non-negative t - 1 lies
// below t >= 1, so the wrap is one integer min.  No predicate: FSETP/ISETP -> FSEL or a predicated FADD costs
// 14 cycles of latency on this part against 6.75 for VIMNMX (tools/ubench/phase_chain.cu: 22.9 -> 14.75 cycles
// per step of the recurrence, same bits).
__device__ __forceinline__ float wrap_unit(float t) {
    return __uint_as_float(min(__float_as_uint(__fadd_rn(t, -1.0f)), __float_as_uint(t)));
}

// math.rs:11-19 with feature "fma"
__device__ __forceinline__ float line_fma(float rise, float run, float x, float y0) {
    return __fmaf_rn(__fdiv_rn(rise, run), x, y0);
}

// try3/envelopes.rs:22-149 (tail frames only)
__device__ __forceinline__ float env_scalar(const EnvP& e, float x) {
    const float rel = e.relf;
    const float end = __fadd_rn(rel, e.R);
    const bool in_release = x >= rel && x < end;
    const bool in_end = x >= end;
    const bool in_attack = !in_release && !in_end && x < e.A;
    const bool in_decay = !in_release && !in_end && !in_attack && x < e.AD;
    const bool in_sustain = !in_release && !in_end && !in_attack && !in_decay && x < rel;
    float rss;
    if (rel < e.A) rss = line_fma(1.0f, e.A, rel, 0.0f);
    else if (rel < e.AD) rss = line_fma(__fsub_rn(e.S, 1.0f), e.D, __fsub_rn(rel, e.A), 1.0f);
    else rss = e.S;
    if (in_attack) return line_fma(1.0f, e.A, x, 0.0f);
    if (in_decay) return line_fma(__fsub_rn(e.S, 1.0f), e.D, __fsub_rn(x, e.A), 1.0f);
    if (in_sustain) return e.S;
    if (in_release) return line_fma(-rss, e.R, __fsub_rn(x, rel), rss);
    return 0.0f;
}

// 2^x, e^x, sin/cos: s2_math.h (binary64 evaluation, one rounding).  The reference calls sleef pow
// (x16, process.rs:244) / libm powf (scalar, process.rs:227) / expf / sinf / cosf, none reproducible
// bit-for-bit on a GPU; their outputs only feed float results, compared with the north-star tolerance.
__device__ __forceinline__ float pow2_ref(float x) { return s2_exp2f(x); }
__device__ __forceinline__ float exp_ref(float x) { return s2_expf(x); }

// process.rs:231-250.  amount == 0 -> pow(2, +-0) == 1 and 1 * f == f exactly: skip the call.
__device__ __forceinline__ float modulate_freq(float f, float m, float amount) {
    if (amount == 0.0f) return f;
    return __fmul_rn(pow2_ref(__fmul_rn(m, amount)), f);
}

__device__ __forceinline__ void make_osc(OscC& o, float fo, float sr) {
    o.fo_bits = __float_as_uint(fo);
    o.P = __fdiv_rn(sr, fo);                       // units.rs:32-41
    o.d = __fdiv_rn(1.0f, o.P);                    // try3/oscillators.rs:378
    o.slope = __fdiv_rn(-2.0f, o.P);
    o.half = __fdiv_rn(o.P, 2.0f);
    o.ts1 = __fdiv_rn(-2.0f, o.half);
    o.ts2 = __fdiv_rn(2.0f, o.half);
}

// FILTER template values = S2_FILTER_* of include/s2_cuda.h
enum { FILT_ONE_POLE = 0, FILT_BIQUAD_LP = 1, FILT_BIQUAD_HP = 2, FILT_BIQUAD_BP = 3, FILT_FIRST_LP = 4, FILT_FIRST_HP = 5 };

template <int FILTER>
__device__ __forceinline__ void make_filt(FiltC& c, float fl, float damp, float sr) {
    c.fl_bits = __float_as_uint(fl);
    const float pi = 3.14159274101257324219f;
    if (FILTER == 0) {
        // try3/filters.rs:21: (-2.0 * pi * freq / sample_rate).exp()
        float t = __fmul_rn(-2.0f, pi);
        t = __fmul_rn(t, fl);
        t = __fdiv_rn(t, sr);
        const float k = exp_ref(t);
        c.c0 = k;
        c.c1 = __fsub_rn(1.0f, k);
        c.c2 = 0.0f;
    } else if (FILTER == FILT_BIQUAD_LP || FILTER == FILT_BIQUAD_HP) {
        // try3/dsp_filters.rs:99-109 (low-pass), :149-159 (high-pass: alpha = (1/2 + beta + gamma) / 4)
        float th = __fmul_rn(2.0f, pi);
        th = __fmul_rn(th, fl);
        th = __fdiv_rn(th, sr);
        float s, co;
        s2_sincosf(th, &s, &co);
        const float hd = __fmul_rn(damp, 0.5f);                      // damp / 2.0: scaling by a power of two, same bits
        const float num = __fsub_rn(1.0f, __fmul_rn(hd, s));
        const float den = __fadd_rn(1.0f, __fmul_rn(hd, s));
        const float beta = __fmul_rn(0.5f, __fdiv_rn(num, den));
        const float gamma = __fmul_rn(__fadd_rn(0.5f, beta), co);
        const float hb = __fadd_rn(0.5f, beta);
        const float alpha = __fmul_rn(FILTER == FILT_BIQUAD_LP ? __fsub_rn(hb, gamma) : __fadd_rn(hb, gamma), 0.25f);   // ... / 4.0, likewise
        // y = 2*(alpha*s + gamma*y1 - beta*y2): scaling by 2 commutes with round-to-nearest, so the
        // doubling is folded into the coefficients (exact unless an intermediate is subnormal).
        c.c0 = __fmul_rn(2.0f, alpha);
        c.c1 = __fmul_rn(2.0f, beta);
        c.c2 = __fmul_rn(2.0f, gamma);
    } else if (FILTER == FILT_BIQUAD_BP) {
        // try3/dsp_filters.rs:197-207; `damp` carries the quality factor
        float th = __fmul_rn(2.0f, pi);
        th = __fmul_rn(th, fl);
        th = __fdiv_rn(th, sr);
        const float tn = s2_tanf(__fdiv_rn(th, __fmul_rn(2.0f, damp)));
        const float beta = __fmul_rn(0.5f, __fdiv_rn(__fsub_rn(1.0f, tn), __fadd_rn(1.0f, tn)));
        float s, co;
        s2_sincosf(th, &s, &co);
        const float gamma = __fmul_rn(__fadd_rn(0.5f, beta), co);
        const float alpha = __fmul_rn(__fsub_rn(0.5f, beta), 0.5f);
        c.c0 = __fmul_rn(2.0f, alpha);
        c.c1 = __fmul_rn(2.0f, beta);
        c.c2 = __fmul_rn(2.0f, gamma);
    } else {
        // try3/dsp_filters.rs:30-32 (first-order low-pass) / :64-66 (high-pass): c0 = alpha, c2 = gamma
        float th = __fmul_rn(2.0f, pi);
        th = __fmul_rn(th, fl);
        th = __fdiv_rn(th, sr);
        float s, co;
        s2_sincosf(th, &s, &co);
        const float gamma = __fdiv_rn(co, __fadd_rn(1.0f, s));
        c.c0 = __fmul_rn(FILTER == FILT_FIRST_LP ? __fsub_rn(1.0f, gamma) : __fadd_rn(1.0f, gamma), 0.5f);
        c.c1 = 0.0f;
        c.c2 = gamma;
    }
}

// The compact envelope the classifier and the per-frame evaluation need (env_stage / env_x16); the
// scalar-tail envelope (env_scalar) needs the full EnvP.
struct EnvQ { float A, AD, S, Rs, E, sA, sD, sR; };
__device__ __forceinline__ EnvQ compact(const EnvP& e) { return {e.A, e.AD, e.S, e.Rs, e.E, e.sA, e.sD, e.sR}; }

// Decode one voice's parameter column (struct-of-arrays, see s2_internal.h).
__device__ __forceinline__ Lane load_lane(const float* __restrict__ P, uint32_t vp, float sr) {
    Lane L;
    L.kind = __float_as_uint(P[P_KIND * vp]);
    const uint32_t seed = __float_as_uint(P[P_SEED * vp]);
    L.rot = (seed << 5) | (seed >> 27);
    L.pitch = P[P_PITCH * vp];
    L.gain = P[P_GAIN * vp];
    L.namt = P[P_NOISE * vp];
    L.lpf = P[P_LPF * vp];
    L.damp = P[P_DAMP * vp];
    L.amt_osc = P[P_AMT_OSC * vp];
    L.amt_lpf = P[P_AMT_LPF * vp];
    const uint32_t release = __float_as_uint(P[P_RELEASE * vp]);
    make_env(L.amp, P[P_AA * vp], P[P_AD * vp], P[P_AS * vp], P[P_AR * vp], release, sr);
    make_env(L.mod, P[P_MA * vp], P[P_MD * vp], P[P_MS * vp], P[P_MR * vp], release, sr);
    return L;
}

// ------------------------------------------------------------------------------------------
// One frame of one voice.

// phased + basic oscillators (try3/oscillators.rs:217-239 then :60-199) and the phase step
// (:377-381).  LITERAL keeps both `%`; the fast form drops them where they are provably no-ops:
//   * RN(P * phase) < P for every phase < 1 (P - P*2^-24 lies more than half an ulp below P),
//     so `offset % period` returns its argument;
//   * phase + 1/P < 2 when 1/P < 1, so `% 1.0` is a conditional exact subtraction.
template <int KIND, bool LITERAL>
__device__ __forceinline__ float osc_step(uint32_t kind, const OscC& o, float& ph, const float* sintab) {
    float x = __fmul_rn(o.P, ph);                  // period.mul_add(phase, 0.0)
    if (LITERAL) x = fmodf(x, o.P);
    const uint32_t k = KIND >= 0 ? (uint32_t)KIND : kind;
    float y;
    if (k == 1u) {                                 // Saw
        y = __fmaf_rn(o.slope, x, 1.0f);
    } else if (k == 0u) {                          // Square
        y = x < o.half ? 1.0f : -1.0f;
    } else if (k == 2u) {                          // Triangle
        const float a = __fmaf_rn(o.ts1, x, 1.0f);
        const float b = __fmaf_rn(o.ts2, __fsub_rn(x, o.half), -1.0f);
        y = x < o.half ? a : b;
    } else {                                       // Sine: try3/lookup.rs:46-85 on SIN_TABLE
        const float tv = __fdiv_rn(__fmul_rn(x, 1024.0f), o.P);
        const uint32_t i1 = __float2uint_rz(tv);   // `as u32`: truncating, saturating
        const uint32_t i2 = (i1 + 1u) & 1023u;
        const float s1 = i1 < 1024u ? sintab[i1] : 0.0f;   // gather_or_default
        const float s2 = sintab[i2];
        y = __fmaf_rn(__fsub_rn(s2, s1), __fsub_rn(tv, __uint2float_rn(i1)), s1);
    }
    const float t = __fadd_rn(ph, o.d);
    if (LITERAL) ph = fmodf(t, 1.0f);
    else ph = wrap_unit(t);
    return y;
}

// try3/hashnoise.rs:33-68.  value / 65535 is replaced by fma(v, hi, v*lo) with hi + lo = 1/65535
// to 48 bits: equal to the IEEE quotient for all 65,536 possible values (tests/test_host_logic.py).
// (q * 2) - 1 is one fma because q * 2 is exact.
__device__ __forceinline__ float noise_fast(uint32_t rot, uint32_t n) {
    const uint32_t h = (rot ^ n) * 0x9e3779b9u;
    const float v = __uint2float_rn(h & 0xffffu);
    const float q = __fmaf_rn(v, 0x1.0001p-16f, __fmul_rn(v, 0x1.0001p-48f));
    return __fmaf_rn(q, 2.0f, -1.0f);
}

__device__ __forceinline__ float noise_literal(uint32_t rot, uint32_t n) {
    const uint32_t off = __float2uint_rz(__uint2float_rn(n));   // u32 -> f32 -> u32 (process.rs:347-348)
    const uint32_t h = (rot ^ off) * 0x9e3779b9u;
    const float v = __uint2float_rn(h & 0xffffu);
    const float q = __fdiv_rn(v, 65535.0f);
    return __fsub_rn(__fmul_rn(q, 2.0f), 1.0f);
}

template <int FILTER>
__device__ __forceinline__ float filt_step(float u, const FiltC& c, FiltS& s) {
    if (FILTER == 0) {
        // try3/filters.rs:23-33: a0.mul_add(input, -b1 * last), b1 = -k
        const float y = __fmaf_rn(c.c1, u, __fmul_rn(c.c0, s.y1));
        s.y1 = y;
        return y;
    } else if (FILTER == FILT_BIQUAD_LP || FILTER == FILT_BIQUAD_HP || FILTER == FILT_BIQUAD_BP) {
        // try3/dsp_filters.rs:116-128: 2*(alpha*(x + 2*x1 + x2) + gamma*y1 - beta*y2); x + 2*x1 is one fma
        // because 2*x1 is exact.  :166-176 (high-pass) has x - 2*x1 + x2, :214-224 (band-pass) x - x2.
        float sx;
        if (FILTER == FILT_BIQUAD_LP) sx = __fadd_rn(__fmaf_rn(2.0f, s.x1, u), s.x2);
        else if (FILTER == FILT_BIQUAD_HP) sx = __fadd_rn(__fmaf_rn(-2.0f, s.x1, u), s.x2);
        else sx = __fsub_rn(u, s.x2);
        float t = __fmul_rn(c.c0, sx);
        t = __fadd_rn(t, __fmul_rn(c.c2, s.y1));
        t = __fsub_rn(t, __fmul_rn(c.c1, s.y2));
        s.x2 = s.x1; s.x1 = u; s.y2 = s.y1; s.y1 = t;
        return t;
    } else {
        // try3/dsp_filters.rs:37-41 / :71-75: alpha * (x +- x1) + gamma * y1 (no mul_add in the source)
        const float xs = FILTER == FILT_FIRST_LP ? __fadd_rn(u, s.x1) : __fsub_rn(u, s.x1);
        const float t = __fadd_rn(__fmul_rn(c.c0, xs), __fmul_rn(c.c2, s.y1));
        s.x1 = u; s.y1 = t;
        return t;
    }
}

// ------------------------------------------------------------------------------------------
// Lane-vector arithmetic.  A lane carries NV voices (NV = 1: one voice, scalar FP32 instructions;
// NV = 2: two voices packed in a float2 and computed with Blackwell's packed-FP32 instructions
// FADD2 / FMUL2 / FFMA2, which retire two IEEE-754 round-to-nearest results per issue slot — the
// render loop is issue-bound, not FLOP-bound: profiles/r1_notes.md).  Each element is rounded
// exactly like the scalar instruction, so parity is unchanged.

template <int NV> struct VT;
template <> struct VT<1> { using type = float; };
template <> struct VT<2> { using type = float2; };
template <int NV> using vf = typename VT<NV>::type;

__device__ __forceinline__ float vget(float v, int) { return v; }
__device__ __forceinline__ float vget(float2 v, int e) { return e ? v.y : v.x; }
__device__ __forceinline__ void vset(float& v, int, float x) { v = x; }
__device__ __forceinline__ void vset(float2& v, int e, float x) { if (e) v.y = x; else v.x = x; }
template <int NV> __device__ __forceinline__ vf<NV> vsplat(float x);
template <> __device__ __forceinline__ float vsplat<1>(float x) { return x; }
template <> __device__ __forceinline__ float2 vsplat<2>(float x) { return make_float2(x, x); }

// CONTRACTION HAZARD (ptxas 12.9, sm_100a): a packed multiply whose result feeds a packed add is fused
// into FFMA2 — with the __fmul2_rn/__fadd2_rn builtins AND with explicit `mul.rn.f32x2` / `add.rn.f32x2`
// PTX, -fmad=false notwithstanding (tools/ubench/fuse_check.cu; scalar __fmul_rn + __fadd_rn is not
// fused).  That moves results by an ulp and can flip a square wave's sign.  Rule used in this file: the
// result of pmul2 never feeds padd2; where the reference adds to a product, the add is done with scalar
// __fadd_rn per element (vadd(float2, float2) below always is).
__device__ __forceinline__ float2 padd2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 pmul2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 pfma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

__device__ __forceinline__ float vadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float vmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float vfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return pmul2(a, b); }
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return pfma2(a, b, c); }

// Per-lane constants and state of the fast path, NV voices wide.
template <int NV> struct FastV {
    // oscillator (derived from the period; negations are stored so the loop only adds)
    vf<NV> P, d, slope, nhalf, ts1, ts2;
    // patch
    vf<NV> gain, namt;
    // filter: one-pole c0 = k, c1 = 1 - k; biquad c0 = 2*alpha, nc1 = -2*beta, c2 = 2*gamma
    vf<NV> c0, c1, c2;
    // envelope segment g = es * (x + nex0) + ey0
    vf<NV> es, nex0, ey0;
    // carried state
    vf<NV> ph, x1, x2, y1, y2;
};

struct FastEnv { float es, ex0, ey0; };   // g = es * (x - ex0) + ey0 reproduces each stage bit-exactly

// Fast chunk: period, cutoff and envelope segment are constant over the 32 frames of every voice
// of the warp.  KIND >= 0: every voice of the warp runs that oscillator (banks are sorted by kind).
template <int NV, int FILTER, int KIND, bool GCONST, bool NAMT0, int TRACE>
__device__ __forceinline__ void chunk_fast(FastV<NV>& F, const uint32_t (&kind)[NV], const uint32_t (&rot)[NV],
                                           const uint32_t (&n0)[NV], float* __restrict__ tile, int lane,
                                           const float* sintab) {
    const vf<NV> one = vsplat<NV>(1.0f), none = vsplat<NV>(-1.0f), two = vsplat<NV>(2.0f);
    uint32_t n[NV];
    vf<NV> xf;
#pragma unroll
    for (int e = 0; e < NV; e++) { n[e] = n0[e]; vset(xf, e, __uint2float_rn(n0[e])); }   // exact: n0 + 32 <= 2^24
    // 8 frames per trip: long enough for the scheduler to overlap neighbouring frames, short enough to
    // live in the instruction cache.
#pragma unroll 2
    for (int j = 0; j < kChunk / 4; j++) {
        float o4[NV][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const vf<NV> ph0 = F.ph;
            // ---- oscillator: x = period.mul_add(phase, 0); `% period` is a no-op (see osc_step)
            const vf<NV> x = vmul(F.P, ph0);
            vf<NV> osc;
            if (KIND == 1) {                                       // Saw: fma(-2/P, x, 1)
                osc = vfma(F.slope, x, one);
            } else if (KIND == 0) {                                // Square: x < P/2 ? 1 : -1
                // sign(x - half) picks +-1: (x - half) is -0 never, +0 when equal -> -1 like `<`
                const vf<NV> dl = vadd(x, F.nhalf);
#pragma unroll
                for (int e = 0; e < NV; e++)
                    vset(osc, e, __uint_as_float((__float_as_uint(vget(dl, e)) & 0x80000000u) ^ 0xbf800000u));
            } else if (KIND == 2) {                                // Triangle
                const vf<NV> dl = vadd(x, F.nhalf);
                const vf<NV> a = vfma(F.ts1, x, one);
                const vf<NV> b = vfma(F.ts2, dl, none);
#pragma unroll
                for (int e = 0; e < NV; e++) vset(osc, e, vget(dl, e) < 0.0f ? vget(a, e) : vget(b, e));
            } else {                                               // Sine, or a warp of mixed kinds
#pragma unroll
                for (int e = 0; e < NV; e++) {
                    const uint32_t k = KIND == 3 ? 3u : kind[e];
                    const float xe = vget(x, e), Pe = vget(F.P, e), he = -vget(F.nhalf, e);
                    float y;
                    if (k == 1u) y = __fmaf_rn(vget(F.slope, e), xe, 1.0f);
                    else if (k == 0u) y = xe < he ? 1.0f : -1.0f;
                    else if (k == 2u) {
                        const float a = __fmaf_rn(vget(F.ts1, e), xe, 1.0f);
                        const float b = __fmaf_rn(vget(F.ts2, e), __fsub_rn(xe, he), -1.0f);
                        y = xe < he ? a : b;
                    } else {                                       // try3/lookup.rs:46-85 on SIN_TABLE
                        const float tv = __fdiv_rn(__fmul_rn(xe, 1024.0f), Pe);
                        const uint32_t i1 = __float2uint_rz(tv);
                        const uint32_t i2 = (i1 + 1u) & 1023u;
                        const float s1 = i1 < 1024u ? sintab[i1] : 0.0f;
                        const float s2 = sintab[i2];
                        y = __fmaf_rn(__fsub_rn(s2, s1), __fsub_rn(tv, __uint2float_rn(i1)), s1);
                    }
                    vset(osc, e, y);
                }
            }
            // ---- phase step: t = phase + 1/P; `% 1.0` == subtract 1 when t >= 1 (t < 2, exact)
            const vf<NV> t = vadd(ph0, F.d);
            vf<NV> w;
#pragma unroll
            for (int e = 0; e < NV; e++) vset(w, e, vget(t, e) >= 1.0f ? 1.0f : 0.0f);
            F.ph = vfma(w, none, t);                               // t - w, exact product
            // ---- noise (try3/hashnoise.rs:33-68): integer hash, then v/65535*2-1 (see noise_fast)
            vf<NV> v;
#pragma unroll
            for (int e = 0; e < NV; e++) {
                const uint32_t h = (rot[e] ^ n[e]) * 0x9e3779b9u;
                vset(v, e, __uint2float_rn(h & 0xffffu));
                n[e] += 1u;
            }
            const vf<NV> q = vfma(v, vsplat<NV>(0x1.0001p-16f), vmul(v, vsplat<NV>(0x1.0001p-48f)));
            const vf<NV> nz = vfma(q, two, none);
            // ---- process.rs:341-358: gain and noise amount are ADDED on the x16 path.
            // nz + 0.0 == nz bit-for-bit (nz is never -0.0), so NAMT0 drops that add.
            const vf<NV> u = vadd(vadd(osc, F.gain), NAMT0 ? nz : vadd(nz, F.namt));
            // ---- filter
            vf<NV> y;
            if (FILTER == 0) {
                // try3/filters.rs:23-33: a0.mul_add(input, k * last)
                y = vfma(F.c1, u, vmul(F.c0, F.y1));
                F.y1 = y;
            } else {
                // try3/dsp_filters.rs:116-128 (see filt_step): 2*(alpha*(x + 2*x1 + x2) + gamma*y1 - beta*y2)
                vf<NV> sx = vfma(two, F.x1, u);
                sx = vadd(sx, F.x2);
                vf<NV> tt = vmul(F.c0, sx);
                tt = vadd(tt, vmul(F.c2, F.y1));
                y = vadd(tt, vmul(F.c1, F.y2));                             // c1 holds -2*beta: exact negation
                F.x2 = F.x1; F.x1 = u; F.y2 = F.y1; F.y1 = y;
            }
            // ---- amp envelope (old/simdtest.rs:270-330 on one segment) and gain (process.rs:373-378)
            vf<NV> g;
            if (GCONST) g = F.ey0;
            else {
                g = vadd(vmul(F.es, vadd(xf, F.nex0)), F.ey0);
                xf = vadd(xf, one);
            }
            const vf<NV> out = TRACE == TRACE_PHASE ? ph0 : vmul(y, g);
#pragma unroll
            for (int e = 0; e < NV; e++) o4[e][i] = vget(out, e);
        }
#pragma unroll
        for (int e = 0; e < NV; e++)
            *reinterpret_cast<float4*>(tile + (e * 32 + lane) * kTileStride + 4 * j) =
                make_float4(o4[e][0], o4[e][1], o4[e][2], o4[e][3]);
    }
}

// Time-packed fast chunk for one voice per lane: the render loop is issue-bound, not FLOP-bound
// (profiles/r1_notes.md), and FADD2/FMUL2/FFMA2 retire two IEEE-754 results per issue slot.  The two
// recurrences (phase, filter) stay scalar — a packed op has twice the latency — while everything that
// is feed-forward (waveform, noise map, gain/noise combine, envelope, output gain) is computed for
// frames (i, i+1) of the voice in one packed instruction.  Element-wise rounding is identical.
// ALIGNED8: the voice's frame offset and its rotated seed are multiples of 8 (seeds below 2^27 rotate to multiples
// of 32), so (seed' ^ (offset + i)) == (seed' ^ offset) + i and the hash of frame i is one add with an immediate.
// GCONST = false: the amp envelope is evaluated per frame with its full stage chain (env_x16), so
// attack / decay / release ramps and their boundaries stay on the fast path.
template <int FILTER, int KIND, bool GCONST, bool NAMT0, bool ALIGNED8, int TRACE>
__device__ __forceinline__ void chunk_fast_tp(FastV<1>& F, const EnvP* __restrict__ amp, uint32_t kind, uint32_t rot,
                                              uint32_t n0, float* __restrict__ row, const float* sintab) {
    const float2 one2 = make_float2(1.0f, 1.0f), none2 = make_float2(-1.0f, -1.0f), two2 = make_float2(2.0f, 2.0f);
    const float2 P2 = make_float2(F.P, F.P), slope2 = make_float2(F.slope, F.slope);
    const float2 ts1_2 = make_float2(F.ts1, F.ts1), ts2_2 = make_float2(F.ts2, F.ts2);
    const float2 gain2 = make_float2(F.gain, F.gain), namt2 = make_float2(F.namt, F.namt);
    const float2 ey0_2 = make_float2(F.ey0, F.ey0);
    const float hbig = __fmul_rn(-F.nhalf, 0x1p60f);               // (P / 2) * 2^60, exact
    uint32_t n = n0;
    float xf = __uint2float_rn(n0);                               // exact: n0 + 32 <= 2^24
    EnvP A;
    SegEnv sg = {0.0f, 0.0f, 0.0f, 0u};
    uint32_t ne = n0;                                             // frame offset of the next envelope pair
    if (!GCONST) { A = *amp; sg = seg_env(A, n0); }
    FiltS fs = {F.x1, F.x2, F.y1, F.y2};
    FiltC fc;
    fc.c0 = F.c0; fc.c1 = FILTER == 0 ? F.c1 : -F.c1; fc.c2 = F.c2; fc.fl_bits = 0;   // F.c1 holds -2*beta for the biquad
    float ph = F.ph;
    // 8 frames per trip: long enough to overlap neighbouring frames, short enough for the instruction cache
#pragma unroll 1
    for (int jt = 0; jt < kChunk / S2_TRIP; jt++) {
    const uint32_t nb = rot ^ n;                                  // hash input base of this trip
    // ALIGNED8 (offset and rot both multiples of 8): nb ^ i == nb + i for i < 8, so the hash of frame i is
    // nb * C + i * C: one multiply per trip and one add with an immediate per frame
    const uint32_t nbc = nb * 0x9e3779b9u;
#pragma unroll
    for (int jj = 0; jj < S2_TRIP / 4; jj++) {
        const int j = (S2_TRIP / 4) * jt + jj;
        float o4[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // ---- phase recurrence, two frames (try3/oscillators.rs:377-381; see osc_step)
            const float pa = ph;
            const float ta = __fadd_rn(pa, F.d);
            const float pb = wrap_unit(ta);
            const float tb = __fadd_rn(pb, F.d);
            ph = wrap_unit(tb);
            const float2 ph2 = make_float2(pa, pb);
            // ---- waveform: x = period.mul_add(phase, 0); `% period` is a no-op
            const float2 x2 = pmul2(P2, ph2);
            float2 osc2;
            if (KIND == 1) {
                osc2 = pfma2(slope2, x2, one2);
            } else if (KIND == 0) {
                // x < P/2 ? +1 : -1 without a compare or bit surgery: s = sat(2^60 * (P/2 - x)) is exactly 1 when
                // x < P/2 and exactly 0 otherwise (the fma is exact in sign; two distinct binary32 values of this
                // magnitude differ by >= 2^-24, so the product is >= 2^36 before the clamp; equality gives +0; a
                // NaN clamps to 0, i.e. -1, as `NaN < h` is false), and 2 s - 1 is exact.
                const float2 sq = make_float2(__saturatef(__fmaf_rn(x2.x, -0x1p60f, hbig)),
                                              __saturatef(__fmaf_rn(x2.y, -0x1p60f, hbig)));
                osc2 = pfma2(sq, two2, none2);
            } else if (KIND == 2) {
                const float2 dl = make_float2(__fadd_rn(x2.x, F.nhalf), __fadd_rn(x2.y, F.nhalf));
                const float2 a = pfma2(ts1_2, x2, one2);
                const float2 b = pfma2(ts2_2, dl, none2);
                osc2.x = dl.x < 0.0f ? a.x : b.x;
                osc2.y = dl.y < 0.0f ? a.y : b.y;
            } else {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t k = KIND == 3 ? 3u : kind;
                    const float xe = e ? x2.y : x2.x, he = -F.nhalf;
                    float y;
                    if (k == 1u) y = __fmaf_rn(F.slope, xe, 1.0f);
                    else if (k == 0u) y = xe < he ? 1.0f : -1.0f;
                    else if (k == 2u) {
                        const float a = __fmaf_rn(F.ts1, xe, 1.0f);
                        const float b = __fmaf_rn(F.ts2, __fsub_rn(xe, he), -1.0f);
                        y = xe < he ? a : b;
                    } else {                                       // try3/lookup.rs:46-85 on SIN_TABLE
                        const float tv = __fdiv_rn(__fmul_rn(xe, 1024.0f), F.P);
                        const uint32_t i1 = __float2uint_rz(tv);
                        const uint32_t i2 = (i1 + 1u) & 1023u;
                        const float s1 = i1 < 1024u ? sintab[i1] : 0.0f;
                        const float s2 = sintab[i2];
                        y = __fmaf_rn(__fsub_rn(s2, s1), __fsub_rn(tv, __uint2float_rn(i1)), s1);
                    }
                    if (e) osc2.y = y; else osc2.x = y;
                }
            }
            // ---- noise (try3/hashnoise.rs:33-68)
            const uint32_t fi = 4u * jj + 2u * h;                 // frame index inside the trip (compile-time)
            const uint32_t ha = ALIGNED8 ? nbc + fi * 0x9e3779b9u : (rot ^ (n + fi)) * 0x9e3779b9u;
            const uint32_t hb = ALIGNED8 ? nbc + (fi + 1u) * 0x9e3779b9u : (rot ^ (n + fi + 1u)) * 0x9e3779b9u;
            const float2 v2 = make_float2(__uint2float_rn(ha & 0xffffu), __uint2float_rn(hb & 0xffffu));
            const float2 q2 = pfma2(v2, make_float2(0x1.0001p-16f, 0x1.0001p-16f),
                                         pmul2(v2, make_float2(0x1.0001p-48f, 0x1.0001p-48f)));
            const float2 nz2 = pfma2(q2, two2, none2);
            // ---- process.rs:341-358 (ADD, x16 quirk); nz + 0.0 == nz bit-for-bit
            const float2 u2 = padd2(padd2(osc2, gain2), NAMT0 ? nz2 : padd2(nz2, namt2));
            // ---- filter recurrence, scalar
            const float ya = filt_step<FILTER>(u2.x, fc, fs);
            const float yb = filt_step<FILTER>(u2.y, fc, fs);
            // ---- envelope segment and output gain
            float2 g2;
            if (GCONST) g2 = ey0_2;
            else {
                const float xb = __fadd_rn(xf, 1.0f);
                if (ne + 2u <= sg.nend) {
                    // both frames inside the current segment: its line, bit-exact (see SegEnv)
                    g2.x = seg_eval(sg, xf);
                    g2.y = seg_eval(sg, xb);
                } else {
                    // a stage boundary: the full stage chain for these two frames, then the next segment
                    g2.x = env_x16(A, xf);
                    g2.y = env_x16(A, xb);
                    sg = seg_env(A, ne + 2u);
                }
                ne += 2u;
                xf = __fadd_rn(xf, 2.0f);
            }
            const float2 out2 = TRACE == TRACE_PHASE ? ph2 : pmul2(make_float2(ya, yb), g2);
            o4[2 * h] = out2.x;
            o4[2 * h + 1] = out2.y;
        }
        *reinterpret_cast<float4*>(row + 4 * j) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
    n += (uint32_t)S2_TRIP;
    }
    F.ph = ph;
    F.x1 = fs.x1; F.x2 = fs.x2; F.y1 = fs.y1; F.y2 = fs.y2;
}

// Modulated-cutoff chunk (one voice per lane): the period is constant but the mod envelope is moving,
// so the cutoff — and with it the filter coefficients — changes every frame (process.rs:148-152,
// 363-371; the first 200 ms of every note of the default patch, synth.rs:141-150).  Everything else
// keeps its fast form; envelopes are evaluated per frame with their full stage chain.
// SHARED: every voice of the warp has the same cutoff trajectory (same cutoff, damping, modulation amount, mod
// envelope and frame offset — a detune / pitch sweep of one patch, BASELINE config 5), so the 32 frames'
// coefficients were computed once, one frame per lane, into `ctab` ([c0 | c1 | c2 | fl bits][32], see
// modcut_coefficients): the same make_filt on the same inputs, hence the same bits, at 1/32 of the work.
template <int FILTER, int KIND, int TRACE, bool SHARED, class ENV>
__device__ __forceinline__ void chunk_modcut(FastV<1>& F, const ENV* __restrict__ amp, const ENV* __restrict__ mod,
                                             float lpf, float amt_lpf, float damp, float sr, FiltC& fc, uint32_t kind,
                                             uint32_t rot, uint32_t n0, float* __restrict__ row, const float* sintab,
                                             const float* __restrict__ ctab) {
    const ENV A = *amp;
    ENV M;
    if (!SHARED) M = *mod;
    SegEnv sa = seg_env(A, n0), sm = {0.0f, 0.0f, 0.0f, 0u};
    if (!SHARED) sm = seg_env(M, n0);
    OscC o;
    o.P = F.P; o.d = F.d; o.slope = F.slope; o.half = -F.nhalf; o.ts1 = F.ts1; o.ts2 = F.ts2; o.fo_bits = 0;
    FiltS fs = {F.x1, F.x2, F.y1, F.y2};
    float ph = F.ph;
    uint32_t n = n0;
    float xf = __uint2float_rn(n0);                               // exact: n0 + 32 <= 2^24
#pragma unroll 1
    for (int j = 0; j < kChunk / 4; j++) {
        float o4[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float g;
            if (n < sa.nend) g = seg_eval(sa, xf); else { g = env_x16(A, xf); sa = seg_env(A, n + 1u); }
            if (SHARED) {
                const int fi = 4 * j + i;
                fc.c0 = ctab[fi]; fc.c1 = ctab[32 + fi]; fc.c2 = ctab[64 + fi];
            } else {
                float m;
                if (n < sm.nend) m = seg_eval(sm, xf); else { m = env_x16(M, xf); sm = seg_env(M, n + 1u); }
                // branch-free on purpose: 2^(m * 0) * f == f exactly, and re-deriving unchanged coefficients
                // gives the same bits, so voices whose cutoff does not move lose nothing and the warp does not
                // diverge around the binary64 code
                const float fl = __fmul_rn(pow2_ref(__fmul_rn(m, amt_lpf)), lpf);
                make_filt<FILTER>(fc, fl, damp, sr);
            }
            const float ph0 = ph;
            const float osc = osc_step<KIND, false>(kind, o, ph, sintab);
            const float nz = noise_fast(rot, n);
            const float u = __fadd_rn(__fadd_rn(osc, F.gain), __fadd_rn(nz, F.namt));
            const float y = filt_step<FILTER>(u, fc, fs);
            o4[i] = TRACE == TRACE_PHASE ? ph0 : __fmul_rn(y, g);
            n += 1u;
            xf = __fadd_rn(xf, 1.0f);
        }
        *reinterpret_cast<float4*>(row + 4 * j) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
    if (SHARED) fc.fl_bits = __float_as_uint(ctab[96 + kChunk - 1]);      // keep the memo key of the last frame
    F.ph = ph;
    F.x1 = fs.x1; F.x2 = fs.x2; F.y1 = fs.y1; F.y2 = fs.y2;
}

// One frame per lane: the filter coefficients of frames n0 .. n0 + 31 of a cutoff trajectory shared by the warp.
template <int FILTER, class ENV>
__device__ __forceinline__ void modcut_coefficients(const ENV& M, float lpf, float amt_lpf, float damp, float sr,
                                                    uint32_t n0, int lane, float* __restrict__ ctab) {
    const float m = env_x16(M, __uint2float_rn(n0 + (uint32_t)lane));
    const float fl = __fmul_rn(pow2_ref(__fmul_rn(m, amt_lpf)), lpf);
    FiltC c;
    make_filt<FILTER>(c, fl, damp, sr);
    ctab[lane] = c.c0; ctab[32 + lane] = c.c1; ctab[64 + lane] = c.c2; ctab[96 + lane] = fl;
}

// General frame: the normative per-sample semantics, x16 or scalar-tail flavour.
template <int FILTER, int TRACE>
__device__ __forceinline__ float general_frame(const Lane& L, float sr, uint32_t n, bool scalar_sem, OscC& o, FiltC& c,
                               float& ph, FiltS& fs, const float* sintab) {
    const float x = __uint2float_rn(n);           // offset as f32
    float g, m;
    if (!scalar_sem) { g = env_x16(L.amp, x); m = env_x16(L.mod, x); }
    else { g = env_scalar(L.amp, x); m = env_scalar(L.mod, x); }
    const float fo = modulate_freq(L.pitch, m, L.amt_osc);
    const float fl = modulate_freq(L.lpf, m, L.amt_lpf);
    if (__float_as_uint(fo) != o.fo_bits) make_osc(o, fo, sr);
    if (__float_as_uint(fl) != c.fl_bits) make_filt<FILTER>(c, fl, L.damp, sr);
    const float ph0 = ph;
    const float osc = osc_step<-1, true>(L.kind, o, ph, sintab);
    const float nz = noise_literal(L.rot, n);
    float u;
    if (!scalar_sem) u = __fadd_rn(__fadd_rn(osc, L.gain), __fadd_rn(nz, L.namt));
    else u = __fadd_rn(__fmul_rn(osc, L.gain), __fmul_rn(nz, L.namt));   // process.rs:287-294
    const float y = filt_step<FILTER>(u, c, fs);
    return TRACE == TRACE_PHASE ? ph0 : __fmul_rn(y, g);
}

// ------------------------------------------------------------------------------------------

template <int NV, int FILTER, int KIND, int TRACE>
__device__ __forceinline__ void chunk_fast_dispatch(bool gconst, bool namt0, bool aligned8, FastV<NV>& F, const EnvP* amp0,
                                                    const uint32_t (&kind)[NV],
                                                    const uint32_t (&rot)[NV], const uint32_t (&n)[NV],
                                                    float* tile, int lane, const float* sintab) {
    if constexpr (NV == 1) {
        float* row = tile + lane * kTileStride;
        // the sustain / tail steady state gets the fully specialised loop; envelope ramps, added noise
        // amounts and odd offsets are a small share of a render and share more general variants
        if (gconst && namt0 && aligned8) chunk_fast_tp<FILTER, KIND, true, true, true, TRACE>(F, amp0, kind[0], rot[0], n[0], row, sintab);
        else if (gconst) chunk_fast_tp<FILTER, KIND, true, false, false, TRACE>(F, amp0, kind[0], rot[0], n[0], row, sintab);
        else chunk_fast_tp<FILTER, KIND, false, false, false, TRACE>(F, amp0, kind[0], rot[0], n[0], row, sintab);
    } else {
        if (gconst) {
            if (namt0) chunk_fast<NV, FILTER, KIND, true, true, TRACE>(F, kind, rot, n, tile, lane, sintab);
            else chunk_fast<NV, FILTER, KIND, true, false, TRACE>(F, kind, rot, n, tile, lane, sintab);
        } else {
            chunk_fast<NV, FILTER, KIND, false, false, TRACE>(F, kind, rot, n, tile, lane, sintab);
        }
    }
}

// Per-voice state that only the classifier, the general path and the epilogue touch.  It lives in
// shared memory ("coefficient and state tiles"), not in registers: the fast loop then owns the whole
// 128-register budget that keeps all 13.8 warps per SM resident.  An odd word count keeps the 32
// lanes of a warp on distinct banks.
struct Cold {
    Lane L;
    OscC oc;
    FiltC fc;
    FastEnv fe;
    uint32_t n_safe;       // fast constants are valid for frame offsets [.., n_safe)
    uint32_t n_gc;         // the amp envelope is a constant (sustain / end) for offsets [.., n_gc); 0 = ramping
    uint32_t vi;           // slot index (state/params column)
    uint32_t out_row;      // caller-visible voice index, 0xffffffff = no such voice
    uint32_t flags;        // bit 0 active, bit 1 mod envelope matters
};
constexpr int kColdWords = (sizeof(Cold) / 4) | 1;
constexpr int kRowPtrWords = 2;    // one 64-bit output-row base per tile row (0 = the row has no output)
constexpr int kCoefWords = 128;    // shared moving-cutoff coefficients of one chunk: [c0 | c1 | c2 | fl][32]

template <int NV>
__host__ __device__ constexpr size_t warp_smem_floats() {
    return 32 * NV * kTileStride + 32 * NV * kColdWords + 32 * NV * kRowPtrWords + kCoefWords;
}

}  // namespace s2
