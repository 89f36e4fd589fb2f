// s2_kernels.cu — sm_100a render kernels for synth2's oscillator -> filter -> envelope -> mix path.
//
// Mapping: one lane = one voice, one warp = 32 consecutive voices, time runs
// sequentially inside the lane (the f32 phase recurrence of oscillators.rs:377-381 is not
// associative, so its rounding sequence is replayed exactly).  A warp produces a 32 voices x 32
// frames tile in shared memory and writes it back transposed, so every STG.128 of the warp
// covers four full 128-byte lines of voice-major output.
//
// Arithmetic contract: every operation below that feeds a discrete decision (phase, table index,
// envelope stage, noise hash) is the reference's binary32 operation, spelled with __f*_rn
// intrinsics so nothing is contracted or reassociated (the file is also built with -fmad=false).
// Reference line numbers are relative to /root/reference/components/s2_lib/src/.
#include "s2_internal.h"
#include "s2_math.h"

// 1: biquad products are fused into the running sum (FFMA); 0: every product and sum rounded
// separately, as the source reads.  A resonant low-cutoff biquad in f32 direct form amplifies
// per-frame rounding differences by ~1/(1-r) (~10^3 at 100 Hz, damping 0.2): the fused form drifted
// 1.8e-4 from the oracle in tests, over the 1e-4 bar, so the unfused form is the default.
#ifndef S2_FUSED_BIQUAD
#define S2_FUSED_BIQUAD 0
#endif
// 1: the phase recurrence of a 32-frame chunk runs as its own pass ahead of everything else.
#ifndef S2_SPLIT_PHASE
#define S2_SPLIT_PHASE 0
#endif

namespace s2 {

__device__ const uint32_t d_sin_bits[1024] = {
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
__device__ __forceinline__ int env_stage(const EnvP& e, float x) {
    return x < e.A ? 0 : (x < e.AD ? 1 : (x < e.Rs ? 2 : (x < e.E ? 3 : 4)));
}

// old/simdtest.rs:270-330 for one lane; line = (rise/run)*x + y0, never fused (:247-261)
__device__ __forceinline__ float env_x16(const EnvP& e, float x) {
    switch (env_stage(e, x)) {
    case 0: return __fadd_rn(__fmul_rn(e.sA, x), 0.0f);
    case 1: return __fadd_rn(__fmul_rn(e.sD, __fsub_rn(x, e.A)), 1.0f);
    case 2: return e.S;
    case 3: return __fadd_rn(__fmul_rn(e.sR, __fsub_rn(x, e.Rs)), e.S);
    default: return 0.0f;
    }
}

// math.rs:11-19 with feature "fma"
__device__ __forceinline__ float line_fma(float rise, float run, float x, float y0) {
    return __fmaf_rn(__fdiv_rn(rise, run), x, y0);
}

// try3/envelopes.rs:22-149 (tail frames only)
__device__ float env_scalar(const EnvP& e, float x) {
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
    } else {
        // try3/dsp_filters.rs:99-109
        float th = __fmul_rn(2.0f, pi);
        th = __fmul_rn(th, fl);
        th = __fdiv_rn(th, sr);
        float s, co;
        s2_sincosf(th, &s, &co);
        const float hd = __fdiv_rn(damp, 2.0f);
        const float num = __fsub_rn(1.0f, __fmul_rn(hd, s));
        const float den = __fadd_rn(1.0f, __fmul_rn(hd, s));
        const float beta = __fmul_rn(0.5f, __fdiv_rn(num, den));
        const float gamma = __fmul_rn(__fadd_rn(0.5f, beta), co);
        const float alpha = __fdiv_rn(__fsub_rn(__fadd_rn(0.5f, beta), gamma), 4.0f);
        // y = 2*(alpha*s + gamma*y1 - beta*y2): scaling by 2 commutes with round-to-nearest, so the
        // doubling is folded into the coefficients (exact unless an intermediate is subnormal).
        c.c0 = __fmul_rn(2.0f, alpha);
        c.c1 = __fmul_rn(2.0f, beta);
        c.c2 = __fmul_rn(2.0f, gamma);
    }
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
    else ph = t >= 1.0f ? __fadd_rn(t, -1.0f) : t;
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
    } else {
        // try3/dsp_filters.rs:116-128: 2*(alpha*(x + 2*x1 + x2) + gamma*y1 - beta*y2);
        // x + 2*x1 is one fma because 2*x1 is exact
        float sx = __fmaf_rn(2.0f, s.x1, u);
        sx = __fadd_rn(sx, s.x2);
        float t = __fmul_rn(c.c0, sx);
        t = __fadd_rn(t, __fmul_rn(c.c2, s.y1));
        t = __fsub_rn(t, __fmul_rn(c.c1, s.y2));
        s.x2 = s.x1; s.x1 = u; s.y2 = s.y1; s.y1 = t;
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
#if S2_SPLIT_PHASE
    // Pass A — the phase recurrence alone (try3/oscillators.rs:377-381): t = phase + 1/P; `% 1.0` is
    // "subtract 1 when t >= 1" (t < 2, exact).  It is the only chain every other operation of a frame
    // hangs from; running it first (32 frames, parked in the voice's own tile row) leaves pass B
    // feed-forward except for the filter state, which the scheduler can software-pipeline.
#pragma unroll 2
    for (int j = 0; j < kChunk / 4; j++) {
        float p4[NV][4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int e = 0; e < NV; e++) p4[e][i] = vget(F.ph, e);
            const vf<NV> t = vadd(F.ph, F.d);
            vf<NV> w;
#pragma unroll
            for (int e = 0; e < NV; e++) vset(w, e, vget(t, e) >= 1.0f ? 1.0f : 0.0f);
            F.ph = vfma(w, none, t);                               // t - w, exact product
        }
#pragma unroll
        for (int e = 0; e < NV; e++)
            *reinterpret_cast<float4*>(tile + (e * 32 + lane) * kTileStride + 4 * j) =
                make_float4(p4[e][0], p4[e][1], p4[e][2], p4[e][3]);
    }
#endif
    // 8 frames per trip: long enough for the scheduler to overlap neighbouring frames, short enough to
    // live in the instruction cache.
#pragma unroll 2
    for (int j = 0; j < kChunk / 4; j++) {
        float o4[NV][4];
#if S2_SPLIT_PHASE
        float4 pin[NV];
#pragma unroll
        for (int e = 0; e < NV; e++)
            pin[e] = *reinterpret_cast<const float4*>(tile + (e * 32 + lane) * kTileStride + 4 * j);
#endif
#pragma unroll
        for (int i = 0; i < 4; i++) {
#if S2_SPLIT_PHASE
            vf<NV> ph0;
#pragma unroll
            for (int e = 0; e < NV; e++)
                vset(ph0, e, i == 0 ? pin[e].x : i == 1 ? pin[e].y : i == 2 ? pin[e].z : pin[e].w);
#else
            const vf<NV> ph0 = F.ph;
#endif
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
#if !S2_SPLIT_PHASE
            // ---- phase step: t = phase + 1/P; `% 1.0` == subtract 1 when t >= 1 (t < 2, exact)
            const vf<NV> t = vadd(ph0, F.d);
            vf<NV> w;
#pragma unroll
            for (int e = 0; e < NV; e++) vset(w, e, vget(t, e) >= 1.0f ? 1.0f : 0.0f);
            F.ph = vfma(w, none, t);                               // t - w, exact product
#endif
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
#if S2_FUSED_BIQUAD
                // products fused into the running sum (2 roundings fewer per frame, tolerance-level
                // difference from the unfused source; only y1 sits on the frame-to-frame critical path)
                const vf<NV> r = vfma(F.c1, F.y2, vmul(F.c0, sx));          // c1 holds -2*beta
                y = vfma(F.c2, F.y1, r);
#else
                vf<NV> tt = vmul(F.c0, sx);
                tt = vadd(tt, vmul(F.c2, F.y1));
                y = vadd(tt, vmul(F.c1, F.y2));                             // c1 holds -2*beta: exact negation
#endif
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
// ALIGNED8: the voice's frame offset is a multiple of 8 at every 8-frame trip, so offset + i == offset ^ i
// and the noise hash input of frame i is one LOP3 with an immediate.
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
    uint32_t n = n0;
    float xf = __uint2float_rn(n0);                               // exact: n0 + 32 <= 2^24
    EnvP A;
    if (!GCONST) A = *amp;
    FiltS fs = {F.x1, F.x2, F.y1, F.y2};
    FiltC fc;
    fc.c0 = F.c0; fc.c1 = FILTER == 0 ? F.c1 : -F.c1; fc.c2 = F.c2; fc.fl_bits = 0;   // F.c1 holds -2*beta for the biquad
    float ph = F.ph;
    // 8 frames per trip: long enough to overlap neighbouring frames, short enough for the instruction cache
#pragma unroll 1
    for (int jt = 0; jt < kChunk / 8; jt++) {
    const uint32_t nb = rot ^ n;                                  // hash input base of this trip
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
        const int j = 2 * jt + jj;
        float o4[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // ---- phase recurrence, two frames (try3/oscillators.rs:377-381; see osc_step)
            const float pa = ph;
            const float ta = __fadd_rn(pa, F.d);
            const float pb = ta >= 1.0f ? __fadd_rn(ta, -1.0f) : ta;
            const float tb = __fadd_rn(pb, F.d);
            ph = tb >= 1.0f ? __fadd_rn(tb, -1.0f) : tb;
            const float2 ph2 = make_float2(pa, pb);
            // ---- waveform: x = period.mul_add(phase, 0); `% period` is a no-op
            const float2 x2 = pmul2(P2, ph2);
            float2 osc2;
            if (KIND == 1) {
                osc2 = pfma2(slope2, x2, one2);
            } else if (KIND == 0) {
                // scalar adds: x2 is a packed product (contraction hazard above)
                const float2 dl = make_float2(__fadd_rn(x2.x, F.nhalf), __fadd_rn(x2.y, F.nhalf));
                osc2.x = __uint_as_float((__float_as_uint(dl.x) & 0x80000000u) ^ 0xbf800000u);
                osc2.y = __uint_as_float((__float_as_uint(dl.y) & 0x80000000u) ^ 0xbf800000u);
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
            const uint32_t ha = (ALIGNED8 ? (nb ^ fi) : (rot ^ (n + fi))) * 0x9e3779b9u;
            const uint32_t hb = (ALIGNED8 ? (nb ^ (fi + 1u)) : (rot ^ (n + fi + 1u))) * 0x9e3779b9u;
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
                g2.x = env_x16(A, xf);
                g2.y = env_x16(A, __fadd_rn(xf, 1.0f));
                xf = __fadd_rn(xf, 2.0f);
            }
            const float2 out2 = TRACE == TRACE_PHASE ? ph2 : pmul2(make_float2(ya, yb), g2);
            o4[2 * h] = out2.x;
            o4[2 * h + 1] = out2.y;
        }
        *reinterpret_cast<float4*>(row + 4 * j) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
    n += 8u;
    }
    F.ph = ph;
    F.x1 = fs.x1; F.x2 = fs.x2; F.y1 = fs.y1; F.y2 = fs.y2;
}

// Modulated-cutoff chunk (one voice per lane): the period is constant but the mod envelope is moving,
// so the cutoff — and with it the filter coefficients — changes every frame (process.rs:148-152,
// 363-371; the first 200 ms of every note of the default patch, synth.rs:141-150).  Everything else
// keeps its fast form; envelopes are evaluated per frame with their full stage chain.
template <int FILTER, int KIND, int TRACE>
__device__ __forceinline__ void chunk_modcut(FastV<1>& F, const EnvP* __restrict__ amp, const EnvP* __restrict__ mod,
                                             float lpf, float amt_lpf, float damp, float sr, FiltC& fc, uint32_t kind,
                                             uint32_t rot, uint32_t n0, float* __restrict__ row, const float* sintab) {
    const EnvP A = *amp, M = *mod;
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
            const float g = env_x16(A, xf);
            const float m = env_x16(M, xf);
            const float fl = modulate_freq(lpf, m, amt_lpf);
            if (__float_as_uint(fl) != fc.fl_bits) make_filt<FILTER>(fc, fl, damp, sr);
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
    F.ph = ph;
    F.x1 = fs.x1; F.x2 = fs.x2; F.y1 = fs.y1; F.y2 = fs.y2;
}

// General frame: the normative per-sample semantics, x16 or scalar-tail flavour.
template <int FILTER, int TRACE>
__device__ float general_frame(const Lane& L, float sr, uint32_t n, bool scalar_sem, OscC& o, FiltC& c,
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
constexpr int kRowPtrWords = 18;   // 8 row pointers (16 words) + 2 pad: conflict-free LDS.64

template <int NV>
__host__ __device__ constexpr size_t warp_smem_floats() {
    return 32 * NV * kTileStride + 32 * NV * kColdWords + 32 * NV * kRowPtrWords;
}

// One warp renders 32*NV consecutive slots; lane l owns slots base + l (+ 32 for its second voice).
// 65,536 voices = 13.8 one-warp blocks per SM: all of them must be resident at once (a second wave would
// serialise), and at 144 registers only 13 fit (measured: +20 % time), so NV = 1 is held to 128.
template <int NV, int FILTER, int TRACE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) __maxnreg__(NV == 1 ? 128 : 255)
render_kernel(const RenderArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int kRows = 32 * NV;
    constexpr int kTileFloats = kRows * kTileStride;
    float* wsm = smem + warp * warp_smem_floats<NV>();
    float* tile = wsm;
    float* cold_base = wsm + kTileFloats;
    float* sintab = smem + kWarpsPerBlock * warp_smem_floats<NV>();
    if (a.has_sine) {
        for (int i = threadIdx.x; i < 1024; i += kWarpsPerBlock * 32) sintab[i] = __uint_as_float(d_sin_bits[i]);
        __syncthreads();
    }

    const uint32_t gwarp = blockIdx.x * kWarpsPerBlock + warp;
    const uint32_t vbase = gwarp * kRows;
    if (vbase >= a.n_voices) return;
    const uint32_t vp = a.vpad;
    const float sr = a.sample_rate;
    auto cold = [&](int e) -> Cold& { return *reinterpret_cast<Cold*>(cold_base + (e * 32 + lane) * kColdWords); };

    uint32_t kind[NV], rot[NV], n[NV];
    bool active[NV];
    FastV<NV> F;
#pragma unroll
    for (int e = 0; e < NV; e++) {
        Cold& C = cold(e);
        const uint32_t v = vbase + e * 32 + lane;
        const bool exists = v < a.n_voices;
        const uint32_t vi = exists ? v : vbase;    // out-of-range lanes shadow slot vbase's loads, never store
        const float* __restrict__ P = a.params + vi;
        active[e] = exists && __float_as_uint(P[P_ACTIVE * vp]) != 0u;
        C.vi = vi;
        C.out_row = exists ? __float_as_uint(P[P_ROW * vp]) : 0xffffffffu;
        Lane L;
        L.kind = kind[e] = __float_as_uint(P[P_KIND * vp]);
        const uint32_t seed = __float_as_uint(P[P_SEED * vp]);
        L.rot = rot[e] = (seed << 5) | (seed >> 27);
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
        C.L = L;
        C.flags = (active[e] ? 1u : 0u) | ((L.amt_osc != 0.0f || L.amt_lpf != 0.0f) ? 2u : 0u);

        const float* __restrict__ S = a.state + vi;
        vset(F.ph, e, __float_as_uint(S[S_HAS_PHASE * vp]) != 0u ? S[S_PHASE * vp] : 0.0f);   // process.rs:316
        n[e] = __float_as_uint(S[S_OFFSET * vp]);
        if (FILTER == 0) {
            vset(F.y1, e, S[S_LAST * vp]); vset(F.x1, e, 0.0f); vset(F.x2, e, 0.0f); vset(F.y2, e, 0.0f);
        } else {
            vset(F.x1, e, S[S_X1 * vp]); vset(F.x2, e, S[S_X2 * vp]);
            vset(F.y1, e, S[S_Y1 * vp]); vset(F.y2, e, S[S_Y2 * vp]);
        }
        vset(F.gain, e, L.gain);
        vset(F.namt, e, L.namt);
        // memoised derived constants (keys = exact input bits; kNoKey -> first use derives them)
        OscC oc;
        oc.fo_bits = __float_as_uint(S[S_FO_KEY * vp]);
        oc.P = S[S_OSC_P * vp]; oc.d = S[S_OSC_D * vp]; oc.slope = S[S_OSC_SLOPE * vp];
        oc.half = S[S_OSC_HALF * vp]; oc.ts1 = S[S_OSC_TS1 * vp]; oc.ts2 = S[S_OSC_TS2 * vp];
        C.oc = oc;
        FiltC fc;
        fc.fl_bits = __float_as_uint(S[S_DAMP_KEY * vp]) == __float_as_uint(L.damp) ? __float_as_uint(S[S_FL_KEY * vp]) : kNoKey;
        fc.c0 = S[S_FC_C0 * vp]; fc.c1 = S[S_FC_C1 * vp]; fc.c2 = S[S_FC_C2 * vp];
        C.fc = fc;
        C.fe = {0.0f, 0.0f, 0.0f};
        C.n_safe = 0u;
        C.n_gc = 0u;
        vset(F.P, e, 0.0f); vset(F.d, e, 0.0f); vset(F.slope, e, 0.0f); vset(F.nhalf, e, 0.0f);
        vset(F.ts1, e, 0.0f); vset(F.ts2, e, 0.0f);
        vset(F.c0, e, 0.0f); vset(F.c1, e, 0.0f); vset(F.c2, e, 0.0f);
        vset(F.es, e, 0.0f); vset(F.nex0, e, 0.0f); vset(F.ey0, e, 0.0f);
    }

    // Warp-uniform oscillator kind -> straight-line specialised loop; mixed warps use the per-voice select.
    bool lane_any_active = false, lane_namt0 = true, lane_al8 = true;
#pragma unroll
    for (int e = 0; e < NV; e++) {
        lane_any_active |= active[e];
        lane_namt0 &= !active[e] || __float_as_uint(vget(F.namt, e)) == 0u;      // +0.0 only
        lane_al8 &= !active[e] || (n[e] & 7u) == 0u;
    }
    const uint32_t amask = __ballot_sync(0xffffffffu, lane_any_active);
    int wkind = -1;
    {
        uint32_t mykind = 0xffu;       // first active kind of this lane
#pragma unroll
        for (int e = NV - 1; e >= 0; e--) if (active[e]) mykind = kind[e];
        const int leader = amask ? __ffs(amask) - 1 : 0;
        const uint32_t k0 = __shfl_sync(0xffffffffu, mykind, leader);
        bool same = true;
#pragma unroll
        for (int e = 0; e < NV; e++) same &= !active[e] || kind[e] == k0;
        if (__all_sync(0xffffffffu, same)) wkind = (int)k0;
    }
    const bool namt0 = __all_sync(0xffffffffu, lane_namt0);
    // offsets advance by whole 32-frame chunks while on the fast path, so this holds for the launch
    const bool aligned8 = __all_sync(0xffffffffu, lane_al8);

    const uint32_t frames = a.frames;
    const uint32_t f16 = frames & ~15u;            // x16 region (process.rs:26-37), then the scalar tail
    const size_t stride = a.row_stride;
    float* __restrict__ gout = a.voice_out;
    // Output rows are indexed by the caller's voice index, which the bank may have permuted into
    // kind-uniform warps (P_ROW).  Transposed write-back: lanes 8q..8q+7 cover 128 contiguous bytes of
    // tile row 4*i + q, so each STG.128 of the warp writes four full 128-byte lines.
    const int q = lane >> 3, c4 = (lane & 7) * 4;
    bool lane_rows_ok = gout != nullptr;
#pragma unroll
    for (int i = 0; i < 8 * NV; i++) {
        const int row = 4 * i + q;                 // tile row = e * 32 + source lane
        const uint32_t r = __shfl_sync(0xffffffffu, cold((4 * i) / 32).out_row, row & 31);
        lane_rows_ok &= r != 0xffffffffu;
        reinterpret_cast<unsigned long long*>(cold_base + kRows * kColdWords)[((i / 8) * 32 + lane) * (kRowPtrWords / 2) + (i & 7)] =
            (gout && r != 0xffffffffu) ? reinterpret_cast<unsigned long long>(gout + (size_t)r * stride + c4) : 0ull;
    }
    const bool all_rows = __all_sync(0xffffffffu, lane_rows_ok);   // every tile row has an output row
    float* __restrict__ gbus = a.bus_partials ? a.bus_partials + (size_t)gwarp * frames : nullptr;

    // fast_left: frames for which every voice of the warp stays on the fast path (warp-uniform), with
    // the loop variant (gconst) chosen when it was computed; 0 = classify before the next chunk.
    uint32_t fast_left = 0, gc_left = 0;
    bool nv2_gconst = false;

    for (uint32_t t0 = 0; t0 < frames; t0 += kChunk) {
        const uint32_t cnt = min((uint32_t)kChunk, frames - t0);
        const bool full = cnt == kChunk && t0 + kChunk <= f16;
        bool warp_fast = full && fast_left >= (uint32_t)kChunk;
        bool warp_semi = false;
        if (!warp_fast) {
            bool lane_ok = true, lane_semi_ok = true;
#pragma unroll
            for (int e = 0; e < NV; e++) {
                Cold& C = cold(e);
                bool fast = active[e] && full;
                bool semi = false;
                if (fast && n[e] + kChunk > C.n_safe) {
                    // (Re)classify this voice: which envelope segments is frame n in, and until when.
                    fast = false;
                    const uint32_t ne = n[e];
                    const float x0 = __uint2float_rn(ne);
                    const EnvP A = C.L.amp;
                    const EnvP M = C.L.mod;
                    const bool mm = (C.flags & 2u) != 0u;
                    const int sa = env_stage(A, x0);
                    const int sm = env_stage(M, x0);
                    const bool mconst = !mm || sm == 2 || sm == 4;
                    uint32_t n_safe = 0u;
                    if (NV == 1 && !mconst && C.L.amt_osc == 0.0f && ne + kChunk <= (1u << 24)) {
                        // the mod envelope moves but only the cutoff follows it: the period is the
                        // per-voice constant sr / pitch -> modulated-cutoff chunk
                        OscC oc = C.oc;
                        if (__float_as_uint(C.L.pitch) != oc.fo_bits) { make_osc(oc, C.L.pitch, sr); C.oc = oc; }
                        if (oc.d < 1.0f && oc.P > 1.0f) {
                            semi = true;
                            vset(F.P, e, oc.P); vset(F.d, e, oc.d); vset(F.slope, e, oc.slope);
                            vset(F.nhalf, e, -oc.half); vset(F.ts1, e, oc.ts1); vset(F.ts2, e, oc.ts2);
                        }
                    }
                    if (mconst && ne < (1u << 24)) {
                        const float ba = sa == 0 ? A.A : sa == 1 ? A.AD : sa == 2 ? A.Rs : sa == 3 ? A.E : 4.0e9f;
                        const float bm = !mm ? 4.0e9f : (sm == 2 ? M.Rs : 4.0e9f);
                        // first integer offset whose f32 image reaches the boundary (exact below 2^24).
                        // NV = 1 evaluates the amp envelope per frame when it ramps, so only the mod
                        // envelope bounds the fast constants there.
                        const uint32_t na = __float2uint_ru(ba);
                        n_safe = min(NV == 1 ? __float2uint_ru(bm) : min(na, __float2uint_ru(bm)), 1u << 24);
                        C.n_gc = (sa == 2 || sa == 4) ? min(na, 1u << 24) : 0u;
                        FastEnv fe;
                        fe.es = sa == 0 ? A.sA : sa == 1 ? A.sD : sa == 3 ? A.sR : 0.0f;
                        fe.ex0 = sa == 1 ? A.A : sa == 3 ? A.Rs : 0.0f;
                        fe.ey0 = sa == 1 ? 1.0f : (sa == 2 || sa == 3) ? A.S : 0.0f;
                        C.fe = fe;
                        const float m = (mm && sm == 2) ? M.S : 0.0f;
                        const float fo = modulate_freq(C.L.pitch, m, C.L.amt_osc);
                        const float fl = modulate_freq(C.L.lpf, m, C.L.amt_lpf);
                        OscC oc = C.oc;
                        FiltC fc = C.fc;
                        if (__float_as_uint(fo) != oc.fo_bits) { make_osc(oc, fo, sr); C.oc = oc; }
                        if (__float_as_uint(fl) != fc.fl_bits) { make_filt<FILTER>(fc, fl, C.L.damp, sr); C.fc = fc; }
                        // the fast phase step needs 1/P < 1 (and a sane period)
                        const bool sane = oc.d < 1.0f && oc.P > 1.0f;
                        if (!sane) n_safe = 0u;
                        fast = sane && ne + kChunk <= n_safe;
                        // publish into the lane vectors
                        vset(F.P, e, oc.P); vset(F.d, e, oc.d); vset(F.slope, e, oc.slope);
                        vset(F.nhalf, e, -oc.half); vset(F.ts1, e, oc.ts1); vset(F.ts2, e, oc.ts2);
                        vset(F.c0, e, fc.c0);
                        vset(F.c1, e, FILTER == 0 ? fc.c1 : -fc.c1);
                        vset(F.c2, e, fc.c2);
                        vset(F.es, e, fe.es); vset(F.nex0, e, -fe.ex0); vset(F.ey0, e, fe.ey0);
                    }
                    C.n_safe = n_safe;
                }
                lane_ok &= fast || !active[e];
                lane_semi_ok &= fast || semi || !active[e];
            }
            warp_fast = full && amask != 0u && __all_sync(0xffffffffu, lane_ok);
            warp_semi = NV == 1 && !warp_fast && full && amask != 0u && __all_sync(0xffffffffu, lane_semi_ok);
            if (warp_fast) {
                uint32_t lane_left = 0xffffffffu;
#pragma unroll
                for (int e = 0; e < NV; e++)
                    if (active[e]) lane_left = min(lane_left, cold(e).n_safe - n[e]);
                fast_left = __reduce_min_sync(0xffffffffu, lane_left);
                gc_left = 0;           // recomputed below
                bool lane_gconst = true;
#pragma unroll
                for (int e = 0; e < NV; e++) lane_gconst &= vget(F.es, e) == 0.0f;
                nv2_gconst = __all_sync(0xffffffffu, lane_gconst);
            } else {
                fast_left = 0;
            }
        }
        if (NV == 1 && warp_fast && gc_left < (uint32_t)kChunk) {
            // how long does every voice's amp envelope stay constant from here?
            uint32_t lane_gc = 0xffffffffu;
#pragma unroll
            for (int e = 0; e < NV; e++) {
                if (!active[e]) continue;
                Cold& C = cold(e);
                if (n[e] >= C.n_gc) {      // left the constant segment (or never in one): look again
                    const EnvP A = C.L.amp;
                    const float x0 = __uint2float_rn(n[e]);
                    const int sa = env_stage(A, x0);
                    const float ba = sa == 2 ? A.Rs : 4.0e9f;
                    C.n_gc = (sa == 2 || sa == 4) ? min(__float2uint_ru(ba), 1u << 24) : 0u;
                    vset(F.ey0, e, sa == 2 ? A.S : 0.0f);
                }
                lane_gc = min(lane_gc, C.n_gc > n[e] ? C.n_gc - n[e] : 0u);
            }
            gc_left = __reduce_min_sync(0xffffffffu, lane_gc);
        }
        const bool gconst = NV == 1 ? gc_left >= (uint32_t)kChunk : nv2_gconst;

        if (warp_fast) {
            // inactive voices run the same code on zeroed constants; their rows are cleared below
            fast_left -= kChunk;
            gc_left = gc_left >= (uint32_t)kChunk ? gc_left - kChunk : 0u;
            switch (wkind) {
            case 0: chunk_fast_dispatch<NV, FILTER, 0, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
            case 1: chunk_fast_dispatch<NV, FILTER, 1, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
            case 2: chunk_fast_dispatch<NV, FILTER, 2, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
            case 3: chunk_fast_dispatch<NV, FILTER, 3, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
            default: chunk_fast_dispatch<NV, FILTER, -1, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
            }
#pragma unroll
            for (int e = 0; e < NV; e++) n[e] += kChunk;
        } else if (warp_semi) {
            if constexpr (NV == 1) {
                // voices that are fully constant run the same code: their cutoff simply does not move
                Cold& C = cold(0);
                FiltC fc = C.fc;
                float* row = tile + lane * kTileStride;
                const float lpf = C.L.lpf, amt = C.L.amt_lpf, damp = C.L.damp;
                switch (wkind) {
                case 0: chunk_modcut<FILTER, 0, TRACE>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab); break;
                case 1: chunk_modcut<FILTER, 1, TRACE>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab); break;
                case 2: chunk_modcut<FILTER, 2, TRACE>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab); break;
                default: chunk_modcut<FILTER, -1, TRACE>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab); break;
                }
                if (active[0]) C.fc = fc;
                n[0] += kChunk;
            }
        } else {
#pragma unroll
            for (int e = 0; e < NV; e++) {
                if (!active[e]) continue;
                Cold& C = cold(e);
                const Lane L = C.L;
                OscC oc = C.oc;
                FiltC fc = C.fc;
                float ph = vget(F.ph, e);
                FiltS fs = {vget(F.x1, e), vget(F.x2, e), vget(F.y1, e), vget(F.y2, e)};
                float* row = tile + (e * 32 + lane) * kTileStride;
                for (uint32_t i = 0; i < cnt; i++) {
                    const bool scalar_sem = t0 + i >= f16;
                    row[i] = general_frame<FILTER, TRACE>(L, sr, n[e], scalar_sem, oc, fc, ph, fs, sintab);
                    n[e] += 1u;
                }
                C.oc = oc;
                C.fc = fc;
                vset(F.ph, e, ph);
                vset(F.x1, e, fs.x1); vset(F.x2, e, fs.x2); vset(F.y1, e, fs.y1); vset(F.y2, e, fs.y2);
                C.n_safe = 0u;         // oc/fc may have moved: republish through the classifier
            }
        }
#pragma unroll
        for (int e = 0; e < NV; e++) {
            if (!active[e]) {
                float* row = tile + (e * 32 + lane) * kTileStride;
#pragma unroll
                for (int j = 0; j < kChunk / 4; j++)
                    *reinterpret_cast<float4*>(row + 4 * j) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        __syncwarp();
        if (gout) {
            const size_t tb = (size_t)t0 * sizeof(float);
            const unsigned long long* rp = reinterpret_cast<const unsigned long long*>(cold_base + kRows * kColdWords);
            if (cnt == kChunk && all_rows) {
#pragma unroll
                for (int i = 0; i < 8 * NV; i++) {
                    char* dst = reinterpret_cast<char*>(rp[((i / 8) * 32 + lane) * (kRowPtrWords / 2) + (i & 7)]);
                    const float4 val = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                    __stcs(reinterpret_cast<float4*>(dst + tb), val);
                }
            } else if (cnt == kChunk) {
#pragma unroll
                for (int i = 0; i < 8 * NV; i++) {
                    char* dst = reinterpret_cast<char*>(rp[((i / 8) * 32 + lane) * (kRowPtrWords / 2) + (i & 7)]);
                    if (dst) {
                        const float4 val = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                        __stcs(reinterpret_cast<float4*>(dst + tb), val);
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < NV; e++) {
                    for (uint32_t r = 0; r < 32u; r++) {
                        const uint32_t orow = __shfl_sync(0xffffffffu, cold(e).out_row, r);
                        if (orow != 0xffffffffu && (uint32_t)lane < cnt)
                            gout[(size_t)orow * stride + t0 + lane] = tile[(e * 32 + r) * kTileStride + lane];
                    }
                }
            }
        }
        if (gbus) {
            if ((uint32_t)lane < cnt) {
                float acc;
                if (a.n_voices <= 32u) {
                    // synth.rs:176-202: voices are accumulated in index order, starting from 0.0 — the
                    // reference's exact summation order (a bank this narrow is one warp, identity slots)
                    acc = 0.0f;
#pragma unroll 8
                    for (int r = 0; r < kRows; r++) acc = __fadd_rn(acc, tile[r * kTileStride + lane]);
                } else {
                    // wide banks: fixed 4-way tree over the warp's rows (deterministic; four independent
                    // chains instead of one 32-deep dependent chain)
                    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
                    for (int r = 0; r < kRows; r += 4) {
                        a0 = __fadd_rn(a0, tile[(r + 0) * kTileStride + lane]);
                        a1 = __fadd_rn(a1, tile[(r + 1) * kTileStride + lane]);
                        a2 = __fadd_rn(a2, tile[(r + 2) * kTileStride + lane]);
                        a3 = __fadd_rn(a3, tile[(r + 3) * kTileStride + lane]);
                    }
                    acc = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
                }
                gbus[t0 + lane] = acc;
            }
        }
        __syncwarp();      // every lane is done reading the tile before the next chunk overwrites it
    }

#pragma unroll
    for (int e = 0; e < NV; e++) {
        if (!active[e]) continue;
        const Cold& C = cold(e);
        float* __restrict__ S = a.state + C.vi;
        S[S_PHASE * vp] = vget(F.ph, e);
        S[S_HAS_PHASE * vp] = __uint_as_float(1u);
        const uint32_t start = __float_as_uint(S[S_OFFSET * vp]);
        const uint32_t nxt = start + frames < start ? 0xffffffffu : start + frames;   // saturating (synth.rs:197)
        S[S_OFFSET * vp] = __uint_as_float(nxt);
        if (FILTER == 0) S[S_LAST * vp] = vget(F.y1, e);
        else {
            S[S_X1 * vp] = vget(F.x1, e); S[S_X2 * vp] = vget(F.x2, e);
            S[S_Y1 * vp] = vget(F.y1, e); S[S_Y2 * vp] = vget(F.y2, e);
        }
        const OscC oc = C.oc;
        const FiltC fc = C.fc;
        S[S_FO_KEY * vp] = __uint_as_float(oc.fo_bits);
        S[S_OSC_P * vp] = oc.P; S[S_OSC_D * vp] = oc.d; S[S_OSC_SLOPE * vp] = oc.slope;
        S[S_OSC_HALF * vp] = oc.half; S[S_OSC_TS1 * vp] = oc.ts1; S[S_OSC_TS2 * vp] = oc.ts2;
        S[S_FL_KEY * vp] = __uint_as_float(fc.fl_bits);
        S[S_DAMP_KEY * vp] = C.L.damp;
        S[S_FC_C0 * vp] = fc.c0; S[S_FC_C1 * vp] = fc.c1; S[S_FC_C2 * vp] = fc.c2;
    }
}

// Bus reduction, two stages, fixed summation tree (deterministic):
//   stage 1  block (fx, sy) adds the 64 per-warp partial rows [64*sy, 64*sy + 64) for 32 frames: 8 threads
//            per frame take 8 rows each (independent loads, index order), then the 8 sums are added in
//            order -> seg[sy][t].  4,096 blocks at the bench shape: bandwidth-bound instead of the
//            latency-bound single pass it replaces (51 us -> ~10 us for 32 MB of partials).
//   stage 2  bus[t] = seg[0][t] + seg[1][t] + ... in order (n_seg <= a few dozen, L2-resident).
// Reads are coalesced over t (128 bytes per row per warp).
constexpr uint32_t kBusSegWarps = 64;

__global__ void __launch_bounds__(256) bus_reduce_stage1(const float* __restrict__ partials, uint32_t n_warps,
                                                         uint32_t frames, float* __restrict__ seg) {
    __shared__ float sm[8][33];
    const uint32_t tx = threadIdx.x & 31u, sub = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * 32u + tx;
    const uint32_t w0 = blockIdx.y * kBusSegWarps + sub * 8u;
    float acc = 0.0f;
    if (t < frames) {
        float v[8];
#pragma unroll
        for (uint32_t k = 0; k < 8u; k++) v[k] = (w0 + k < n_warps) ? partials[(size_t)(w0 + k) * frames + t] : 0.0f;
        acc = v[0];
#pragma unroll
        for (uint32_t k = 1; k < 8u; k++) acc = __fadd_rn(acc, v[k]);
    }
    sm[sub][tx] = acc;
    __syncthreads();
    if (sub == 0 && t < frames) {
        float total = sm[0][tx];
#pragma unroll
        for (int k = 1; k < 8; k++) total = __fadd_rn(total, sm[k][tx]);
        seg[(size_t)blockIdx.y * frames + t] = total;
    }
}

__global__ void bus_reduce_stage2(const float* __restrict__ seg, uint32_t n_seg, uint32_t frames,
                                  float* __restrict__ bus) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= frames) return;
    float acc = seg[t];
    for (uint32_t k = 1; k < n_seg; k++) acc = __fadd_rn(acc, seg[(size_t)k * frames + t]);
    bus[t] = acc;
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ staged, const float* __restrict__ row_index_bits,
                                  uint32_t* __restrict__ dst, uint32_t n) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) dst[s] = staged[__float_as_uint(row_index_bits[s])];
}

cudaError_t launch_gather_u32(const uint32_t* staged, const float* row_index_bits, uint32_t* dst_row,
                              uint32_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    gather_u32_kernel<<<(n + 255) / 256, 256, 0, stream>>>(staged, row_index_bits, dst_row, n);
    return cudaGetLastError();
}

uint32_t render_warps(uint32_t n_voices, int nv) { return (n_voices + 32u * nv - 1u) / (32u * nv); }

template <int NV, int FILTER, int TRACE>
static cudaError_t launch_t(const RenderArgs& a, cudaStream_t stream) {
    const uint32_t n_warps = render_warps(a.n_voices, NV);
    const uint32_t blocks = (n_warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const size_t smem = (size_t)kWarpsPerBlock * warp_smem_floats<NV>() * sizeof(float) + (a.has_sine ? 4096 : 0);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(render_kernel<NV, FILTER, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_set = true;
    }
    render_kernel<NV, FILTER, TRACE><<<blocks, kWarpsPerBlock * 32, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int NV>
static cudaError_t launch_nv(const RenderArgs& a, uint32_t filter_kind, int trace, cudaStream_t stream) {
    if (filter_kind == 0)
        return trace == TRACE_PHASE ? launch_t<NV, 0, TRACE_PHASE>(a, stream) : launch_t<NV, 0, TRACE_NONE>(a, stream);
    return trace == TRACE_PHASE ? launch_t<NV, 1, TRACE_PHASE>(a, stream) : launch_t<NV, 1, TRACE_NONE>(a, stream);
}

cudaError_t launch_render(const RenderArgs& a, uint32_t filter_kind, int trace, int nv, cudaStream_t stream) {
    if (a.n_voices == 0 || a.frames == 0) return cudaSuccess;
    return nv == 2 ? launch_nv<2>(a, filter_kind, trace, stream) : launch_nv<1>(a, filter_kind, trace, stream);
}

uint32_t bus_segments(uint32_t n_warps) { return (n_warps + kBusSegWarps - 1) / kBusSegWarps; }

cudaError_t launch_bus_reduce(const float* partials, uint32_t n_warps, uint32_t frames, float* seg_scratch,
                              float* bus, cudaStream_t stream) {
    if (frames == 0) return cudaSuccess;
    const uint32_t n_seg = bus_segments(n_warps);
    bus_reduce_stage1<<<dim3((frames + 31) / 32, n_seg), 256, 0, stream>>>(partials, n_warps, frames, seg_scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    bus_reduce_stage2<<<(frames + 255) / 256, 256, 0, stream>>>(seg_scratch, n_seg, frames, bus);
    return cudaGetLastError();
}

}  // namespace s2
