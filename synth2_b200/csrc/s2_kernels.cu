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

// 2^x.  The reference calls sleef pow (x16, process.rs:244) / libm powf (scalar, process.rs:227);
// neither is reproducible bit-for-bit on a GPU.  Evaluating in binary64 and rounding once gives the
// correctly rounded binary32 result in all but ~1e-8 of cases, which is what glibc returns too.
__device__ __forceinline__ float pow2_ref(float x) { return (float)exp2((double)x); }
__device__ __forceinline__ float exp_ref(float x) { return (float)exp((double)x); }

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
        const float s = (float)sin((double)th);
        const float co = (float)cos((double)th);
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
// Fast chunk: period, cutoff and envelope segment are constant over the 32 frames of every lane.

struct FastEnv { float es, ex0, ey0; };   // g = es * (x - ex0) + ey0 reproduces each stage bit-exactly

template <int FILTER, int KIND, bool GCONST, int TRACE>
__device__ __forceinline__ void chunk_fast(const Lane& L, const OscC& o, const FiltC& c, const FastEnv& fe,
                                           float& ph, FiltS& fs, uint32_t n0, float* __restrict__ row,
                                           const float* sintab) {
    uint32_t n = n0;
    float xf = __uint2float_rn(n0);               // exact: the caller guarantees n0 + 32 <= 2^24
    // 8 frames per trip: long enough for the scheduler to overlap the two recurrences (phase, filter)
    // of neighbouring frames, short enough (~3 KB of SASS) to live in the instruction cache.
#pragma unroll 2
    for (int j = 0; j < kChunk / 4; j++) {
        float o4[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float ph0 = ph;
            const float osc = osc_step<KIND, false>(L.kind, o, ph, sintab);
            const float nz = noise_fast(L.rot, n);
            // process.rs:341-358: gain and noise amount are ADDED on the x16 path
            const float u = __fadd_rn(__fadd_rn(osc, L.gain), __fadd_rn(nz, L.namt));
            const float y = filt_step<FILTER>(u, c, fs);
            float g;
            if (GCONST) g = fe.ey0;
            else g = __fadd_rn(__fmul_rn(fe.es, __fsub_rn(xf, fe.ex0)), fe.ey0);
            o4[i] = TRACE == TRACE_PHASE ? ph0 : __fmul_rn(y, g);   // process.rs:373-378
            n += 1u;
            xf = __fadd_rn(xf, 1.0f);
        }
        *reinterpret_cast<float4*>(row + 4 * j) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
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

template <int FILTER, int KIND, int TRACE>
__device__ __forceinline__ void chunk_fast_dispatch(bool gconst, const Lane& L, const OscC& o, const FiltC& c,
                                                    const FastEnv& fe, float& ph, FiltS& fs, uint32_t n0,
                                                    float* row, const float* sintab) {
    if (gconst) chunk_fast<FILTER, KIND, true, TRACE>(L, o, c, fe, ph, fs, n0, row, sintab);
    else chunk_fast<FILTER, KIND, false, TRACE>(L, o, c, fe, ph, fs, n0, row, sintab);
}

template <int FILTER, int TRACE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
render_seq_kernel(const RenderArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    float* tile = smem + warp * (32 * kTileStride);
    float* sintab = smem + kWarpsPerBlock * (32 * kTileStride);
    if (a.has_sine) {
        for (int i = threadIdx.x; i < 1024; i += kWarpsPerBlock * 32) sintab[i] = __uint_as_float(d_sin_bits[i]);
        __syncthreads();
    }

    const uint32_t gwarp = blockIdx.x * kWarpsPerBlock + warp;
    const uint32_t vbase = gwarp * 32u;
    if (vbase >= a.n_voices) return;
    const uint32_t v = vbase + lane;
    const bool exists = v < a.n_voices;
    const uint32_t vi = exists ? v : vbase;        // out-of-range lanes shadow lane 0's loads, never store
    const float* __restrict__ P = a.params + vi;
    const uint32_t vp = a.vpad;
    const float sr = a.sample_rate;

    const bool active = exists && __float_as_uint(P[P_ACTIVE * vp]) != 0u;

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

    float* __restrict__ S = a.state + vi;
    float ph = __float_as_uint(S[S_HAS_PHASE * vp]) != 0u ? S[S_PHASE * vp] : 0.0f;   // process.rs:316
    uint32_t n = __float_as_uint(S[S_OFFSET * vp]);
    FiltS fs;
    if (FILTER == 0) { fs.y1 = S[S_LAST * vp]; fs.x1 = fs.x2 = fs.y2 = 0.0f; }
    else { fs.x1 = S[S_X1 * vp]; fs.x2 = S[S_X2 * vp]; fs.y1 = S[S_Y1 * vp]; fs.y2 = S[S_Y2 * vp]; }

    OscC oc; oc.fo_bits = 0x7fc00001u;   // impossible frequency bits -> first use derives the constants
    oc.P = oc.d = oc.slope = oc.half = oc.ts1 = oc.ts2 = 0.0f;
    FiltC fc; fc.fl_bits = 0x7fc00001u; fc.c0 = fc.c1 = fc.c2 = 0.0f;
    FastEnv fe = {0.0f, 0.0f, 0.0f};
    uint32_t n_safe = 0u;                // fast constants are valid for offsets [.., n_safe)
    const bool mod_matters = L.amt_osc != 0.0f || L.amt_lpf != 0.0f;

    // Warp-uniform oscillator kind -> straight-line specialised loop; mixed warps use the per-lane select.
    const uint32_t amask = __ballot_sync(0xffffffffu, active);
    int wkind = -1;
    {
        const int leader = amask ? __ffs(amask) - 1 : 0;
        const uint32_t k0 = __shfl_sync(0xffffffffu, L.kind, leader);
        if (__all_sync(0xffffffffu, !active || L.kind == k0)) wkind = (int)k0;
    }

    const uint32_t frames = a.frames;
    const uint32_t f16 = frames & ~15u;            // x16 region (process.rs:26-37), then the scalar tail
    float* myrow = tile + lane * kTileStride;
    const size_t stride = a.row_stride;
    float* __restrict__ gout = a.voice_out;
    // Output rows are indexed by the caller's voice index, which the bank may have permuted into
    // kind-uniform warps (P_ROW).  Lane (q, c4) writes 16 bytes of rows 4*i + q, i = 0..7.
    const uint32_t my_out_row = exists ? __float_as_uint(P[P_ROW * vp]) : 0xffffffffu;
    const int q = lane >> 3, c4 = (lane & 7) * 4;
    float* rowp[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t r = __shfl_sync(0xffffffffu, my_out_row, 4 * i + q);
        rowp[i] = (gout && r != 0xffffffffu) ? gout + (size_t)r * stride + c4 : nullptr;
    }
    float* __restrict__ gbus = a.bus_partials ? a.bus_partials + (size_t)gwarp * frames : nullptr;

    for (uint32_t t0 = 0; t0 < frames; t0 += kChunk) {
        const uint32_t cnt = min((uint32_t)kChunk, frames - t0);
        bool fast = active && cnt == kChunk && t0 + kChunk <= f16;
        if (fast && n + kChunk > n_safe) {
            // (Re)classify this lane: which envelope segments is frame n in, and until when.
            fast = false;
            const float x0 = __uint2float_rn(n);
            const int sa = env_stage(L.amp, x0);
            const int sm = env_stage(L.mod, x0);
            const bool mconst = !mod_matters || sm == 2 || sm == 4;
            if (mconst && n < (1u << 24)) {
                const float ba = sa == 0 ? L.amp.A : sa == 1 ? L.amp.AD : sa == 2 ? L.amp.Rs : sa == 3 ? L.amp.E : 4.0e9f;
                const float bm = !mod_matters ? 4.0e9f : (sm == 2 ? L.mod.Rs : 4.0e9f);
                // first integer offset whose f32 image reaches the boundary (exact below 2^24)
                uint32_t lim = min(__float2uint_ru(ba), __float2uint_ru(bm));
                n_safe = min(lim, 1u << 24);
                fe.es = sa == 0 ? L.amp.sA : sa == 1 ? L.amp.sD : sa == 3 ? L.amp.sR : 0.0f;
                fe.ex0 = sa == 1 ? L.amp.A : sa == 3 ? L.amp.Rs : 0.0f;
                fe.ey0 = sa == 1 ? 1.0f : (sa == 2 || sa == 3) ? L.amp.S : 0.0f;
                const float m = (mod_matters && sm == 2) ? L.mod.S : 0.0f;
                const float fo = modulate_freq(L.pitch, m, L.amt_osc);
                const float fl = modulate_freq(L.lpf, m, L.amt_lpf);
                if (__float_as_uint(fo) != oc.fo_bits) make_osc(oc, fo, sr);
                if (__float_as_uint(fl) != fc.fl_bits) make_filt<FILTER>(fc, fl, L.damp, sr);
                // the fast phase step needs 1/P < 1 (and a sane period)
                fast = n + kChunk <= n_safe && oc.d < 1.0f && oc.P > 1.0f;
                if (!(oc.d < 1.0f && oc.P > 1.0f)) n_safe = 0u;
            } else {
                n_safe = 0u;
            }
        }
        const bool lane_ok = fast || !active;
        const bool warp_fast = cnt == kChunk && t0 + kChunk <= f16 && __all_sync(0xffffffffu, lane_ok) && amask != 0u;

        if (warp_fast) {
            // inactive lanes run the same code on zeroed constants; their rows are cleared below
            const bool gconst = __all_sync(0xffffffffu, fe.es == 0.0f);
            switch (wkind) {
            case 0: chunk_fast_dispatch<FILTER, 0, TRACE>(gconst, L, oc, fc, fe, ph, fs, n, myrow, sintab); break;
            case 1: chunk_fast_dispatch<FILTER, 1, TRACE>(gconst, L, oc, fc, fe, ph, fs, n, myrow, sintab); break;
            case 2: chunk_fast_dispatch<FILTER, 2, TRACE>(gconst, L, oc, fc, fe, ph, fs, n, myrow, sintab); break;
            case 3: chunk_fast_dispatch<FILTER, 3, TRACE>(gconst, L, oc, fc, fe, ph, fs, n, myrow, sintab); break;
            default: chunk_fast_dispatch<FILTER, -1, TRACE>(gconst, L, oc, fc, fe, ph, fs, n, myrow, sintab); break;
            }
            n += kChunk;
        } else if (active) {
            for (uint32_t i = 0; i < cnt; i++) {
                const bool scalar_sem = t0 + i >= f16;
                myrow[i] = general_frame<FILTER, TRACE>(L, sr, n, scalar_sem, oc, fc, ph, fs, sintab);
                n += 1u;
            }
        }
        if (!active) {
#pragma unroll
            for (int j = 0; j < kChunk / 4; j++)
                *reinterpret_cast<float4*>(myrow + 4 * j) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        __syncwarp();

        if (gout) {
            if (cnt == kChunk) {
                // transposed write-back: lanes 8q..8q+7 cover 128 contiguous bytes of row 4*i + q
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (rowp[i]) {
                        const float4 val = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                        __stcs(reinterpret_cast<float4*>(rowp[i] + t0), val);
                    }
                }
            } else {
                for (uint32_t r = 0; r < 32u && vbase + r < a.n_voices; r++) {
                    const uint32_t orow = __shfl_sync(0xffffffffu, my_out_row, r);
                    if ((uint32_t)lane < cnt)
                        gout[(size_t)orow * stride + t0 + lane] = tile[r * kTileStride + lane];
                }
            }
        }
        if (gbus) {
            // synth.rs:176-202: voices are accumulated in index order, starting from 0.0
            if ((uint32_t)lane < cnt) {
                float acc = 0.0f;
#pragma unroll 8
                for (int r = 0; r < 32; r++) acc = __fadd_rn(acc, tile[r * kTileStride + lane]);
                gbus[t0 + lane] = acc;
            }
        }
        __syncwarp();
    }

    if (active) {
        S[S_PHASE * vp] = ph;
        S[S_HAS_PHASE * vp] = __uint_as_float(1u);
        const uint32_t start = __float_as_uint(S[S_OFFSET * vp]);
        const uint32_t nxt = start + frames < start ? 0xffffffffu : start + frames;   // saturating (synth.rs:197)
        S[S_OFFSET * vp] = __uint_as_float(nxt);
        if (FILTER == 0) S[S_LAST * vp] = fs.y1;
        else { S[S_X1 * vp] = fs.x1; S[S_X2 * vp] = fs.x2; S[S_Y1 * vp] = fs.y1; S[S_Y2 * vp] = fs.y2; }
    }
}

// bus[t] = ((0 + p_0[t]) + p_1[t]) + ... over warps in index order; coalesced over t.
__global__ void bus_reduce_kernel(const float* __restrict__ partials, uint32_t n_warps, uint32_t frames,
                                  float* __restrict__ bus) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= frames) return;
    float acc = n_warps ? partials[t] : 0.0f;     // 0.0 + p_0 == p_0 (p_0 is never -0.0: it starts from +0.0)
    for (uint32_t w = 1; w < n_warps; w++) acc = __fadd_rn(acc, partials[(size_t)w * frames + t]);
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

template <int FILTER, int TRACE>
static cudaError_t launch_t(const RenderArgs& a, cudaStream_t stream) {
    const uint32_t n_warps = (a.n_voices + 31u) / 32u;
    const uint32_t blocks = (n_warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const size_t smem = (size_t)kWarpsPerBlock * 32 * kTileStride * sizeof(float) + (a.has_sine ? 4096 : 0);
    render_seq_kernel<FILTER, TRACE><<<blocks, kWarpsPerBlock * 32, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_render(const RenderArgs& a, uint32_t filter_kind, int trace, cudaStream_t stream) {
    if (a.n_voices == 0 || a.frames == 0) return cudaSuccess;
    if (filter_kind == 0) {
        return trace == TRACE_PHASE ? launch_t<0, TRACE_PHASE>(a, stream) : launch_t<0, TRACE_NONE>(a, stream);
    }
    return trace == TRACE_PHASE ? launch_t<1, TRACE_PHASE>(a, stream) : launch_t<1, TRACE_NONE>(a, stream);
}

cudaError_t launch_bus_reduce(const float* partials, uint32_t n_warps, uint32_t frames, float* bus,
                              cudaStream_t stream) {
    if (frames == 0) return cudaSuccess;
    bus_reduce_kernel<<<(frames + 255) / 256, 256, 0, stream>>>(partials, n_warps, frames, bus);
    return cudaGetLastError();
}

}  // namespace s2
