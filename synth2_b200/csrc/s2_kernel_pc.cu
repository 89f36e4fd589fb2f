// s2_kernel_pc.cu — render kernel with a producer/consumer warp pair per voice group.
//
// Why (profiles/r1_notes.md): with one warp per 32 voices a 65,536-voice bank gives only 3.46 warps per
// SM sub-partition, and the loop is latency-bound there (throughput still grows 15 % from 3 to 4 warps
// per scheduler; issue slots are 69 % busy).  The work of a frame splits along its two recurrences:
//
//   producer warp   phase recurrence -> waveform, noise, gain/noise combine        -> u[frame]
//   consumer warp   u[frame] -> filter recurrence -> envelope, output gain -> tile -> HBM, bus
//
// Both warps own the same 32 voices (lane l = voice l in both), exchange u through a double-buffered
// shared-memory tile with named barriers (FULL/EMPTY per buffer), and the consumer's output overwrites u
// in place, so the tile that is written back is the exchange buffer itself.  Twice the warps per
// scheduler for the same voices, each with one short recurrence.
//
// The consumer keeps every decision: it classifies chunks exactly like s2_kernels.cu, and hands the
// producer "runs" of consecutive fast chunks (constant period).  Chunks that are not fast (modulated
// cutoff or pitch, offsets >= 2^24, the x16/tail seam, ragged ends) are rendered by the consumer alone
// with the same code as the one-warp kernel; the phase travels through shared memory at run boundaries.
#include "s2_device.cuh"

namespace s2 {

namespace {

constexpr int kPcThreads = 64;

// Producer/consumer hand-shake through mbarrier objects in shared memory (one elected lane arrives after
// __syncwarp, the other warp waits on the phase parity).  Hardware named barriers would cap the SM at
// 64 / 6 = 10 resident blocks (measured), and this kernel needs 14.
enum { MB_RUN = 0, MB_FULL0 = 1, MB_FULL1 = 2, MB_EMPTY0 = 3, MB_EMPTY1 = 4, MB_COUNT = 5 };

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// all lanes of the signalling warp call this; their earlier shared-memory writes are ordered before it
__device__ __forceinline__ void mbar_arrive_warp(unsigned long long* bar, int lane) {
    __syncwarp();
    if (lane == 0) {
        unsigned long long st;
        asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
        (void)st;
    }
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

struct PcCtrl {            // written by the consumer before it arrives on BAR_RUN
    uint32_t chunks;       // length of the run, in 32-frame chunks
    uint32_t exit;         // 1: no more runs
    int32_t wkind;         // warp-uniform oscillator kind or -1
    uint32_t namt0, aligned8;
    uint32_t pad[3];
    unsigned long long bars[MB_COUNT + 1];
};

// Per-voice state outside the hot loops, in shared memory (see Cold in s2_device.cuh; this one is trimmed
// to what fits 14 blocks per SM: 65,536 voices need all 2,048 blocks resident at once).
struct ColdPc {
    uint32_t kind, rot;
    float pitch, gain, namt, lpf, damp, amt_osc, amt_lpf;
    EnvQ amp, mod;
    OscC oc;
    FiltC fc;
    uint32_t n_safe;       // fast constants are valid for frame offsets [.., n_safe)
    uint32_t n_gc;         // the amp envelope is constant for offsets [.., n_gc); 0 = ramping
    uint32_t flags;        // bit 0 active, bit 1 mod envelope matters
    // hand-over at run boundaries
    float ph;              // in: phase at the first frame of the run; out: phase after its last frame
    uint32_t n;            // frame offset at the first frame of the run
};

constexpr int kPcColdWords = (sizeof(ColdPc) / 4) | 1;
constexpr size_t kPcSmemFloats = 2 * 32 * kTileStride + 32 * kPcColdWords + sizeof(PcCtrl) / 4;

// ---- producer: 32 frames of u for one voice per lane (the A half of chunk_fast_tp) ---------------------
template <int KIND, bool NAMT0, bool ALIGNED8>
__device__ __forceinline__ void produce_chunk(const OscC& o, float gain, float namt, uint32_t kind, uint32_t rot,
                                              uint32_t& n, float& ph, float* __restrict__ row, const float* sintab) {
    const float2 one2 = make_float2(1.0f, 1.0f), none2 = make_float2(-1.0f, -1.0f), two2 = make_float2(2.0f, 2.0f);
    const float2 P2 = make_float2(o.P, o.P), slope2 = make_float2(o.slope, o.slope);
    const float2 ts1_2 = make_float2(o.ts1, o.ts1), ts2_2 = make_float2(o.ts2, o.ts2);
    const float2 gain2 = make_float2(gain, gain), namt2 = make_float2(namt, namt);
    const float nhalf = -o.half;
#pragma unroll 1
    for (int jt = 0; jt < kChunk / 8; jt++) {
        const uint32_t nb = rot ^ n;
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
            float o4[4];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                // phase recurrence, two frames (try3/oscillators.rs:377-381; see osc_step)
                const float pa = ph;
                const float ta = __fadd_rn(pa, o.d);
                const float pb = wrap_unit(ta);
                const float tb = __fadd_rn(pb, o.d);
                ph = wrap_unit(tb);
                const float2 x2 = pmul2(P2, make_float2(pa, pb));
                float2 osc2;
                if (KIND == 1) {
                    osc2 = pfma2(slope2, x2, one2);
                } else if (KIND == 0) {
                    // scalar adds: x2 is a packed product (contraction hazard, s2_device.cuh)
                    const float2 dl = make_float2(__fadd_rn(x2.x, nhalf), __fadd_rn(x2.y, nhalf));
                    osc2.x = __uint_as_float((__float_as_uint(dl.x) & 0x80000000u) ^ 0xbf800000u);
                    osc2.y = __uint_as_float((__float_as_uint(dl.y) & 0x80000000u) ^ 0xbf800000u);
                } else if (KIND == 2) {
                    const float2 dl = make_float2(__fadd_rn(x2.x, nhalf), __fadd_rn(x2.y, nhalf));
                    const float2 a = pfma2(ts1_2, x2, one2);
                    const float2 b = pfma2(ts2_2, dl, none2);
                    osc2.x = dl.x < 0.0f ? a.x : b.x;
                    osc2.y = dl.y < 0.0f ? a.y : b.y;
                } else {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const uint32_t k = KIND == 3 ? 3u : kind;
                        const float xe = e ? x2.y : x2.x;
                        float y;
                        if (k == 1u) y = __fmaf_rn(o.slope, xe, 1.0f);
                        else if (k == 0u) y = xe < o.half ? 1.0f : -1.0f;
                        else if (k == 2u) {
                            const float a = __fmaf_rn(o.ts1, xe, 1.0f);
                            const float b = __fmaf_rn(o.ts2, __fsub_rn(xe, o.half), -1.0f);
                            y = xe < o.half ? a : b;
                        } else {                                   // try3/lookup.rs:46-85 on SIN_TABLE
                            const float tv = __fdiv_rn(__fmul_rn(xe, 1024.0f), o.P);
                            const uint32_t i1 = __float2uint_rz(tv);
                            const uint32_t i2 = (i1 + 1u) & 1023u;
                            const float s1 = i1 < 1024u ? sintab[i1] : 0.0f;
                            const float s2 = sintab[i2];
                            y = __fmaf_rn(__fsub_rn(s2, s1), __fsub_rn(tv, __uint2float_rn(i1)), s1);
                        }
                        if (e) osc2.y = y; else osc2.x = y;
                    }
                }
                // noise (try3/hashnoise.rs:33-68; see noise_fast)
                const uint32_t fi = 4u * jj + 2u * h;
                const uint32_t ha = (ALIGNED8 ? (nb ^ fi) : (rot ^ (n + fi))) * 0x9e3779b9u;
                const uint32_t hb = (ALIGNED8 ? (nb ^ (fi + 1u)) : (rot ^ (n + fi + 1u))) * 0x9e3779b9u;
                const float2 v2 = make_float2(__uint2float_rn(ha & 0xffffu), __uint2float_rn(hb & 0xffffu));
                const float2 q2 = pfma2(v2, make_float2(0x1.0001p-16f, 0x1.0001p-16f),
                                        pmul2(v2, make_float2(0x1.0001p-48f, 0x1.0001p-48f)));
                const float2 nz2 = pfma2(q2, two2, none2);
                // process.rs:341-358 (ADD, x16 quirk); nz + 0.0 == nz bit-for-bit
                const float2 u2 = padd2(padd2(osc2, gain2), NAMT0 ? nz2 : padd2(nz2, namt2));
                o4[2 * h] = u2.x;
                o4[2 * h + 1] = u2.y;
            }
            *reinterpret_cast<float4*>(row + 8 * jt + 4 * jj) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
        n += 8u;
    }
}

template <bool NAMT0, bool ALIGNED8>
__device__ __forceinline__ void produce_dispatch(int wkind, const OscC& o, float gain, float namt, uint32_t kind,
                                                 uint32_t rot, uint32_t& n, float& ph, float* row, const float* sintab) {
    switch (wkind) {
    case 0: produce_chunk<0, NAMT0, ALIGNED8>(o, gain, namt, kind, rot, n, ph, row, sintab); break;
    case 1: produce_chunk<1, NAMT0, ALIGNED8>(o, gain, namt, kind, rot, n, ph, row, sintab); break;
    case 2: produce_chunk<2, NAMT0, ALIGNED8>(o, gain, namt, kind, rot, n, ph, row, sintab); break;
    case 3: produce_chunk<3, NAMT0, ALIGNED8>(o, gain, namt, kind, rot, n, ph, row, sintab); break;
    default: produce_chunk<-1, NAMT0, ALIGNED8>(o, gain, namt, kind, rot, n, ph, row, sintab); break;
    }
}

// ---- consumer: filter + envelope + gain over 32 frames of u, in place (the B half of chunk_fast_tp) -----
template <int FILTER, bool GCONST>
__device__ __forceinline__ void consume_chunk(const FiltC& fc, FiltS& fs, float gconst_level, const EnvQ* __restrict__ amp,
                                              uint32_t n0, float* __restrict__ row) {
    float xf = __uint2float_rn(n0);                               // exact: n0 + 32 <= 2^24
    EnvQ A;
    if (!GCONST) A = *amp;
    const float2 g0 = make_float2(gconst_level, gconst_level);
#pragma unroll 2
    for (int j = 0; j < kChunk / 4; j++) {
        const float4 u4 = *reinterpret_cast<const float4*>(row + 4 * j);
        const float y0 = filt_step<FILTER>(u4.x, fc, fs);
        const float y1 = filt_step<FILTER>(u4.y, fc, fs);
        const float y2 = filt_step<FILTER>(u4.z, fc, fs);
        const float y3 = filt_step<FILTER>(u4.w, fc, fs);
        float2 ga = g0, gb = g0;
        if (!GCONST) {
            ga.x = env_x16(A, xf);
            ga.y = env_x16(A, __fadd_rn(xf, 1.0f));
            gb.x = env_x16(A, __fadd_rn(xf, 2.0f));
            gb.y = env_x16(A, __fadd_rn(xf, 3.0f));
            xf = __fadd_rn(xf, 4.0f);
        }
        const float2 oa = pmul2(make_float2(y0, y1), ga);          // process.rs:373-378
        const float2 ob = pmul2(make_float2(y2, y3), gb);
        *reinterpret_cast<float4*>(row + 4 * j) = make_float4(oa.x, oa.y, ob.x, ob.y);
    }
}

// ---- cold paths of the consumer, kept out of line so that their registers do not count against the
// ---- 72-register budget of the pipelined loop (14 blocks x 64 threads must be resident per SM) --------

struct ConsState { uint32_t n; float ph; FiltS fs; FiltC fcr; float glevel; };

// (Re)classify one voice at frame offset st.n: which envelope segments it is in and until when.
// Returns bit 0: fast (constant period and cutoff), bit 1: semi (constant period, moving cutoff).
struct ClassifyOut { uint32_t res; FiltC fcr; float glevel; };

template <int FILTER>
__device__ __noinline__ ClassifyOut pc_classify(ColdPc& C, uint32_t n, float sr, FiltC fcr_in, float glevel_in) {
    ClassifyOut out;
    out.fcr = fcr_in;
    out.glevel = glevel_in;
    const float x0 = __uint2float_rn(n);
    const EnvQ A = C.amp;
    const EnvQ M = C.mod;
    const bool mm = (C.flags & 2u) != 0u;
    const int sa = env_stage(A, x0);
    const int sm = env_stage(M, x0);
    const bool mconst = !mm || sm == 2 || sm == 4;
    uint32_t n_safe = 0u, res = 0u;
    if (!mconst && C.amt_osc == 0.0f && n + kChunk <= (1u << 24)) {
        // the mod envelope moves but only the cutoff follows it: the period is sr / pitch
        OscC oc = C.oc;
        if (__float_as_uint(C.pitch) != oc.fo_bits) { make_osc(oc, C.pitch, sr); C.oc = oc; }
        if (oc.d < 1.0f && oc.P > 1.0f) res |= 2u;
    }
    if (mconst && n < (1u << 24)) {
        const float ba = sa == 2 ? A.Rs : 4.0e9f;
        const float bm = !mm ? 4.0e9f : (sm == 2 ? M.Rs : 4.0e9f);
        // first integer offset whose f32 image reaches the boundary (exact below 2^24)
        n_safe = min(__float2uint_ru(bm), 1u << 24);
        C.n_gc = (sa == 2 || sa == 4) ? min(__float2uint_ru(ba), 1u << 24) : 0u;
        out.glevel = sa == 2 ? A.S : 0.0f;
        const float m = (mm && sm == 2) ? M.S : 0.0f;
        const float fo = modulate_freq(C.pitch, m, C.amt_osc);
        const float fl = modulate_freq(C.lpf, m, C.amt_lpf);
        OscC oc = C.oc;
        FiltC fc = C.fc;
        if (__float_as_uint(fo) != oc.fo_bits) { make_osc(oc, fo, sr); C.oc = oc; }
        if (__float_as_uint(fl) != fc.fl_bits) { make_filt<FILTER>(fc, fl, C.damp, sr); C.fc = fc; }
        out.fcr = fc;
        const bool sane = oc.d < 1.0f && oc.P > 1.0f;     // the fast phase step needs 1/P < 1
        if (!sane) n_safe = 0u;
        if (sane && n + kChunk <= n_safe) res |= 1u;
    }
    C.n_safe = n_safe;
    out.res = res;
    return out;
}

// amp envelope constant from st.n on?  Refreshes C.n_gc / st.glevel and returns the run length.
// returns the run length; *glevel is updated when the voice (re)enters a constant segment
__device__ __noinline__ uint2 pc_refresh_gc(ColdPc& C, uint32_t n, float glevel) {
    if (n >= C.n_gc) {
        const EnvQ A = C.amp;
        const int sa = env_stage(A, __uint2float_rn(n));
        C.n_gc = (sa == 2 || sa == 4) ? min(__float2uint_ru(sa == 2 ? A.Rs : 4.0e9f), 1u << 24) : 0u;
        glevel = sa == 2 ? A.S : 0.0f;
    }
    return make_uint2(C.n_gc > n ? C.n_gc - n : 0u, __float_as_uint(glevel));
}

// modulated cutoff, constant period: the consumer renders the chunk alone (same code as s2_kernels.cu)
template <int FILTER>
__device__ __noinline__ ConsState pc_solo_modcut(ColdPc& C, ConsState st, int wkind, bool active, float sr, float* row,
                                                 const float* sintab) {
    FastV<1> F;
    const OscC oc = C.oc;
    F.P = oc.P; F.d = oc.d; F.slope = oc.slope; F.nhalf = -oc.half; F.ts1 = oc.ts1; F.ts2 = oc.ts2;
    F.gain = C.gain; F.namt = C.namt;
    F.ph = st.ph; F.x1 = st.fs.x1; F.x2 = st.fs.x2; F.y1 = st.fs.y1; F.y2 = st.fs.y2;
    FiltC fc = C.fc;
    const float lpf = C.lpf, amt = C.amt_lpf, damp = C.damp;
    const uint32_t rot = C.rot, kind = C.kind, n = st.n;
    switch (wkind) {
    case 0: chunk_modcut<FILTER, 0, TRACE_NONE, false>(F, &C.amp, &C.mod, lpf, amt, damp, sr, fc, kind, rot, n, row, sintab, nullptr); break;
    case 1: chunk_modcut<FILTER, 1, TRACE_NONE, false>(F, &C.amp, &C.mod, lpf, amt, damp, sr, fc, kind, rot, n, row, sintab, nullptr); break;
    case 2: chunk_modcut<FILTER, 2, TRACE_NONE, false>(F, &C.amp, &C.mod, lpf, amt, damp, sr, fc, kind, rot, n, row, sintab, nullptr); break;
    default: chunk_modcut<FILTER, -1, TRACE_NONE, false>(F, &C.amp, &C.mod, lpf, amt, damp, sr, fc, kind, rot, n, row, sintab, nullptr); break;
    }
    if (active) C.fc = fc;
    st.ph = F.ph; st.fs.x1 = F.x1; st.fs.x2 = F.x2; st.fs.y1 = F.y1; st.fs.y2 = F.y2;
    st.n = n + kChunk;
    return st;
}

// anything else: the normative per-frame semantics (general_frame), x16 or scalar-tail flavour
template <int FILTER>
__device__ __noinline__ ConsState pc_solo_general(ColdPc& C, ConsState st, const float* __restrict__ Pcol, uint32_t vp, float sr,
                                                  uint32_t t0, uint32_t cnt, uint32_t f16, float* row, const float* sintab) {
    const Lane L = load_lane(Pcol, vp, sr);               // the general path needs the full envelopes
    OscC oc = C.oc;
    FiltC fc = C.fc;
    uint32_t n = st.n;
    float ph = st.ph;
    FiltS fs = st.fs;
    for (uint32_t i = 0; i < cnt; i++) {
        const bool scalar_sem = t0 + i >= f16;
        row[i] = general_frame<FILTER, TRACE_NONE>(L, sr, n, scalar_sem, oc, fc, ph, fs, sintab);
        n += 1u;
    }
    C.oc = oc;
    C.fc = fc;
    C.n_safe = 0u;
    st.n = n; st.ph = ph; st.fs = fs;
    return st;
}

}  // namespace

template <int FILTER>
__global__ void __launch_bounds__(kPcThreads, 14)
render_pc_kernel(const RenderArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int role = threadIdx.x >> 5;            // 0 consumer, 1 producer
    const int lane = threadIdx.x & 31;
    constexpr int kTileFloats = 32 * kTileStride;
    float* ubuf = smem;                           // two exchange tiles
    float* cold_base = smem + 2 * kTileFloats;
    PcCtrl* ctrl = reinterpret_cast<PcCtrl*>(cold_base + 32 * kPcColdWords);
    float* sintab = smem + kPcSmemFloats;
    if (a.has_sine) {
        for (int i = threadIdx.x; i < 1024; i += kPcThreads) sintab[i] = __uint_as_float(d_sin_bits[i]);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < MB_COUNT; i++) mbar_init(&ctrl->bars[i], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long* bars = ctrl->bars;

    const uint32_t vbase = a.slot_begin + blockIdx.x * 32u;
    if (vbase >= a.slot_end) return;
    const uint32_t gwarp = vbase / 32u;
    ColdPc& C = *reinterpret_cast<ColdPc*>(cold_base + lane * kPcColdWords);

    // =============================================================== producer
    if (role == 1) {
        uint32_t par_run = 0, par_empty = 0;   // bit b = parity to wait for on EMPTY[b]
        for (;;) {
            mbar_wait(&bars[MB_RUN], par_run);
            par_run ^= 1u;
            if (ctrl->exit) return;
            const uint32_t chunks = ctrl->chunks;
            const int wkind = ctrl->wkind;
            const bool fastvar = ctrl->namt0 && ctrl->aligned8;
            const OscC o = C.oc;
            const float gain = C.gain, namt = C.namt;
            const uint32_t kind = C.kind, rot = C.rot;
            float ph = C.ph;
            uint32_t n = C.n;
            for (uint32_t c = 0; c < chunks; c++) {
                const int b = (int)(c & 1u);
                if (c >= 2u) { mbar_wait(&bars[MB_EMPTY0 + b], (par_empty >> b) & 1u); par_empty ^= 1u << b; }
                float* row = ubuf + b * kTileFloats + lane * kTileStride;
                if (fastvar) produce_dispatch<true, true>(wkind, o, gain, namt, kind, rot, n, ph, row, sintab);
                else produce_dispatch<false, false>(wkind, o, gain, namt, kind, rot, n, ph, row, sintab);
                if (c + 1u == chunks) C.ph = ph;                  // hand the phase back with the last tile
                mbar_arrive_warp(&bars[MB_FULL0 + b], lane);
            }
        }
    }

    // =============================================================== consumer
    const uint32_t vp = a.vpad;
    const float sr = a.sample_rate;
    const uint32_t v = vbase + lane;
    const bool exists = v < a.slot_end;
    const uint32_t vi = exists ? v : vbase;       // out-of-range lanes shadow slot vbase's loads, never store
    bool active;
    ConsState st;                                 // n, phase, filter state, current filter constants, gain level
    st.glevel = 0.0f;
    const float* __restrict__ Pcol = a.params + vi;
    const uint32_t my_out_row = exists ? __float_as_uint(Pcol[P_ROW * vp]) : 0xffffffffu;
    {
        active = exists && __float_as_uint(Pcol[P_ACTIVE * vp]) != 0u;
        const Lane L = load_lane(Pcol, vp, sr);
        C.kind = L.kind; C.rot = L.rot; C.pitch = L.pitch; C.gain = L.gain; C.namt = L.namt;
        C.lpf = L.lpf; C.damp = L.damp; C.amt_osc = L.amt_osc; C.amt_lpf = L.amt_lpf;
        C.amp = compact(L.amp);
        C.mod = compact(L.mod);
        C.flags = (active ? 1u : 0u) | ((L.amt_osc != 0.0f || L.amt_lpf != 0.0f) ? 2u : 0u);
        const float* __restrict__ S = a.state + vi;
        st.ph = __float_as_uint(S[S_HAS_PHASE * vp]) != 0u ? S[S_PHASE * vp] : 0.0f;   // process.rs:316
        st.n = __float_as_uint(S[S_OFFSET * vp]);
        if (FILTER == 0) { st.fs.y1 = S[S_LAST * vp]; st.fs.x1 = st.fs.x2 = st.fs.y2 = 0.0f; }
        else { st.fs.x1 = S[S_X1 * vp]; st.fs.x2 = S[S_X2 * vp]; st.fs.y1 = S[S_Y1 * vp]; st.fs.y2 = S[S_Y2 * vp]; }
        OscC oc;
        oc.fo_bits = __float_as_uint(S[S_FO_KEY * vp]);
        oc.P = S[S_OSC_P * vp]; oc.d = S[S_OSC_D * vp]; oc.slope = S[S_OSC_SLOPE * vp];
        oc.half = S[S_OSC_HALF * vp]; oc.ts1 = S[S_OSC_TS1 * vp]; oc.ts2 = S[S_OSC_TS2 * vp];
        C.oc = oc;
        FiltC fc;
        fc.fl_bits = __float_as_uint(S[S_DAMP_KEY * vp]) == __float_as_uint(L.damp) ? __float_as_uint(S[S_FL_KEY * vp]) : kNoKey;
        fc.c0 = S[S_FC_C0 * vp]; fc.c1 = S[S_FC_C1 * vp]; fc.c2 = S[S_FC_C2 * vp];
        C.fc = fc;
        st.fcr = fc;
        C.n_safe = 0u;
        C.n_gc = 0u;
    }
    const uint32_t kind = C.kind;
    const uint32_t amask = __ballot_sync(0xffffffffu, active);
    int wkind = -1;
    {
        const int leader = amask ? __ffs(amask) - 1 : 0;
        const uint32_t k0 = __shfl_sync(0xffffffffu, kind, leader);
        if (__all_sync(0xffffffffu, !active || kind == k0)) wkind = (int)k0;
    }
    const bool namt0 = __all_sync(0xffffffffu, !active || __float_as_uint(C.namt) == 0u);
    const bool aligned8 = __all_sync(0xffffffffu, !active || (st.n & 7u) == 0u);

    const uint32_t frames = a.frames;
    const uint32_t f16 = frames & ~15u;
    const size_t stride = a.row_stride;
    float* __restrict__ gout = a.voice_out;
    const int q = lane >> 3, c4 = (lane & 7) * 4;
    const bool all_rows = __all_sync(0xffffffffu, gout != nullptr && my_out_row != 0xffffffffu);
    float* __restrict__ gbus = a.bus_partials ? a.bus_partials + (size_t)gwarp * frames : nullptr;

    uint32_t fast_left = 0, gc_left = 0;
    uint32_t par_full = 0;          // bit b = parity to wait for on FULL[b]
    uint32_t run_left = 0;          // chunks of the current producer run still to consume
    uint32_t run_c = 0;             // index of the next chunk inside the run
    uint32_t run_len = 0;

    for (uint32_t t0 = 0; t0 < frames; t0 += kChunk) {
        const uint32_t cnt = min((uint32_t)kChunk, frames - t0);
        const bool full = cnt == kChunk && t0 + kChunk <= f16;
        bool warp_fast = full && fast_left >= (uint32_t)kChunk;
        bool warp_semi = false;
        if (!warp_fast) {
            // (a run never outlives fast_left, so no run is open here)
            bool fast = active && full, semi = false;
            if (fast && st.n + kChunk > C.n_safe) {
                const ClassifyOut r = pc_classify<FILTER>(C, st.n, sr, st.fcr, st.glevel);
                st.fcr = r.fcr;
                st.glevel = r.glevel;
                fast = (r.res & 1u) != 0u;
                semi = (r.res & 2u) != 0u;
            }
            warp_fast = full && amask != 0u && __all_sync(0xffffffffu, fast || !active);
            warp_semi = !warp_fast && full && amask != 0u && __all_sync(0xffffffffu, fast || semi || !active);
            if (warp_fast) {
                fast_left = __reduce_min_sync(0xffffffffu, active ? C.n_safe - st.n : 0xffffffffu);
                gc_left = 0;
            } else {
                fast_left = 0;
            }
        }
        if (warp_fast && gc_left < (uint32_t)kChunk) {
            uint32_t lane_gc = 0xffffffffu;
            if (active) {
                const uint2 r = pc_refresh_gc(C, st.n, st.glevel);
                lane_gc = r.x;
                st.glevel = __uint_as_float(r.y);
            }
            gc_left = __reduce_min_sync(0xffffffffu, lane_gc);
        }

        float* tile = ubuf;                                       // solo paths render into tile 0
        if (warp_fast) {
            if (run_left == 0u) {
                // open a run: every fast chunk from here until the constants expire or the render ends
                const uint32_t avail = (f16 - t0) / kChunk;
                run_len = min(fast_left / kChunk, avail);
                run_left = run_len;
                run_c = 0;
                C.ph = st.ph;
                C.n = st.n;
                if (lane == 0) {
                    ctrl->chunks = run_len;
                    ctrl->exit = 0u;
                    ctrl->wkind = wkind;
                    ctrl->namt0 = namt0 ? 1u : 0u;
                    ctrl->aligned8 = aligned8 ? 1u : 0u;
                }
                mbar_arrive_warp(&bars[MB_RUN], lane);
            }
            const int b = (int)(run_c & 1u);
            tile = ubuf + b * kTileFloats;
            mbar_wait(&bars[MB_FULL0 + b], (par_full >> b) & 1u);
            par_full ^= 1u << b;
            float* row = tile + lane * kTileStride;
            if (gc_left >= (uint32_t)kChunk) consume_chunk<FILTER, true>(st.fcr, st.fs, st.glevel, &C.amp, st.n, row);
            else consume_chunk<FILTER, false>(st.fcr, st.fs, st.glevel, &C.amp, st.n, row);
            st.n += kChunk;
            fast_left -= kChunk;
            gc_left = gc_left >= (uint32_t)kChunk ? gc_left - kChunk : 0u;
        } else if (warp_semi) {
            st = pc_solo_modcut<FILTER>(C, st, wkind, active, sr, tile + lane * kTileStride, sintab);
        } else if (active) {
            st = pc_solo_general<FILTER>(C, st, Pcol, vp, sr, t0, cnt, f16, tile + lane * kTileStride, sintab);
        }
        if (!active) {
            float* row = tile + lane * kTileStride;
#pragma unroll
            for (int j = 0; j < kChunk / 4; j++)
                *reinterpret_cast<float4*>(row + 4 * j) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        __syncwarp();
        if (gout) {
            if (cnt == kChunk) {
                // transposed write-back: lanes 8q..8q+7 cover 128 contiguous bytes of tile row 4*i + q
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t orow = __shfl_sync(0xffffffffu, my_out_row, 4 * i + q);
                    if (all_rows || orow != 0xffffffffu) {
                        const float4 val = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                        __stcs(reinterpret_cast<float4*>(gout + (size_t)orow * stride + t0 + c4), val);
                    }
                }
            } else {
                for (uint32_t r = 0; r < 32u; r++) {
                    const uint32_t orow = __shfl_sync(0xffffffffu, my_out_row, r);
                    if (orow != 0xffffffffu && (uint32_t)lane < cnt)
                        gout[(size_t)orow * stride + t0 + lane] = tile[r * kTileStride + lane];
                }
            }
        }
        if (gbus) {
            if ((uint32_t)lane < cnt) {
                float acc;
                if (a.n_voices <= 32u) {
                    acc = 0.0f;                                   // synth.rs:176-202, the reference's order
#pragma unroll 8
                    for (int r = 0; r < 32; r++) acc = __fadd_rn(acc, tile[r * kTileStride + lane]);
                } else {
                    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
                    for (int r = 0; r < 32; r += 4) {
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
        __syncwarp();      // every lane is done reading the tile
        if (warp_fast) {
            // give the tile back if the producer still has chunks to write into it, close the run otherwise
            if (run_c + 2u < run_len) mbar_arrive_warp(&bars[MB_EMPTY0 + (int)(run_c & 1u)], lane);
            run_c++;
            run_left--;
            if (run_left == 0u) st.ph = C.ph;                     // written before the last FULL arrive
        }
    }
    // no run is open here: tell the producer to leave
    if (lane == 0) ctrl->exit = 1u;
    mbar_arrive_warp(&bars[MB_RUN], lane);

    if (active) {
        float* __restrict__ S = a.state + vi;
        S[S_PHASE * vp] = st.ph;
        S[S_HAS_PHASE * vp] = __uint_as_float(1u);
        const uint32_t start = __float_as_uint(S[S_OFFSET * vp]);
        const uint32_t nxt = start + frames < start ? 0xffffffffu : start + frames;   // saturating (synth.rs:197)
        S[S_OFFSET * vp] = __uint_as_float(nxt);
        if (FILTER == 0) S[S_LAST * vp] = st.fs.y1;
        else { S[S_X1 * vp] = st.fs.x1; S[S_X2 * vp] = st.fs.x2; S[S_Y1 * vp] = st.fs.y1; S[S_Y2 * vp] = st.fs.y2; }
        const OscC oc = C.oc;
        const FiltC fc = C.fc;
        S[S_FO_KEY * vp] = __uint_as_float(oc.fo_bits);
        S[S_OSC_P * vp] = oc.P; S[S_OSC_D * vp] = oc.d; S[S_OSC_SLOPE * vp] = oc.slope;
        S[S_OSC_HALF * vp] = oc.half; S[S_OSC_TS1 * vp] = oc.ts1; S[S_OSC_TS2 * vp] = oc.ts2;
        S[S_FL_KEY * vp] = __uint_as_float(fc.fl_bits);
        S[S_DAMP_KEY * vp] = C.damp;
        S[S_FC_C0 * vp] = fc.c0; S[S_FC_C1 * vp] = fc.c1; S[S_FC_C2 * vp] = fc.c2;
    }
}

template <int FILTER>
static cudaError_t launch_pc_t(const RenderArgs& a, cudaStream_t stream) {
    const uint32_t blocks = (a.slot_end - a.slot_begin + 31u) / 32u;
    const size_t smem = kPcSmemFloats * sizeof(float) + (a.has_sine ? 4096 : 0);
    static bool attr_set = false;
    if (!attr_set) {
        // 14 blocks x ~15 KB must be resident per SM: ask for the largest shared-memory carveout
        cudaFuncSetAttribute(render_pc_kernel<FILTER>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    render_pc_kernel<FILTER><<<blocks, kPcThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_render_pc(const RenderArgs& a, uint32_t filter_kind, cudaStream_t stream) {
    if (a.n_voices == 0 || a.frames == 0) return cudaSuccess;
    return filter_kind == 0 ? launch_pc_t<0>(a, stream) : launch_pc_t<1>(a, stream);
}

}  // namespace s2
