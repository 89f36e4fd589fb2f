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
#include "s2_device.cuh"

// Register cap of the one-voice-per-lane kernel.  96 keeps the fast loops spill-free and lets 16+ one-warp
// blocks share an SM, so blocks of overlapping launches (pipelined mode) find room: measured 1.079e12
// voice-samples/s at 96 against 1.047e12 at 128 and 0.967e12 at 80 (spills).
#ifndef S2_MAXREG
#define S2_MAXREG 96
#endif

namespace s2 {

// One warp renders 32*NV consecutive slots; lane l owns slots base + l (+ 32 for its second voice).
// 65,536 voices = 13.8 one-warp blocks per SM: without pipelining all of them must be resident at once (a
// second wave would serialise: measured +20 % time at 144 registers, where only 13 fit).
// Sum of a lane's N float4 tile reads as a pairwise tree of packed adds (depth log2 N instead of a chain of
// N: the adds sit at the end of a chunk where the warp has nothing else to issue).  Fixed order: deterministic.
template <int N>
__device__ __forceinline__ void tree_sum_rows(float4 (&val)[N], float2& b01, float2& b23) {
#pragma unroll
    for (int w = 1; w < N; w <<= 1) {
#pragma unroll
        for (int i = 0; i + w < N; i += 2 * w) {
            const float2 lo = padd2(make_float2(val[i].x, val[i].y), make_float2(val[i + w].x, val[i + w].y));
            const float2 hi = padd2(make_float2(val[i].z, val[i].w), make_float2(val[i + w].z, val[i + w].w));
            val[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
    }
    b01 = make_float2(val[0].x, val[0].y);
    b23 = make_float2(val[0].z, val[0].w);
}

template <int NV, int FILTER, int TRACE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) __maxnreg__(NV == 1 ? S2_MAXREG : 255)
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

    const uint32_t vbase = a.slot_begin + (blockIdx.x * kWarpsPerBlock + warp) * kRows;
    if (vbase >= a.slot_end) return;
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
        const bool exists = v < a.slot_end;
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
        lane_al8 &= !active[e] || ((n[e] | rot[e]) & (uint32_t)(S2_TRIP - 1)) == 0u;
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

    // Do the active voices of the warp share one cutoff trajectory (cutoff, damping, modulation amount, mod
    // envelope, frame offset)?  Then a moving-cutoff chunk computes its 32 frames' coefficients once, one frame
    // per lane (chunk_modcut<..., SHARED>).  Offsets advance together, so this holds for the launch.
    bool filt_uniform = false;
    if (NV == 1) {
        const unsigned am = __ballot_sync(0xffffffffu, active[0]);
        const int uni_leader = am ? __ffs(am) - 1 : 0;
        const Cold& C = cold(0);
        const EnvP& M = C.L.mod;
        const float key[10] = {C.L.lpf, C.L.damp, C.L.amt_lpf, M.A, M.AD, M.S, M.Rs, M.E, M.sD, M.sR};
        // (every lane must execute every shuffle: no short-circuit between them)
        const uint32_t n_lead = __shfl_sync(0xffffffffu, n[0], uni_leader);
        const float sa_lead = __shfl_sync(0xffffffffu, M.sA, uni_leader);
        bool same = n_lead == n[0];
        same &= __float_as_uint(sa_lead) == __float_as_uint(M.sA);
#pragma unroll
        for (int k = 0; k < 10; k++)
            same &= __float_as_uint(__shfl_sync(0xffffffffu, key[k], uni_leader)) == __float_as_uint(key[k]);
        filt_uniform = am != 0u && __all_sync(0xffffffffu, same || !active[0]);
    }

    const uint32_t frames = a.frames;
    const uint32_t f16 = frames & ~15u;            // x16 region (process.rs:26-37), then the scalar tail
    const size_t stride = a.row_stride;
    float* __restrict__ gout = a.voice_out;
    // Output rows are indexed by the caller's voice index, which the bank may have permuted into
    // kind-uniform warps (P_ROW).  Transposed write-back: lanes 8q..8q+7 cover 128 contiguous bytes of
    // tile row 4*i + q, so each STG.128 of the warp writes four full 128-byte lines.
    const int q = lane >> 3, c4 = (lane & 7) * 4;
    // rp[tile row] = base address of that voice's output row (0 = none): 32*NV 64-bit words per warp.  The four
    // q-groups of a store read four neighbouring words (broadcast inside a group): conflict-free LDS.64.
    unsigned long long* rp = reinterpret_cast<unsigned long long*>(cold_base + kRows * kColdWords);
    bool lane_rows_ok = gout != nullptr;
#pragma unroll
    for (int e = 0; e < NV; e++) {
        const uint32_t r = cold(e).out_row;
        lane_rows_ok &= r != 0xffffffffu;
        rp[e * 32 + lane] = (gout && r != 0xffffffffu) ? reinterpret_cast<unsigned long long>(gout + (size_t)r * stride) : 0ull;
    }
    __syncwarp();
    const bool all_rows = __all_sync(0xffffffffu, lane_rows_ok);   // every tile row has an output row

    // fast_left: frames for which every voice of the warp stays on the fast path (warp-uniform), with
    // the loop variant (gconst) chosen when it was computed; 0 = classify before the next chunk.
    uint32_t fast_left = 0, gc_left = 0;
    bool nv2_gconst = false;

    // Runs of fast tiles.  A third of the warp's stall time sits at the chunk boundaries (loop control, dispatch,
    // store-variant selection: uniform-datapath code with exposed latencies, and the warps of an SM reach it in
    // lockstep), so when the store has its simple launch-uniform form — every tile row has an output row or no
    // rows are wanted, and the mix, if any, is the per-warp partial of a wide bank — the fast path renders up to
    // kRunTiles tiles per trip round the outer loop, each followed by its own straight-line write-back (run lengths
    // 2 / 4 / 8 / 16 measured 222 / 221 / 217 / 217 us per 65,536 x 4,096 block).
#ifndef S2_RUN_TILES
#define S2_RUN_TILES 8
#endif
    constexpr uint32_t kRunTiles = S2_RUN_TILES;
    const bool bus_wide_ok = a.bus_partials != nullptr && a.n_voices > 32u && (frames & 3u) == 0u &&
                             (reinterpret_cast<uintptr_t>(a.bus_partials) & 15u) == 0u;
    const bool simple_store = NV == 1 && (gout == nullptr || all_rows) && (a.bus_partials == nullptr || bus_wide_ok) &&
                              (gout != nullptr || a.bus_partials != nullptr);
    auto store_simple = [&](uint32_t ts) {
        const size_t tb = ((size_t)ts + (size_t)c4) * sizeof(float);
        if (a.bus_partials == nullptr) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                char* dst = reinterpret_cast<char*>(rp[4 * i + q]);
                const float4 val = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                __stcs(reinterpret_cast<float4*>(dst + tb), val);
            }
        } else {
            float4 val[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                val[i] = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                if (gout) __stcs(reinterpret_cast<float4*>(reinterpret_cast<char*>(rp[4 * i + q]) + tb), val[i]);
            }
            float2 s01, s23;
            tree_sum_rows<8>(val, s01, s23);
#pragma unroll
            for (int sh = 8; sh <= 16; sh <<= 1) {
                s01 = padd2(s01, make_float2(__shfl_xor_sync(0xffffffffu, s01.x, sh), __shfl_xor_sync(0xffffffffu, s01.y, sh)));
                s23 = padd2(s23, make_float2(__shfl_xor_sync(0xffffffffu, s23.x, sh), __shfl_xor_sync(0xffffffffu, s23.y, sh)));
            }
            float* gb = a.bus_partials + (size_t)(a.slot_begin / (uint32_t)kRows + blockIdx.x * (uint32_t)kWarpsPerBlock + (uint32_t)warp) * a.frames;
            if (lane < 8) *reinterpret_cast<float4*>(gb + ts + c4) = make_float4(s01.x, s01.y, s23.x, s23.y);
        }
    };

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
            uint32_t reps = 1;
            if (simple_store) {
                uint32_t room = min(fast_left, f16 - t0) / (uint32_t)kChunk;      // >= 1: warp_fast
                if (gconst) room = min(room, gc_left / (uint32_t)kChunk);          // >= 1: gconst
                reps = min(room, kRunTiles);
            }
            for (uint32_t r = 0;;) {
                switch (wkind) {
                case 0: chunk_fast_dispatch<NV, FILTER, 0, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
                case 1: chunk_fast_dispatch<NV, FILTER, 1, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
                case 2: chunk_fast_dispatch<NV, FILTER, 2, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
                case 3: chunk_fast_dispatch<NV, FILTER, 3, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
                default: chunk_fast_dispatch<NV, FILTER, -1, TRACE>(gconst, namt0, aligned8, F, &cold(0).L.amp, kind, rot, n, tile, lane, sintab); break;
                }
#pragma unroll
                for (int e = 0; e < NV; e++) n[e] += kChunk;
                if (!simple_store) break;                    // one tile, written back by the common code below
                if (!active[0]) {
                    float* row = tile + lane * kTileStride;
#pragma unroll
                    for (int j = 0; j < kChunk / 4; j++)
                        *reinterpret_cast<float4*>(row + 4 * j) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
                __syncwarp();
                store_simple(t0 + r * (uint32_t)kChunk);
                __syncwarp();      // every lane is done reading the tile before the next one overwrites it
                if (++r == reps) break;
            }
            fast_left -= reps * (uint32_t)kChunk;
            gc_left = gc_left >= reps * (uint32_t)kChunk ? gc_left - reps * (uint32_t)kChunk : 0u;
            if (simple_store) {
                t0 += (reps - 1u) * (uint32_t)kChunk;        // the loop header adds the last tile
                continue;
            }
        } else if (warp_semi) {
            if constexpr (NV == 1) {
                // voices that are fully constant run the same code: their cutoff simply does not move
                Cold& C = cold(0);
                FiltC fc = C.fc;
                float* row = tile + lane * kTileStride;
                const float lpf = C.L.lpf, amt = C.L.amt_lpf, damp = C.L.damp;
                if (filt_uniform) {
                    // one cutoff trajectory for the whole warp: 32 frames' coefficients, one per lane, once
                    float* ctab = cold_base + kRows * kColdWords + kRows * kRowPtrWords;
                    const int uni_leader = __ffs(__ballot_sync(0xffffffffu, active[0])) - 1;      // filt_uniform => some lane is active
                    const Cold& CL = *reinterpret_cast<const Cold*>(cold_base + uni_leader * kColdWords);
                    const uint32_t n_lead = __shfl_sync(0xffffffffu, n[0], uni_leader);   // inactive lanes hold other offsets
                    modcut_coefficients<FILTER>(CL.L.mod, CL.L.lpf, CL.L.amt_lpf, CL.L.damp, sr, n_lead, lane, ctab);
                    __syncwarp();
                    switch (wkind) {
                    case 0: chunk_modcut<FILTER, 0, TRACE, true>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab, ctab); break;
                    case 1: chunk_modcut<FILTER, 1, TRACE, true>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab, ctab); break;
                    case 2: chunk_modcut<FILTER, 2, TRACE, true>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab, ctab); break;
                    default: chunk_modcut<FILTER, -1, TRACE, true>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab, ctab); break;
                    }
                    __syncwarp();          // ctab is rewritten by the next moving-cutoff chunk
                } else {
                switch (wkind) {
                case 0: chunk_modcut<FILTER, 0, TRACE, false>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab, nullptr); break;
                case 1: chunk_modcut<FILTER, 1, TRACE, false>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab, nullptr); break;
                case 2: chunk_modcut<FILTER, 2, TRACE, false>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab, nullptr); break;
                default: chunk_modcut<FILTER, -1, TRACE, false>(F, &C.L.amp, &C.L.mod, lpf, amt, damp, sr, fc, kind[0], rot[0], n[0], row, sintab, nullptr); break;
                }
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
        // (16-byte stores into this warp's partial row: needs frames % 4 == 0 and an aligned base)
        // this warp's partial row, from launch parameters and the block index only (uniform registers: a pointer
        // kept live across the chunk loop is spilled under the register cap and reloaded from local memory)
        float* __restrict__ gbus = a.bus_partials
            ? a.bus_partials + (size_t)(a.slot_begin / (uint32_t)kRows + blockIdx.x * (uint32_t)kWarpsPerBlock + (uint32_t)warp) * a.frames
            : nullptr;
        const bool wide_bus = gbus != nullptr && a.n_voices > 32u && cnt == kChunk && (frames & 3u) == 0u &&
                              (reinterpret_cast<uintptr_t>(gbus) & 15u) == 0u;
        float2 b01 = make_float2(0.0f, 0.0f), b23 = make_float2(0.0f, 0.0f);
        if (cnt == kChunk && (gout || wide_bus)) {
            // One pass over the tile, transposed: lane (q, c4) reads 16 bytes of rows 4*i + q.  For banks
            // wider than a warp the same registers also feed the bus: each lane adds its 8*NV rows, then the
            // four q-groups are added by two butterfly shuffles (a fixed tree: deterministic; banks of <= 32
            // voices use the reference's sequential order below).  Three straight-line variants so that the
            // common one carries no predicates.
            const size_t tb = ((size_t)t0 + (size_t)c4) * sizeof(float);
            if (all_rows && !wide_bus) {
#pragma unroll
                for (int i = 0; i < 8 * NV; i++) {
                    char* dst = reinterpret_cast<char*>(rp[4 * i + q]);
                    const float4 val = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                    __stcs(reinterpret_cast<float4*>(dst + tb), val);
                }
            } else if (all_rows) {
                float4 val[8 * NV];
#pragma unroll
                for (int i = 0; i < 8 * NV; i++) {
                    char* dst = reinterpret_cast<char*>(rp[4 * i + q]);
                    val[i] = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                    __stcs(reinterpret_cast<float4*>(dst + tb), val[i]);
                }
                tree_sum_rows<8 * NV>(val, b01, b23);
            } else {
                float4 val[8 * NV];
#pragma unroll
                for (int i = 0; i < 8 * NV; i++) {
                    val[i] = *reinterpret_cast<const float4*>(tile + (4 * i + q) * kTileStride + c4);
                    if (gout) {
                        char* dst = reinterpret_cast<char*>(rp[4 * i + q]);
                        if (dst) __stcs(reinterpret_cast<float4*>(dst + tb), val[i]);
                    }
                }
                tree_sum_rows<8 * NV>(val, b01, b23);
            }
        } else if (gout) {
#pragma unroll
            for (int e = 0; e < NV; e++) {
                for (uint32_t r = 0; r < 32u; r++) {
                    const uint32_t orow = __shfl_sync(0xffffffffu, cold(e).out_row, r);
                    if (orow != 0xffffffffu && (uint32_t)lane < cnt)
                        gout[(size_t)orow * stride + t0 + lane] = tile[(e * 32 + r) * kTileStride + lane];
                }
            }
        }
        if (wide_bus) {
#pragma unroll
            for (int sh = 8; sh <= 16; sh <<= 1) {
                b01 = padd2(b01, make_float2(__shfl_xor_sync(0xffffffffu, b01.x, sh), __shfl_xor_sync(0xffffffffu, b01.y, sh)));
                b23 = padd2(b23, make_float2(__shfl_xor_sync(0xffffffffu, b23.x, sh), __shfl_xor_sync(0xffffffffu, b23.y, sh)));
            }
            if (lane < 8) *reinterpret_cast<float4*>(gbus + t0 + c4) = make_float4(b01.x, b01.y, b23.x, b23.y);
        } else if (gbus) {
            if ((uint32_t)lane < cnt) {
                float acc;
                if (a.n_voices <= 32u) {
                    // synth.rs:176-202: voices are accumulated in index order, starting from 0.0 — the
                    // reference's exact summation order (a bank this narrow is one warp, identity slots)
                    acc = 0.0f;
#pragma unroll 8
                    for (int r = 0; r < kRows; r++) acc = __fadd_rn(acc, tile[r * kTileStride + lane]);
                } else {
                    // ragged last chunk of a wide bank: fixed 4-way tree over the warp's rows
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

// Bus reduction: one kernel, fixed summation order (deterministic).  Block fx owns 32 frames; its 8 warps take the
// partial rows w, w + 8, w + 16, ... (lane = frame: 128-byte coalesced reads), each with four accumulators over
// consecutive rows of its sequence, eight loads in flight; the eight warp sums are then added in order.
// Few, long-lived blocks on purpose: the reduction runs next to the render kernels of the following step on a
// machine whose every slot is taken by long-running render blocks.  The 4,096 short blocks of the two-stage
// form it replaces queued for those slots one by one and cost ~30 us of step time for 19 us of work; 128 blocks
// of 8 warps need one slot each and displace next to nothing (it is latency-bound at ~20 us, off the critical path).
constexpr uint32_t kBusSegWarps = 64;      // (scratch sizing of the callers: segments of the former first stage)

__global__ void __launch_bounds__(256) bus_reduce_kernel(const float* __restrict__ partials, uint32_t n_rows,
                                                         size_t row_stride, uint32_t frames, float* __restrict__ bus) {
    __shared__ float sm[8][33];
    const uint32_t tx = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * 32u + tx;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    if (t < frames) {
        const float* __restrict__ col = partials + t;
        uint32_t r = w;
        for (; r + 56u < n_rows; r += 64u) {                  // rows r, r+8, ..., r+56: eight loads in flight
            float v[8];
#pragma unroll
            for (uint32_t k = 0; k < 8u; k++) v[k] = col[(size_t)(r + 8u * k) * row_stride];
            a0 = __fadd_rn(a0, v[0]); a1 = __fadd_rn(a1, v[1]); a2 = __fadd_rn(a2, v[2]); a3 = __fadd_rn(a3, v[3]);
            a0 = __fadd_rn(a0, v[4]); a1 = __fadd_rn(a1, v[5]); a2 = __fadd_rn(a2, v[6]); a3 = __fadd_rn(a3, v[7]);
        }
        for (uint32_t k = 0; r < n_rows; r += 8u, k++) {     // the rest, same rotation over the accumulators
            const float v = col[(size_t)r * row_stride];
            if ((k & 3u) == 0u) a0 = __fadd_rn(a0, v);
            else if ((k & 3u) == 1u) a1 = __fadd_rn(a1, v);
            else if ((k & 3u) == 2u) a2 = __fadd_rn(a2, v);
            else a3 = __fadd_rn(a3, v);
        }
    }
    sm[w][tx] = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
    __syncthreads();
    if (w == 0 && t < frames) {
        float total = sm[0][tx];
#pragma unroll
        for (int k = 1; k < 8; k++) total = __fadd_rn(total, sm[k][tx]);
        bus[t] = total;
    }
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
    const uint32_t n_warps = render_warps(a.slot_end - a.slot_begin, NV);
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

template <int NV, int FILTER>
static cudaError_t launch_f(const RenderArgs& a, int trace, cudaStream_t stream) {
    return trace == TRACE_PHASE ? launch_t<NV, FILTER, TRACE_PHASE>(a, stream) : launch_t<NV, FILTER, TRACE_NONE>(a, stream);
}

cudaError_t launch_render(const RenderArgs& a, uint32_t filter_kind, int trace, int nv, cudaStream_t stream) {
    if (a.n_voices == 0 || a.frames == 0) return cudaSuccess;
    switch (filter_kind) {
    case FILT_ONE_POLE: return nv == 2 ? launch_f<2, FILT_ONE_POLE>(a, trace, stream) : launch_f<1, FILT_ONE_POLE>(a, trace, stream);
    case FILT_BIQUAD_LP: return nv == 2 ? launch_f<2, FILT_BIQUAD_LP>(a, trace, stream) : launch_f<1, FILT_BIQUAD_LP>(a, trace, stream);
    // the remaining dsp_filters.rs filters: one voice per lane only
    case FILT_BIQUAD_HP: return launch_f<1, FILT_BIQUAD_HP>(a, trace, stream);
    case FILT_BIQUAD_BP: return launch_f<1, FILT_BIQUAD_BP>(a, trace, stream);
    case FILT_FIRST_LP: return launch_f<1, FILT_FIRST_LP>(a, trace, stream);
    case FILT_FIRST_HP: return launch_f<1, FILT_FIRST_HP>(a, trace, stream);
    default: return cudaErrorInvalidValue;
    }
}

uint32_t bus_segments(uint32_t n_warps) { return (n_warps + kBusSegWarps - 1) / kBusSegWarps; }

cudaError_t launch_bus_reduce(const float* partials, uint32_t n_warps, size_t row_stride, uint32_t frames,
                              float* /*seg_scratch*/, float* bus, cudaStream_t stream) {
    if (frames == 0) return cudaSuccess;
    bus_reduce_kernel<<<(frames + 31) / 32, 256, 0, stream>>>(partials, n_warps, row_stride, frames, bus);
    return cudaGetLastError();
}

}  // namespace s2
