// s2_capi.cu — host side of libs2cuda.so: the C ABI declared in include/s2_cuda.h.
//
// Two objects:
//   s2_bank  — V voices resident on one GPU (SoA parameters + carried state), rendered by
//              s2::launch_render; the batched form of process::process_layer_buf_simd.
//   s2_synth — mirror of synth::Synth (s2_lib/src/try3/synth.rs): 8 voice slots, the default
//              patch, oldest-voice allocation, note_on / note_off / sample.  The voice bookkeeping
//              runs on the host exactly like the reference's; only rendering goes to the GPU.
// No CPU rendering exists in this library.
#include "../../include/s2_cuda.h"
#include "s2_internal.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(S2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                \
    } while (0)

bool finite_nonneg(float x) { return std::isfinite(x) && x >= 0.0f; }

}  // namespace

namespace s2 {
// error reporting for the other host translation units of the library (s2_patch.cpp)
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace s2

namespace {

int validate_voice(const s2_voice_desc& d, size_t index, uint32_t filter_kind = S2_FILTER_ONE_POLE) {
    // second-order filters: damping (band-pass: the quality factor) divides or scales the coefficients
    // (dsp_filters.rs:99-109, 197-207; 0.2 is the documented minimum, :94-96); zero gives tan(inf) / a zero denominator
    if (filter_kind >= S2_FILTER_BIQUAD_LP && filter_kind <= S2_FILTER_BIQUAD_BP && !(d.damping > 0.0f))
        return fail(S2_ERR_INVALID, "voice %zu: damping must be > 0 for the second-order filters", index);
    if (d.osc_kind > S2_OSC_SINE) return fail(S2_ERR_INVALID, "voice %zu: osc_kind %u", index, d.osc_kind);
    if (!(std::isfinite(d.pitch_hz) && d.pitch_hz > 0.0f))
        return fail(S2_ERR_INVALID, "voice %zu: pitch_hz must be finite and > 0", index);
    if (!(std::isfinite(d.lpf_freq_hz) && d.lpf_freq_hz >= 0.0f))
        return fail(S2_ERR_INVALID, "voice %zu: lpf_freq_hz must be finite and >= 0", index);
    const float ms[6] = {d.amp_attack_ms, d.amp_decay_ms, d.amp_release_ms,
                         d.mod_attack_ms, d.mod_decay_ms, d.mod_release_ms};
    for (float m : ms)
        if (!finite_nonneg(m)) return fail(S2_ERR_INVALID, "voice %zu: envelope times must be finite and >= 0", index);
    const float fl[7] = {d.osc_gain, d.noise_amt, d.damping, d.amp_sustain, d.mod_sustain,
                         d.mod_env_to_osc_freq, d.mod_env_to_lpf_freq};
    for (float f : fl)
        if (!std::isfinite(f)) return fail(S2_ERR_INVALID, "voice %zu: non-finite parameter", index);
    return S2_OK;
}

uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
float ubits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

void pack_voice(const s2_voice_desc& d, uint32_t row, float* col, size_t pitch) {
    using namespace s2;
    col[P_ROW * pitch] = ubits(row);
    col[P_KIND * pitch] = ubits(d.osc_kind);
    col[P_SEED * pitch] = ubits(d.noise_seed);
    col[P_PITCH * pitch] = d.pitch_hz;
    col[P_GAIN * pitch] = d.osc_gain;
    col[P_NOISE * pitch] = d.noise_amt;
    col[P_LPF * pitch] = d.lpf_freq_hz;
    col[P_DAMP * pitch] = d.damping;
    col[P_AA * pitch] = d.amp_attack_ms;
    col[P_AD * pitch] = d.amp_decay_ms;
    col[P_AS * pitch] = d.amp_sustain;
    col[P_AR * pitch] = d.amp_release_ms;
    col[P_MA * pitch] = d.mod_attack_ms;
    col[P_MD * pitch] = d.mod_decay_ms;
    col[P_MS * pitch] = d.mod_sustain;
    col[P_MR * pitch] = d.mod_release_ms;
    col[P_AMT_OSC * pitch] = d.mod_env_to_osc_freq;
    col[P_AMT_LPF * pitch] = d.mod_env_to_lpf_freq;
    col[P_RELEASE * pitch] = ubits(d.release_offset);
    col[P_ACTIVE * pitch] = ubits(d.active ? 1u : 0u);
}

void pack_state(const s2_voice_state& s, float* col, size_t pitch) {
    using namespace s2;
    col[S_PHASE * pitch] = s.phase;
    col[S_HAS_PHASE * pitch] = ubits(s.has_phase ? 1u : 0u);
    col[S_OFFSET * pitch] = ubits(s.frame_offset);
    col[S_LAST * pitch] = s.lpf_last;
    col[S_X1 * pitch] = s.x1;
    col[S_X2 * pitch] = s.x2;
    col[S_Y1 * pitch] = s.y1;
    col[S_Y2 * pitch] = s.y2;
    // derived-constant memo: empty (the kernel re-derives and refills it)
    for (int k = S_FO_KEY; k < S_COUNT; k++) col[k * pitch] = 0.0f;
    col[S_FO_KEY * pitch] = ubits(kNoKey);
    col[S_FL_KEY * pitch] = ubits(kNoKey);
}

struct VoiceBook {           // host mirror of what the device will hold, O(1) per render
    uint32_t start_offset;   // frame offset when the voice was (re)started
    uint64_t start_total;    // bank->total_frames at that moment
    uint32_t active;
    uint32_t osc_kind;
};

}  // namespace

// Partial-sum buffers of the pipelined mix.  The reduction of step i runs on the (high-priority) mix stream
// while the sub-banks already render step i+1; with only two buffers step i+2 would have to wait for it,
// and a reduction that is slow to get SM slots next to the render blocks then stalls the whole pipeline.
constexpr int kMixBufs = 4;
constexpr int kTsBufs = 4;                 // the phase pre-pass may run this many blocks ahead of the renders
constexpr size_t kTsMaxVoices = 16384;     // beyond this a bank fills the machine with one voice per lane

struct s2_bank {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t sample_rate = 0;
    uint32_t filter_kind = 0;
    size_t n_voices = 0;
    size_t vpad = 0;
    float* d_params = nullptr;
    float* d_state = nullptr;
    float* d_partials = nullptr;
    size_t partials_cap = 0;     // floats
    unsigned int* d_bus_counters = nullptr;   // s2::kBusCounters zeroed words of the wide bus reduction
    float* d_bus = nullptr;
    size_t bus_cap = 0;          // floats
    std::vector<VoiceBook> book;  // indexed by voice
    // "Slots": device arrays are indexed by slot, not by the caller's voice index.  Banks wider than
    // one warp are sorted by (active, oscillator kind, when the amp envelope rests) at creation so that the
    // 32 lanes of a warp run the same code for as long as possible, then dealt over the voice ranges; output
    // rows, state get/set and per-voice calls keep the caller's indices.  Banks of <= 32 voices keep
    // the identity order, which also keeps the bus sum in the reference's voice order.
    std::vector<uint32_t> slot_of_voice, voice_of_slot;
    bool identity = true;
    uint32_t force_path = 0;      // test hook (S2_FORCE_PATH, see RenderArgs)
    uint32_t* d_stage = nullptr;  // staging for the bulk note-off table when slots are permuted
    // Pipelined mode (s2_bank_set_pipeline): the bank is cut into n_sub contiguous slot ranges, each
    // rendered on its own internal stream.  Consecutive render calls then overlap across sub-banks (no
    // device-wide barrier between blocks), which keeps every SM sub-partition supplied with warps while
    // the slower ones finish (profiles/r1_notes.md: +14 % at 65,536 voices).
    int n_sub = 1;
    cudaStream_t sub[8] = {};
    cudaStream_t mix = nullptr;           // bus reduction and copies of the pipelined mode
    cudaStream_t upload = nullptr;        // note-off table uploads (independent of the mix stream)
    cudaEvent_t ev_gather[8][2] = {};     // sub-bank k has applied the table staged in buffer q
    cudaEvent_t ev_sub[8] = {};           // last work issued on sub-stream k
    cudaEvent_t ev_mix[kMixBufs] = {};    // reduction of partial buffer p finished
    cudaEvent_t ev_stage[2] = {};         // note-off table p has landed in its staging buffer
    cudaEvent_t ev_tail = nullptr;        // last work issued on the mix stream
    uint32_t* d_stage2[2] = {};
    float* d_partials2[kMixBufs] = {};
    size_t partials2_cap = 0;
    uint64_t step = 0, table_step = 0;
    int table_pending = -1;               // staging buffer index holding a table not yet applied, or -1
    // Time-split mode (s2_bank_set_time_split, s2_kernel_ts.cu): narrow one-pole banks render a block as 32
    // time segments per voice.  The phase pre-pass runs on its own stream, ahead of the renders.
    bool ts_enabled = false;
    bool ts_main_dirty = true;            // b->stream carries state writes the pre-pass stream has not seen
    cudaStream_t ts = nullptr;
    float* d_seg_phase[kTsBufs] = {};     // [n_voices][3][32] segment-start phases (and the two before) of block p
    cudaEvent_t ev_k1[kTsBufs] = {}, ev_k2[kTsBufs] = {}, ev_main = nullptr;
    uint64_t ts_step = 0, ts_blocks = 0;  // ts_blocks: blocks rendered through the time-split kernels
    std::vector<s2_voice_desc> descs;     // host copy of the voice descriptions (banks <= kTsMaxVoices only)
    uint64_t total_frames = 0;
    uint64_t max_offset = 0;     // upper bound of any active voice's frame offset
    size_t n_sine = 0;
};

namespace {

uint32_t current_offset(const s2_bank* b, size_t i) {
    const VoiceBook& vb = b->book[i];
    if (!vb.active) return vb.start_offset;
    const uint64_t o = (uint64_t)vb.start_offset + (b->total_frames - vb.start_total);
    return o > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)o;   // saturating_add, synth.rs:197
}

// Waits (on the host) for everything the bank has issued on its internal streams.
int bank_drain(s2_bank* b) {
    for (int k = 0; k < b->n_sub && b->n_sub > 1; k++) CUDA_TRY(cudaStreamSynchronize(b->sub[k]));
    if (b->mix) CUDA_TRY(cudaStreamSynchronize(b->mix));
    if (b->upload) CUDA_TRY(cudaStreamSynchronize(b->upload));
    if (b->ts) CUDA_TRY(cudaStreamSynchronize(b->ts));
    return S2_OK;
}

uint32_t sub_begin(const s2_bank* b, int k) {
    // contiguous slot ranges, multiples of 64 slots so that warps (32 or 64 slots) never straddle two
    const uint64_t groups = (b->n_voices + 63) / 64;
    const uint64_t g = groups * (uint64_t)k / (uint64_t)b->n_sub;
    const uint64_t s = g * 64;
    return (uint32_t)(s < b->n_voices ? s : b->n_voices);
}

int bank_render_pipelined(s2_bank* b, size_t frames, float* d_voice_out, size_t row_stride, float* d_bus_out) {
    const uint32_t n_warps = s2::render_warps((uint32_t)b->n_voices);
    const int p = (int)(b->step % (uint64_t)kMixBufs);
    float* partials = nullptr;
    if (d_bus_out) {
        const size_t need = ((size_t)n_warps + s2::bus_segments(n_warps)) * frames;
        if (need > b->partials2_cap) {
            int rc = bank_drain(b);
            if (rc) return rc;
            for (int i = 0; i < kMixBufs; i++) {
                if (b->d_partials2[i]) CUDA_TRY(cudaFree(b->d_partials2[i]));
                b->d_partials2[i] = nullptr;
            }
            b->partials2_cap = 0;
            for (int i = 0; i < kMixBufs; i++) CUDA_TRY(cudaMalloc(&b->d_partials2[i], need * sizeof(float)));
            b->partials2_cap = need;
        }
        partials = b->d_partials2[p];
    }
    s2::RenderArgs a;
    a.params = b->d_params;
    a.state = b->d_state;
    a.n_voices = (uint32_t)b->n_voices;
    a.vpad = (uint32_t)b->vpad;
    a.sample_rate = (float)b->sample_rate;
    a.frames = (uint32_t)frames;
    a.voice_out = d_voice_out;
    a.row_stride = row_stride;
    a.bus_partials = partials;
    a.has_sine = b->n_sine ? 1u : 0u;
    a.one = 1.0f;
    a.force_path = b->force_path;
    a.staged_release = nullptr;
    a.release_row = nullptr;
    uint32_t* release_row = reinterpret_cast<uint32_t*>(b->d_params + (size_t)s2::P_RELEASE * b->vpad);
    for (int k = 0; k < b->n_sub; k++) {
        a.slot_begin = sub_begin(b, k);
        a.slot_end = sub_begin(b, k + 1);
        if (a.slot_begin >= a.slot_end) continue;
        cudaStream_t sk = b->sub[k];
        a.staged_release = nullptr;
        a.release_row = nullptr;
        if (b->table_pending >= 0) {
            // the staged note-off table: this sub-bank's render applies it to its own slots first (kernel prologue),
            // ordered between its own renders
            CUDA_TRY(cudaStreamWaitEvent(sk, b->ev_stage[b->table_pending], 0));
            a.staged_release = b->d_stage2[b->table_pending];
            a.release_row = release_row;
        }
        if (d_bus_out && b->step >= (uint64_t)kMixBufs) CUDA_TRY(cudaStreamWaitEvent(sk, b->ev_mix[p], 0));   // partials[p] are free again
        CUDA_TRY(s2::launch_render(a, b->filter_kind, s2::TRACE_NONE, sk));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (b->table_pending >= 0) CUDA_TRY(cudaEventRecord(b->ev_gather[k][b->table_pending], sk));   // staging buffer read
        CUDA_TRY(cudaEventRecord(b->ev_sub[k], sk));
    }
    b->table_pending = -1;
    if (d_bus_out) {
        for (int k = 0; k < b->n_sub; k++) CUDA_TRY(cudaStreamWaitEvent(b->mix, b->ev_sub[k], 0));
        if (n_warps == 1) {
            CUDA_TRY(cudaMemcpyAsync(d_bus_out, partials, frames * sizeof(float), cudaMemcpyDeviceToDevice, b->mix));
        } else {
            CUDA_TRY(s2::launch_bus_reduce(partials, n_warps, frames, (uint32_t)frames, partials + (size_t)n_warps * frames,
                                           d_bus_out, b->mix, b->d_bus_counters, (size_t)s2::bus_segments(n_warps) * frames));
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        CUDA_TRY(cudaEventRecord(b->ev_mix[p], b->mix));
    }
    b->step++;
    b->total_frames += frames;
    b->max_offset += frames;
    return S2_OK;
}

// units.rs:44-53 in host f32 arithmetic (IEEE, no contraction: same bits as the device's make_env)
float host_ms_as_samples(float ms, float sr) { return sr * (ms / 1000.0f); }

// May this block go through the time-split kernels?  Every active voice must keep one period for the whole
// block (no pitch modulation).  Returns 0 = no, 1 = yes and every cutoff rests too (mod envelope unused by the
// cutoff, or in sustain / end until the block ends), 2 = yes but some cutoff follows a ramping mod envelope
// (per-frame coefficients).  Conservative: anything unsure renders the general way.
int ts_block_class(const s2_bank* b, size_t frames, const float* d_voice_out, const float* d_bus_out) {
    if (!b->ts_enabled || !d_voice_out) return 0;                               // the mix is summed from the rows
    (void)d_bus_out;
    if (frames < 1024 || (frames & 1023u) != 0 || frames > (1u << 24)) return 0;   // 32 segments of whole chunks
    const float sr = (float)b->sample_rate;
    int cls = 1;
    for (size_t i = 0; i < b->n_voices; i++) {
        if (!b->book[i].active) continue;
        const s2_voice_desc& d = b->descs[i];
        if (d.mod_env_to_osc_freq != 0.0f) return 0;
        const uint64_t n0 = current_offset(b, i);
        if (n0 + frames > (1ull << 24)) return 0;
        const float P = sr / d.pitch_hz;
        const float step = 1.0f / P;
        if (!(step < 1.0f && P > 1.0f)) return 0;
        if (d.mod_env_to_lpf_freq != 0.0f) {
            const float A = host_ms_as_samples(d.mod_attack_ms, sr), D = host_ms_as_samples(d.mod_decay_ms, sr);
            const float R = host_ms_as_samples(d.mod_release_ms, sr);
            const float AD = A + D;
            const float Rs = fmaxf((float)d.release_offset, AD);
            const float E = Rs + R;
            const float x0 = (float)(uint32_t)n0, x1 = (float)(uint32_t)(n0 + frames - 1);
            const bool rest_sustain = x0 >= A && x0 >= AD && x1 < Rs;     // stage 2 from first to last frame
            const bool rest_end = x0 >= A && x0 >= AD && x0 >= Rs && x0 >= E;
            if (!(rest_sustain || rest_end)) cls = 2;
        }
    }
    return cls;
}

int bank_render_time_split(s2_bank* b, size_t frames, float* d_voice_out, size_t row_stride, float* d_bus_out,
                           bool moving) {
    const int p = (int)(b->ts_step % (uint64_t)kTsBufs);
    s2::RenderArgs a;
    a.params = b->d_params;
    a.state = b->d_state;
    a.n_voices = (uint32_t)b->n_voices;
    a.slot_begin = 0;
    a.slot_end = (uint32_t)b->n_voices;
    a.vpad = (uint32_t)b->vpad;
    a.sample_rate = (float)b->sample_rate;
    a.frames = (uint32_t)frames;
    a.voice_out = d_voice_out;
    a.row_stride = row_stride;
    a.bus_partials = nullptr;
    a.has_sine = b->n_sine ? 1u : 0u;
    a.one = 1.0f;
    a.force_path = b->force_path;
    a.staged_release = nullptr;
    a.release_row = nullptr;
    if (b->ts_main_dirty) {
        // the pre-pass reads the carried phase: order it after whatever the bank's stream wrote
        CUDA_TRY(cudaEventRecord(b->ev_main, b->stream));
        CUDA_TRY(cudaStreamWaitEvent(b->ts, b->ev_main, 0));
        b->ts_main_dirty = false;
    }
    if (b->ts_step >= (uint64_t)kTsBufs) CUDA_TRY(cudaStreamWaitEvent(b->ts, b->ev_k2[p], 0));   // seg_phase[p] is free again
    CUDA_TRY(s2::launch_ts_phase(a, b->d_seg_phase[p], b->ts));
    CUDA_TRY(cudaEventRecord(b->ev_k1[p], b->ts));
    CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_k1[p], 0));
    CUDA_TRY(s2::launch_ts_render(a, b->filter_kind, moving, b->d_seg_phase[p], b->stream));
    CUDA_TRY(cudaEventRecord(b->ev_k2[p], b->stream));
    g_launches.fetch_add(2, std::memory_order_relaxed);
    if (d_bus_out) {
        // one warp rendered one voice: the output rows (silent voices included, as zeros) are the partial sums,
        // added by the same reduction kernel
        const uint32_t nv = (uint32_t)b->n_voices;
        if (nv == 1) {
            CUDA_TRY(cudaMemcpyAsync(d_bus_out, d_voice_out, frames * sizeof(float), cudaMemcpyDeviceToDevice, b->stream));
        } else {
            const size_t need = (size_t)s2::bus_segments(nv) * frames;
            if (need > b->partials_cap) {
                if (b->d_partials) CUDA_TRY(cudaFree(b->d_partials));
                b->d_partials = nullptr; b->partials_cap = 0;
                CUDA_TRY(cudaMalloc(&b->d_partials, need * sizeof(float)));
                b->partials_cap = need;
            }
            CUDA_TRY(s2::launch_bus_reduce(d_voice_out, nv, row_stride, (uint32_t)frames, b->d_partials, d_bus_out, b->stream));
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
    }
    b->ts_step++;
    b->ts_blocks++;
    b->total_frames += frames;
    b->max_offset += frames;
    return S2_OK;
}

int bank_render_impl(s2_bank* b, size_t frames, float* d_voice_out, size_t row_stride, float* d_bus_out,
                     int trace) {
    if (!b) return fail(S2_ERR_INVALID, "null bank");
    if (frames == 0) return S2_OK;
    if (frames > 0x7FFFFFFFull) return fail(S2_ERR_INVALID, "frames %zu too large for one call", frames);
    if (d_voice_out) {
        if (((uintptr_t)d_voice_out & 15u) != 0 || (row_stride & 3u) != 0 || row_stride < frames)
            return fail(S2_ERR_INVALID, "voice_out must be 16-byte aligned with row_stride %% 4 == 0 and >= frames");
    }
    if (b->max_offset + frames > 0xFFFFFFFFull) {
        // max_offset is an upper bound that only grows with the frames rendered; voices restarted since then sit at
        // small offsets.  Look at the real ones before failing (the reference panics only when a voice's own offset
        // overflows, process.rs:36): O(V), and only when the bound trips.
        uint64_t mx = 0;
        for (size_t i = 0; i < b->n_voices; i++)
            if (b->book[i].active) mx = std::max<uint64_t>(mx, (uint64_t)b->book[i].start_offset + (b->total_frames - b->book[i].start_total));
        b->max_offset = mx;
        if (mx + frames > 0xFFFFFFFFull) return fail(S2_ERR_OVERFLOW, "frame offset overflow (process.rs:36)");
    }
    CUDA_TRY(cudaSetDevice(b->device));
    if (trace == s2::TRACE_NONE) {
        const int cls = ts_block_class(b, frames, d_voice_out, d_bus_out);
        if (cls) return bank_render_time_split(b, frames, d_voice_out, row_stride, d_bus_out, cls == 2);
    }
    b->ts_main_dirty = true;      // the general kernels below move the carried phase on the bank's stream
    if (b->n_sub > 1 && trace == s2::TRACE_NONE) return bank_render_pipelined(b, frames, d_voice_out, row_stride, d_bus_out);
    if (b->n_sub > 1) { int rc = bank_drain(b); if (rc) return rc; }

    const uint32_t n_warps = s2::render_warps((uint32_t)b->n_voices);
    float* partials = nullptr;
    if (d_bus_out) {
        if (n_warps == 1) {
            partials = d_bus_out;   // a single warp sums its voices in index order: that IS the bus
        } else {
            // per-warp partials followed by the stage-1 segment sums of the bus reduction
            const size_t need = ((size_t)n_warps + s2::bus_segments(n_warps)) * frames;
            if (need > b->partials_cap) {
                if (b->d_partials) CUDA_TRY(cudaFree(b->d_partials));
                b->d_partials = nullptr; b->partials_cap = 0;
                CUDA_TRY(cudaMalloc(&b->d_partials, need * sizeof(float)));
                b->partials_cap = need;
            }
            partials = b->d_partials;
        }
    }

    s2::RenderArgs a;
    a.params = b->d_params;
    a.state = b->d_state;
    a.n_voices = (uint32_t)b->n_voices;
    a.slot_begin = 0;
    a.slot_end = (uint32_t)b->n_voices;
    a.vpad = (uint32_t)b->vpad;
    a.sample_rate = (float)b->sample_rate;
    a.frames = (uint32_t)frames;
    a.voice_out = d_voice_out;
    a.row_stride = row_stride;
    a.bus_partials = partials;
    a.has_sine = b->n_sine ? 1u : 0u;
    a.one = 1.0f;
    a.force_path = b->force_path;
    a.staged_release = nullptr;
    a.release_row = nullptr;
    CUDA_TRY(s2::launch_render(a, b->filter_kind, trace, b->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (d_bus_out && n_warps > 1) {
        CUDA_TRY(s2::launch_bus_reduce(partials, n_warps, frames, (uint32_t)frames, partials + (size_t)n_warps * frames,
                                       d_bus_out, b->stream, b->d_bus_counters, (size_t)s2::bus_segments(n_warps) * frames));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    b->total_frames += frames;
    b->max_offset += frames;
    return S2_OK;
}

}  // namespace

extern "C" {

uint32_t s2_abi_version(void) { return S2_ABI_VERSION; }
const char* s2_last_error(void) { return g_err; }
uint64_t s2_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int s2_device_count(int* count) {
    if (!count) return fail(S2_ERR_INVALID, "null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(S2_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return S2_OK;
}

// synth.rs:208-212
float s2_note_to_pitch(uint8_t note) {
    const float n = (float)note;
    return 440.0f * powf(2.0f, (n - 69.0f) / 12.0f);
}

// synth.rs:125-152
void s2_default_voice(s2_voice_desc* d) {
    if (!d) return;
    memset(d, 0, sizeof *d);
    d->osc_kind = S2_OSC_SAW;
    d->noise_seed = 0;
    d->pitch_hz = 440.0f;
    d->osc_gain = 1.0f;
    d->noise_amt = 0.0f;
    d->lpf_freq_hz = 200.0f;
    d->damping = 1.41421356f;
    d->amp_attack_ms = 100.0f; d->amp_decay_ms = 100.0f; d->amp_sustain = 0.5f; d->amp_release_ms = 100.0f;
    d->mod_attack_ms = 0.0f; d->mod_decay_ms = 200.0f; d->mod_sustain = 0.0f; d->mod_release_ms = 0.0f;
    d->mod_env_to_osc_freq = 0.0f;
    d->mod_env_to_lpf_freq = 10.0f;
    d->frame_offset = 0;
    d->release_offset = S2_NO_RELEASE;
    d->active = 0;
}

int s2_bank_create(int device, uint32_t sample_rate, uint32_t filter_kind, size_t n_voices,
                   const s2_voice_desc* voices, void* stream, s2_bank** out) {
    if (!out) return fail(S2_ERR_INVALID, "null out");
    *out = nullptr;
    if (n_voices == 0 || !voices) return fail(S2_ERR_INVALID, "empty bank");
    if (n_voices > 0x7FFFFFE0ull) return fail(S2_ERR_INVALID, "too many voices");
    if (sample_rate == 0) return fail(S2_ERR_INVALID, "sample_rate must be > 0");
    if (filter_kind > S2_FILTER_FIRST_ORDER_HP) return fail(S2_ERR_INVALID, "filter_kind %u", filter_kind);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(S2_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(S2_ERR_NO_DEVICE, "device %d out of range (%d)", device, ndev);
    for (size_t i = 0; i < n_voices; i++) {
        int rc = validate_voice(voices[i], i, filter_kind);
        if (rc) return rc;
    }
    CUDA_TRY(cudaSetDevice(device));

    s2_bank* b = new (std::nothrow) s2_bank;
    if (!b) return fail(S2_ERR_NOMEM, "out of host memory");
    b->device = device;
    b->stream = (cudaStream_t)stream;
    b->sample_rate = sample_rate;
    b->filter_kind = filter_kind;
    b->n_voices = n_voices;
    b->vpad = (n_voices + 63) & ~(size_t)63;
    b->book.resize(n_voices);
    if (n_voices <= kTsMaxVoices) b->descs.assign(voices, voices + n_voices);
    if (const char* f = getenv("S2_FORCE_PATH")) b->force_path = (f[0] >= '1' && f[0] <= '3') ? (uint32_t)(f[0] - '0') : 0u;   // test hook

    b->voice_of_slot.resize(n_voices);
    b->slot_of_voice.resize(n_voices);
    for (size_t i = 0; i < n_voices; i++) b->voice_of_slot[i] = (uint32_t)i;
    if (n_voices > 32) {
        // (inactive last) x oscillator kind x when the amp envelope comes to rest: the 32 voices of a warp then leave
        // the envelope ramps together instead of waiting for the slowest of a random draw.  Stable: a patch sweep
        // keeps the variants of one cutoff trajectory next to each other.
        // NOT a key: whether the cutoff follows the mod envelope.  Grouping the followers halves the warps that run
        // moving-cutoff chunks, but a step is one launch per voice range and a range's launches run in order: the
        // ranges of followers (1.9x the work per frame for 200 ms) fall behind the others and finish alone on a
        // part-filled machine.  Measured on the bench bank, first 20 blocks: 6.2 ms grouped and contiguous, 7.1 ms
        // grouped and dealt over the ranges, against mixed warps (every warp pays, all ranges in step).
        struct Key { uint32_t kind; float amp_rest; };
        std::vector<Key> keys(n_voices);
        for (size_t i = 0; i < n_voices; i++) {
            const s2_voice_desc& d = voices[i];
            keys[i] = {d.active ? d.osc_kind : 4u, d.amp_attack_ms + d.amp_decay_ms};
        }
        std::stable_sort(b->voice_of_slot.begin(), b->voice_of_slot.end(), [&](uint32_t x, uint32_t y) {
            const Key& a = keys[x];
            const Key& c = keys[y];
            if (a.kind != c.kind) return a.kind < c.kind;
            return a.amp_rest < c.amp_rest;
        });
        // Instead the followers are PAIRED with voices whose cutoff rests: inside each kind, walking the sorted
        // order, a follower and a non-follower that are next in line share an aligned lane pair (2i, 2i + 1), and the
        // resting lane computes half of its partner's moving coefficients (chunk_modcut_pr).  Voices left without
        // a partner come after the pairs, in the sorted order.
        {
            auto follows = [&](uint32_t v) { return voices[v].active && voices[v].mod_env_to_lpf_freq != 0.0f; };
            std::vector<uint32_t> out, fq, nq;
            out.reserve(n_voices);
            size_t g0 = 0;
            while (g0 < n_voices) {
                size_t g1 = g0;
                while (g1 < n_voices && keys[b->voice_of_slot[g1]].kind == keys[b->voice_of_slot[g0]].kind) g1++;
                fq.clear(); nq.clear();
                size_t fi = 0, ni = 0;             // queue heads
                for (size_t i = g0; i < g1; i++) {
                    const uint32_t v = b->voice_of_slot[i];
                    (follows(v) ? fq : nq).push_back(v);
                    if (fi < fq.size() && ni < nq.size()) {
                        if (out.size() & 1u) { out.push_back(nq[ni++]); continue; }    // keep the pairs on even slots
                        out.push_back(fq[fi++]);
                        out.push_back(nq[ni++]);
                    }
                }
                // the unpaired: merge the two leftovers back into the sorted order
                std::vector<uint32_t> left(fq.begin() + fi, fq.end());
                left.insert(left.end(), nq.begin() + ni, nq.end());
                std::stable_sort(left.begin(), left.end(), [&](uint32_t x, uint32_t y) { return keys[x].amp_rest < keys[y].amp_rest; });
                out.insert(out.end(), left.begin(), left.end());
                g0 = g1;
            }
            b->voice_of_slot.swap(out);
        }
        // Deal the sorted warps (32 slots each) round-robin into 8 bins laid out one after the other: every contiguous
        // eighth / quarter / half of the slot range — the voice ranges of s2_bank_set_pipeline — then holds the same
        // mix of kinds and envelope lengths, so the ranges' streams advance together.
        const size_t full_warps = n_voices / 32;
        if (full_warps >= 16) {
            std::vector<uint32_t> dealt;
            dealt.reserve(n_voices);
            for (size_t bin = 0; bin < 8; bin++)
                for (size_t w = bin; w < full_warps; w += 8)
                    dealt.insert(dealt.end(), b->voice_of_slot.begin() + w * 32, b->voice_of_slot.begin() + (w + 1) * 32);
            dealt.insert(dealt.end(), b->voice_of_slot.begin() + full_warps * 32, b->voice_of_slot.end());
            b->voice_of_slot.swap(dealt);
        }
    }
    for (size_t s = 0; s < n_voices; s++) {
        b->slot_of_voice[b->voice_of_slot[s]] = (uint32_t)s;
        if (b->voice_of_slot[s] != s) b->identity = false;
    }

    std::vector<float> hp((size_t)s2::P_COUNT * b->vpad, 0.0f), hs((size_t)s2::S_COUNT * b->vpad, 0.0f);
    for (size_t i = 0; i < n_voices; i++) {
        const size_t slot = b->slot_of_voice[i];
        pack_voice(voices[i], (uint32_t)i, hp.data() + slot, b->vpad);
        s2_voice_state st;
        memset(&st, 0, sizeof st);
        st.frame_offset = voices[i].frame_offset;
        pack_state(st, hs.data() + slot, b->vpad);
        b->book[i] = {voices[i].frame_offset, 0, voices[i].active ? 1u : 0u, voices[i].osc_kind};
        if (voices[i].active && voices[i].frame_offset > b->max_offset) b->max_offset = voices[i].frame_offset;
        if (voices[i].osc_kind == S2_OSC_SINE) b->n_sine++;
    }
    cudaError_t e = cudaMalloc(&b->d_params, hp.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_state, hs.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_bus_counters, s2::kBusCounters * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_bus_counters, 0, s2::kBusCounters * sizeof(unsigned int), b->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_params, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_state, hs.data(), hs.size() * sizeof(float), cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    if (e != cudaSuccess) {
        int rc = fail(e == cudaErrorMemoryAllocation ? S2_ERR_NOMEM : S2_ERR_CUDA, "bank upload: %s", cudaGetErrorString(e));
        s2_bank_destroy(b);
        return rc;
    }
    *out = b;
    return S2_OK;
}

void s2_bank_destroy(s2_bank* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    bank_drain(b);
    cudaStreamSynchronize(b->stream);
    cudaFree(b->d_params);
    cudaFree(b->d_state);
    cudaFree(b->d_partials);
    cudaFree(b->d_bus_counters);
    cudaFree(b->d_bus);
    cudaFree(b->d_stage);
    for (int k = 0; k < 8; k++) {
        if (b->sub[k]) { cudaStreamSynchronize(b->sub[k]); cudaStreamDestroy(b->sub[k]); }
        if (b->ev_sub[k]) cudaEventDestroy(b->ev_sub[k]);
    }
    if (b->mix) { cudaStreamSynchronize(b->mix); cudaStreamDestroy(b->mix); }
    if (b->upload) { cudaStreamSynchronize(b->upload); cudaStreamDestroy(b->upload); }
    for (int k = 0; k < 8; k++) for (int q = 0; q < 2; q++) if (b->ev_gather[k][q]) cudaEventDestroy(b->ev_gather[k][q]);
    for (int i = 0; i < 2; i++) {
        if (b->ev_stage[i]) cudaEventDestroy(b->ev_stage[i]);
        cudaFree(b->d_stage2[i]);
    }
    for (int i = 0; i < kMixBufs; i++) {
        if (b->ev_mix[i]) cudaEventDestroy(b->ev_mix[i]);
        cudaFree(b->d_partials2[i]);
    }
    if (b->ev_tail) cudaEventDestroy(b->ev_tail);
    if (b->ts) { cudaStreamSynchronize(b->ts); cudaStreamDestroy(b->ts); }
    for (int i = 0; i < kTsBufs; i++) {
        if (b->ev_k1[i]) cudaEventDestroy(b->ev_k1[i]);
        if (b->ev_k2[i]) cudaEventDestroy(b->ev_k2[i]);
        cudaFree(b->d_seg_phase[i]);
    }
    if (b->ev_main) cudaEventDestroy(b->ev_main);
    delete b;
}

size_t s2_bank_voices(const s2_bank* b) { return b ? b->n_voices : 0; }

int s2_bank_set_voice(s2_bank* b, size_t index, const s2_voice_desc* voice) {
    if (!b || !voice) return fail(S2_ERR_INVALID, "null argument");
    if (index >= b->n_voices) return fail(S2_ERR_INVALID, "voice index %zu out of range", index);
    int rc = validate_voice(*voice, index, b->filter_kind);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(b->device));
    if ((rc = bank_drain(b)) != S2_OK) return rc;      // pipelined mode: per-voice edits are not pipelined
    float hp[s2::P_COUNT], hs[s2::S_COUNT];
    const size_t slot = b->slot_of_voice[index];   // the voice keeps its slot (a changed kind makes that warp mixed)
    pack_voice(*voice, (uint32_t)index, hp, 1);
    s2_voice_state st;
    memset(&st, 0, sizeof st);            // st::Layer::default(), synth.rs:68
    st.frame_offset = voice->frame_offset;
    pack_state(st, hs, 1);
    // one 4-byte element per SoA row
    CUDA_TRY(cudaMemcpy2DAsync(b->d_params + slot, b->vpad * sizeof(float), hp, sizeof(float), sizeof(float),
                               s2::P_COUNT, cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(cudaMemcpy2DAsync(b->d_state + slot, b->vpad * sizeof(float), hs, sizeof(float), sizeof(float),
                               s2::S_COUNT, cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));   // hp/hs live on this stack frame
    if (b->book[index].osc_kind == S2_OSC_SINE) b->n_sine--;
    if (voice->osc_kind == S2_OSC_SINE) b->n_sine++;
    b->book[index] = {voice->frame_offset, b->total_frames, voice->active ? 1u : 0u, voice->osc_kind};
    if (!b->descs.empty()) b->descs[index] = *voice;
    if (voice->active && voice->frame_offset > b->max_offset) b->max_offset = voice->frame_offset;
    return S2_OK;
}

int s2_bank_release_voice(s2_bank* b, size_t index) {
    if (!b) return fail(S2_ERR_INVALID, "null bank");
    if (index >= b->n_voices) return fail(S2_ERR_INVALID, "voice index %zu out of range", index);
    CUDA_TRY(cudaSetDevice(b->device));
    { int rc = bank_drain(b); if (rc) return rc; }
    const uint32_t rel = current_offset(b, index);   // release_frame_offset = current_frame_offset (synth.rs:75)
    CUDA_TRY(cudaMemcpyAsync(b->d_params + (size_t)s2::P_RELEASE * b->vpad + b->slot_of_voice[index], &rel,
                             sizeof rel, cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (!b->descs.empty()) b->descs[index].release_offset = rel;
    return S2_OK;
}

int s2_bank_set_releases(s2_bank* b, const uint32_t* h_release) {
    if (!b || !h_release) return fail(S2_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(b->device));
    for (size_t i = 0; i < b->descs.size(); i++) b->descs[i].release_offset = h_release[i];
    if (b->n_sub > 1) {
        // stage the table on the mix stream; every sub-bank applies it to its own slots right before its
        // next render (bank_render_pipelined), so no sub-bank waits for another
        const int q = (int)(b->table_step & 1u);
        if (b->table_pending >= 0) {
            // two tables without a render in between: the older one is simply superseded
            b->table_pending = -1;
            b->table_step--;
            return s2_bank_set_releases(b, h_release);
        }
        if (b->table_step >= 2)     // buffer q still feeds the gathers of the table staged two uploads ago
            for (int k = 0; k < b->n_sub; k++) CUDA_TRY(cudaStreamWaitEvent(b->upload, b->ev_gather[k][q], 0));
        CUDA_TRY(cudaMemcpyAsync(b->d_stage2[q], h_release, b->n_voices * sizeof(uint32_t), cudaMemcpyHostToDevice, b->upload));
        CUDA_TRY(cudaEventRecord(b->ev_stage[q], b->upload));
        b->table_pending = q;
        b->table_step++;
        return S2_OK;
    }
    uint32_t* row = reinterpret_cast<uint32_t*>(b->d_params + (size_t)s2::P_RELEASE * b->vpad);
    if (b->identity) {
        CUDA_TRY(cudaMemcpyAsync(row, h_release, b->n_voices * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
        return S2_OK;
    }
    if (!b->d_stage) CUDA_TRY(cudaMalloc(&b->d_stage, b->n_voices * sizeof(uint32_t)));
    CUDA_TRY(cudaMemcpyAsync(b->d_stage, h_release, b->n_voices * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(s2::launch_gather_u32(b->d_stage, b->d_params + (size_t)s2::P_ROW * b->vpad, row,
                                   (uint32_t)b->n_voices, b->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return S2_OK;
}

int s2_bank_render(s2_bank* b, size_t frames, float* d_voice_out, size_t row_stride, float* d_bus_out) {
    return bank_render_impl(b, frames, d_voice_out, row_stride, d_bus_out, s2::TRACE_NONE);
}

int s2_bank_trace_phase(s2_bank* b, size_t frames, float* d_phase_out, size_t row_stride) {
    if (!d_phase_out) return fail(S2_ERR_INVALID, "null phase_out");
    return bank_render_impl(b, frames, d_phase_out, row_stride, nullptr, s2::TRACE_PHASE);
}

int s2_bank_render_bus_host(s2_bank* b, size_t frames, float* d_voice_out, size_t row_stride, float* h_bus_out) {
    int rc = s2_bank_render_bus_host_async(b, frames, d_voice_out, row_stride, h_bus_out);
    if (rc) return rc;
    if (frames == 0) return S2_OK;
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return S2_OK;
}

int s2_bank_render_bus_host_async(s2_bank* b, size_t frames, float* d_voice_out, size_t row_stride, float* h_bus_out) {
    if (!b || !h_bus_out) return fail(S2_ERR_INVALID, "null argument");
    if (frames == 0) return S2_OK;
    CUDA_TRY(cudaSetDevice(b->device));
    if (frames > b->bus_cap) {
        if (b->d_bus) CUDA_TRY(cudaFree(b->d_bus));
        b->d_bus = nullptr; b->bus_cap = 0;
        CUDA_TRY(cudaMalloc(&b->d_bus, frames * sizeof(float)));
        b->bus_cap = frames;
    }
    int rc = bank_render_impl(b, frames, d_voice_out, row_stride, b->d_bus, s2::TRACE_NONE);
    if (rc) return rc;
    if (b->n_sub > 1) {
        // pipelined: the mix lives on the bank's mix stream; the caller's stream only waits for the copy
        CUDA_TRY(cudaMemcpyAsync(h_bus_out, b->d_bus, frames * sizeof(float), cudaMemcpyDeviceToHost, b->mix));
        CUDA_TRY(cudaEventRecord(b->ev_tail, b->mix));
        CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_tail, 0));
        return S2_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(h_bus_out, b->d_bus, frames * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    return S2_OK;
}

int s2_bank_get_state(s2_bank* b, s2_voice_state* out) {
    if (!b || !out) return fail(S2_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(b->device));
    { int rc = bank_drain(b); if (rc) return rc; }
    std::vector<float> hs((size_t)s2::S_COUNT * b->vpad);
    CUDA_TRY(cudaMemcpyAsync(hs.data(), b->d_state, hs.size() * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    const size_t p = b->vpad;
    for (size_t v = 0; v < b->n_voices; v++) {
        const size_t i = b->slot_of_voice[v];
        s2_voice_state& s = out[v];
        s.phase = hs[s2::S_PHASE * p + i];
        s.has_phase = fbits(hs[s2::S_HAS_PHASE * p + i]);
        s.frame_offset = fbits(hs[s2::S_OFFSET * p + i]);
        s.lpf_last = hs[s2::S_LAST * p + i];
        s.x1 = hs[s2::S_X1 * p + i];
        s.x2 = hs[s2::S_X2 * p + i];
        s.y1 = hs[s2::S_Y1 * p + i];
        s.y2 = hs[s2::S_Y2 * p + i];
    }
    return S2_OK;
}

int s2_bank_set_state(s2_bank* b, const s2_voice_state* in) {
    if (!b || !in) return fail(S2_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(b->device));
    { int rc = bank_drain(b); if (rc) return rc; }
    std::vector<float> hs((size_t)s2::S_COUNT * b->vpad, 0.0f);
    uint64_t mx = 0;
    for (size_t i = 0; i < b->n_voices; i++) {
        pack_state(in[i], hs.data() + b->slot_of_voice[i], b->vpad);
        b->book[i].start_offset = in[i].frame_offset;
        b->book[i].start_total = b->total_frames;
        if (b->book[i].active && in[i].frame_offset > mx) mx = in[i].frame_offset;
    }
    b->max_offset = mx;
    CUDA_TRY(cudaMemcpyAsync(b->d_state, hs.data(), hs.size() * sizeof(float), cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return S2_OK;
}

int s2_bank_sync(s2_bank* b) {
    if (!b) return fail(S2_ERR_INVALID, "null bank");
    CUDA_TRY(cudaSetDevice(b->device));
    { int rc = bank_drain(b); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return S2_OK;
}

int s2_bank_set_pipeline(s2_bank* b, int n_sub) {
    if (!b) return fail(S2_ERR_INVALID, "null bank");
    if (n_sub < 1 || n_sub > 8) return fail(S2_ERR_INVALID, "n_sub must be in [1, 8]");
    if (n_sub > 1 && b->ts_enabled) return fail(S2_ERR_INVALID, "time-split and pipelined voice ranges are exclusive");
    CUDA_TRY(cudaSetDevice(b->device));
    { int rc = s2_bank_sync(b); if (rc) return rc; }
    if (b->n_sub > 1 && b->table_pending >= 0) {
        // a note-off table staged for the next pipelined render: apply it now, the ranges are about to change
        CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_stage[b->table_pending], 0));
        CUDA_TRY(s2::launch_gather_u32(b->d_stage2[b->table_pending], b->d_params + (size_t)s2::P_ROW * b->vpad,
                                       reinterpret_cast<uint32_t*>(b->d_params + (size_t)s2::P_RELEASE * b->vpad),
                                       (uint32_t)b->n_voices, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
        b->table_pending = -1;
    }
    if (n_sub > 1) {
        for (int k = 0; k < n_sub; k++) {
            if (!b->sub[k]) CUDA_TRY(cudaStreamCreateWithFlags(&b->sub[k], cudaStreamNonBlocking));
            if (!b->ev_sub[k]) CUDA_TRY(cudaEventCreateWithFlags(&b->ev_sub[k], cudaEventDisableTiming));
            CUDA_TRY(cudaEventRecord(b->ev_sub[k], b->sub[k]));
        }
        if (!b->mix) {
            int prio_lo = 0, prio_hi = 0;
            CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CUDA_TRY(cudaStreamCreateWithPriority(&b->mix, cudaStreamNonBlocking, prio_hi));
        }
        if (!b->upload) CUDA_TRY(cudaStreamCreateWithFlags(&b->upload, cudaStreamNonBlocking));
        for (int k = 0; k < n_sub; k++)
            for (int q = 0; q < 2; q++)
                if (!b->ev_gather[k][q]) {
                    CUDA_TRY(cudaEventCreateWithFlags(&b->ev_gather[k][q], cudaEventDisableTiming));
                    CUDA_TRY(cudaEventRecord(b->ev_gather[k][q], b->sub[k]));
                }
        for (int i = 0; i < kMixBufs; i++)
            if (!b->ev_mix[i]) CUDA_TRY(cudaEventCreateWithFlags(&b->ev_mix[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            if (!b->ev_stage[i]) CUDA_TRY(cudaEventCreateWithFlags(&b->ev_stage[i], cudaEventDisableTiming));
            if (!b->d_stage2[i]) CUDA_TRY(cudaMalloc(&b->d_stage2[i], b->n_voices * sizeof(uint32_t)));
        }
        if (!b->ev_tail) CUDA_TRY(cudaEventCreateWithFlags(&b->ev_tail, cudaEventDisableTiming));
    }
    b->n_sub = n_sub;
    b->step = 0;
    b->table_step = 0;
    b->table_pending = -1;
    return S2_OK;
}

int s2_bank_set_time_split(s2_bank* b, int enable) {
    if (!b) return fail(S2_ERR_INVALID, "null bank");
    CUDA_TRY(cudaSetDevice(b->device));
    { int rc = s2_bank_sync(b); if (rc) return rc; }
    if (!enable) { b->ts_enabled = false; return S2_OK; }
    if (b->filter_kind > S2_FILTER_BIQUAD_LP)
        return fail(S2_ERR_INVALID, "time-split rendering knows the one-pole and the second-order low-pass filters");
    if (b->n_voices > kTsMaxVoices)
        return fail(S2_ERR_INVALID, "time-split rendering is for narrow banks (<= %zu voices)", kTsMaxVoices);
    if (b->n_sub > 1) return fail(S2_ERR_INVALID, "time-split and pipelined voice ranges are exclusive");
    if (!b->ts) CUDA_TRY(cudaStreamCreateWithFlags(&b->ts, cudaStreamNonBlocking));
    for (int i = 0; i < kTsBufs; i++) {
        if (!b->d_seg_phase[i]) CUDA_TRY(cudaMalloc(&b->d_seg_phase[i], b->n_voices * 96 * sizeof(float)));
        if (!b->ev_k1[i]) CUDA_TRY(cudaEventCreateWithFlags(&b->ev_k1[i], cudaEventDisableTiming));
        if (!b->ev_k2[i]) CUDA_TRY(cudaEventCreateWithFlags(&b->ev_k2[i], cudaEventDisableTiming));
    }
    if (!b->ev_main) CUDA_TRY(cudaEventCreateWithFlags(&b->ev_main, cudaEventDisableTiming));
    b->ts_enabled = true;
    b->ts_main_dirty = true;
    b->ts_step = 0;
    return S2_OK;
}

int s2_bank_time_split_blocks(s2_bank* b, uint64_t* blocks) {
    if (!b || !blocks) return fail(S2_ERR_INVALID, "null argument");
    *blocks = b->ts_blocks;
    return S2_OK;
}

int s2_bank_join(s2_bank* b, void* stream) {
    if (!b) return fail(S2_ERR_INVALID, "null bank");
    if (b->n_sub <= 1) return S2_OK;       // everything already is on the bank's stream
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    for (int k = 0; k < b->n_sub; k++) CUDA_TRY(cudaStreamWaitEvent(st, b->ev_sub[k], 0));
    CUDA_TRY(cudaEventRecord(b->ev_tail, b->mix));
    CUDA_TRY(cudaStreamWaitEvent(st, b->ev_tail, 0));
    return S2_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// Synth mirror (s2_lib/src/try3/synth.rs)

namespace {
constexpr int kNumVoices = 8;   // synth.rs:7

struct SynthVoice {             // synth.rs:23-30 (DSP state lives on the device)
    uint8_t note = 0;
    float velocity = 0.0f;      // stored, never read by the DSP (synth.rs:26)
    bool has_current = false;
    bool has_release = false;
    uint32_t release = 0;
};
}  // namespace

struct s2_synth {
    int device = 0;
    s2_bank* bank = nullptr;    // created by the first sample() call, which is when the rate is known
    uint32_t sample_rate = 0;
    SynthVoice voices[kNumVoices];
    // note events that arrived before the bank exists
    s2_voice_desc pending[kNumVoices];
    bool dirty[kNumVoices] = {};
    s2_patch patch;             // what note_on plays: Synth::default_config() unless s2_synth_set_patch changed it
    float* d_score = nullptr;   // device buffer of s2_synth_render_score
    size_t score_cap = 0;
};

namespace {

uint32_t synth_current(const s2_synth* s, int i) {
    if (!s->voices[i].has_current) return 0xFFFFFFFFu;   // unwrap_or(u32::MAX), synth.rs:106-107
    if (s->bank && !s->dirty[i]) return current_offset(s->bank, (size_t)i);
    return s->pending[i].frame_offset;
}

int synth_flush(s2_synth* s) {
    for (int i = 0; i < kNumVoices; i++) {
        if (!s->dirty[i]) continue;
        int rc = s2_bank_set_voice(s->bank, (size_t)i, &s->pending[i]);
        if (rc) return rc;
        s->dirty[i] = false;
    }
    return S2_OK;
}

}  // namespace

extern "C" {

int s2_synth_new(int device, s2_synth** out) {
    if (!out) return fail(S2_ERR_INVALID, "null out");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(S2_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(S2_ERR_NO_DEVICE, "device %d out of range (%d)", device, ndev);
    s2_synth* s = new (std::nothrow) s2_synth;
    if (!s) return fail(S2_ERR_NOMEM, "out of host memory");
    s->device = device;
    for (int i = 0; i < kNumVoices; i++) s2_default_voice(&s->pending[i]);   // Voice::default(): free slot
    s2_default_patch(&s->patch);
    *out = s;
    return S2_OK;
}

void s2_synth_free(s2_synth* s) {
    if (!s) return;
    s2_bank_destroy(s->bank);
    if (s->d_score) cudaFree(s->d_score);
    delete s;
}

// synth.rs:61-70 + next_voice :101-120: the first slot with the strictly greatest offset; free = u32::MAX
int s2_synth_note_on(s2_synth* s, uint8_t note, float velocity) {
    if (!s) return fail(S2_ERR_INVALID, "null synth");
    int oldest = 0;
    for (int i = 1; i < kNumVoices; i++)
        if (synth_current(s, i) > synth_current(s, oldest)) oldest = i;
    SynthVoice& v = s->voices[oldest];
    v.note = note;
    v.velocity = velocity;
    v.has_current = true;
    v.has_release = false;
    v.release = 0;
    s2_voice_desc d = s->patch.voice;
    d.pitch_hz = s2_note_to_pitch(note);
    d.frame_offset = 0;
    d.release_offset = S2_NO_RELEASE;
    d.active = 1;
    s->pending[oldest] = d;
    s->dirty[oldest] = true;
    return S2_OK;
}

// synth.rs:72-99: the LAST active voice holding that note
int s2_synth_note_off(s2_synth* s, uint8_t note) {
    if (!s) return fail(S2_ERR_INVALID, "null synth");
    int found = -1;
    for (int i = 0; i < kNumVoices; i++) {
        const SynthVoice& v = s->voices[i];
        if (v.note == note && v.has_current && !v.has_release) found = i;
    }
    if (found < 0) return S2_OK;
    SynthVoice& v = s->voices[found];
    v.has_release = true;
    v.release = synth_current(s, found);
    s->pending[found].release_offset = v.release;      // the description mirrors the bank (rebuilds start from it)
    if (s->dirty[found] || !s->bank) {
        s->dirty[found] = true;
        return S2_OK;
    }
    return s2_bank_release_voice(s->bank, (size_t)found);
}

// The bank is created by the first render, which is when the rate is known (synth.rs:156 takes it per call).
static int synth_ensure_bank(s2_synth* s, uint32_t sample_rate) {
    if (s->bank && s->sample_rate != sample_rate) {
        // Changing the rate mid-stream re-derives every rate-dependent constant; carry the DSP state over
        // into a bank built for the new rate.
        std::vector<s2_voice_state> st(kNumVoices);
        std::vector<s2_voice_desc> ds(kNumVoices);
        int rc = s2_bank_get_state(s->bank, st.data());
        if (rc) return rc;
        for (int i = 0; i < kNumVoices; i++) {
            ds[i] = s->pending[i];
            if (!s->dirty[i]) ds[i].frame_offset = st[i].frame_offset;
            // a note_off on an uploaded voice went straight to the bank: the description must carry it too
            ds[i].release_offset = s->voices[i].has_release ? s->voices[i].release : S2_NO_RELEASE;
        }
        s2_bank* nb = nullptr;
        rc = s2_bank_create(s->device, sample_rate, s->patch.filter_kind, kNumVoices, ds.data(), nullptr, &nb);
        if (rc) return rc;
        for (int i = 0; i < kNumVoices; i++)
            if (s->dirty[i]) { memset(&st[i], 0, sizeof st[i]); st[i].frame_offset = ds[i].frame_offset; }
        rc = s2_bank_set_state(nb, st.data());
        if (rc) { s2_bank_destroy(nb); return rc; }
        s2_bank_destroy(s->bank);
        s->bank = nb;
        s->sample_rate = sample_rate;
        for (int i = 0; i < kNumVoices; i++) s->dirty[i] = false;
    }
    if (!s->bank) {
        int rc = s2_bank_create(s->device, sample_rate, s->patch.filter_kind, kNumVoices, s->pending, nullptr, &s->bank);
        if (rc) return rc;
        s->sample_rate = sample_rate;
        for (int i = 0; i < kNumVoices; i++) s->dirty[i] = false;
    }
    return synth_flush(s);
}

int s2_synth_sample(s2_synth* s, float* h_buffer, size_t frames, uint32_t sample_rate) {
    if (!s || (!h_buffer && frames)) return fail(S2_ERR_INVALID, "null argument");
    if (frames == 0) return S2_OK;
    if (sample_rate == 0) return fail(S2_ERR_INVALID, "sample_rate must be > 0");
    int rc = synth_ensure_bank(s, sample_rate);
    if (rc) return rc;
    return s2_bank_render_bus_host(s->bank, frames, nullptr, 0, h_buffer);
}

int s2_synth_set_patch(s2_synth* s, const s2_patch* patch) {
    if (!s || !patch) return fail(S2_ERR_INVALID, "null argument");
    if (patch->filter_kind > S2_FILTER_FIRST_ORDER_HP) return fail(S2_ERR_INVALID, "filter_kind %u", patch->filter_kind);
    s2_voice_desc probe = patch->voice;
    probe.pitch_hz = 440.0f;                    // the template's pitch is replaced per note
    int rc = validate_voice(probe, 0, patch->filter_kind);
    if (rc) return rc;
    if (patch->filter_kind != s->patch.filter_kind) {
        // a voice sounds from its note_on until its amp envelope has ended: release start = max(release, A + D)
        // (old/simdtest.rs:283-285), end = start + R; `has_current` alone never clears (synth.rs:23-30)
        for (int i = 0; i < kNumVoices; i++) {
            const SynthVoice& v = s->voices[i];
            if (!v.has_current) continue;
            bool sounding = !v.has_release;
            if (v.has_release) {
                const s2_voice_desc& d = s->pending[i];
                const double sr = s->sample_rate ? (double)s->sample_rate : 48000.0;
                const double ad = sr * ((double)d.amp_attack_ms + (double)d.amp_decay_ms) / 1000.0;
                const double end = std::max((double)v.release, ad) + sr * (double)d.amp_release_ms / 1000.0 + 1.0;
                sounding = (double)synth_current(s, i) < end;
            }
            if (sounding) return fail(S2_ERR_INVALID, "the filter kind can only change while no voice is sounding");
        }
        s2_bank_destroy(s->bank);               // the next render builds a bank of the new kind
        s->bank = nullptr;
        for (int i = 0; i < kNumVoices; i++) { s2_default_voice(&s->pending[i]); s->dirty[i] = false; s->voices[i] = SynthVoice(); }
    }
    s->patch = *patch;
    s->patch.name[sizeof s->patch.name - 1] = 0;
    return S2_OK;
}

// main.rs:138-147: messages are applied between 16-frame chunks, i.e. before the first chunk that starts at
// or after their arrival
static uint64_t score_quantise(uint64_t frame) { return (frame + 15u) & ~(uint64_t)15u; }

int s2_synth_render_score(s2_synth* s, const s2_note_event* events, size_t n_events, uint32_t sample_rate,
                          float* h_buffer, size_t frames) {
    if (!s || (!h_buffer && frames) || (!events && n_events)) return fail(S2_ERR_INVALID, "null argument");
    if (sample_rate == 0) return fail(S2_ERR_INVALID, "sample_rate must be > 0");
    for (size_t i = 0; i < n_events; i++) {
        if (i && events[i].frame < events[i - 1].frame) return fail(S2_ERR_INVALID, "event %zu is out of time order", i);
        if (events[i].on > 1) return fail(S2_ERR_INVALID, "event %zu: on must be 0 or 1", i);
        if (events[i].frame > 0xFFFFFFFFFFFFFFF0ull) return fail(S2_ERR_INVALID, "event %zu: time out of range", i);
    }
    if (frames == 0) return S2_OK;
    CUDA_TRY(cudaSetDevice(s->device));
    if (frames > s->score_cap) {
        if (s->d_score) CUDA_TRY(cudaFree(s->d_score));
        s->d_score = nullptr; s->score_cap = 0;
        CUDA_TRY(cudaMalloc(&s->d_score, frames * sizeof(float)));
        s->score_cap = frames;
    }
    size_t pos = 0, i = 0;
    while (pos < frames) {
        while (i < n_events && score_quantise(events[i].frame) <= pos) {
            const int rc = events[i].on ? s2_synth_note_on(s, events[i].note, events[i].velocity)
                                        : s2_synth_note_off(s, events[i].note);
            if (rc < 0) return rc;
            i++;
        }
        size_t next = frames;
        if (i < n_events && score_quantise(events[i].frame) < frames) next = (size_t)score_quantise(events[i].frame);
        int rc = synth_ensure_bank(s, sample_rate);          // uploads the voices the events touched
        if (rc) return rc;
        rc = bank_render_impl(s->bank, next - pos, nullptr, 0, s->d_score + pos, s2::TRACE_NONE);
        if (rc) return rc;
        pos = next;
    }
    CUDA_TRY(cudaMemcpyAsync(h_buffer, s->d_score, frames * sizeof(float), cudaMemcpyDeviceToHost, s->bank->stream));
    CUDA_TRY(cudaStreamSynchronize(s->bank->stream));
    return S2_OK;
}

int s2_synth_voice_info(s2_synth* s, int slot, uint8_t* note, uint32_t* current_offset_out,
                        uint32_t* release_offset, s2_voice_state* state) {
    if (!s || slot < 0 || slot >= kNumVoices) return fail(S2_ERR_INVALID, "bad slot");
    const SynthVoice& v = s->voices[slot];
    if (note) *note = v.note;
    if (current_offset_out) *current_offset_out = synth_current(s, slot);
    if (release_offset) *release_offset = v.has_release ? v.release : S2_NO_RELEASE;
    if (state) {
        memset(state, 0, sizeof *state);
        if (s->bank && !s->dirty[slot]) {
            std::vector<s2_voice_state> st(kNumVoices);
            int rc = s2_bank_get_state(s->bank, st.data());
            if (rc) return rc;
            *state = st[slot];
        }
    }
    return v.has_current ? 1 : 0;
}

}  // extern "C"
