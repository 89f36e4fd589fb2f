// s2_internal.h — shared between the kernels (s2_kernels.cu) and the C-ABI host layer (s2_capi.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace s2 {

// Device-resident voice parameters: struct-of-arrays, row k at params[k * vpad + v], so the 32
// lanes of a warp (= 32 consecutive voices) load each field with one coalesced request.
// u32 fields are stored as raw bits in the float array.
enum ParamIdx {
    P_KIND = 0, P_SEED, P_PITCH, P_GAIN, P_NOISE, P_LPF, P_DAMP,
    P_AA, P_AD, P_AS, P_AR,   // amp ADSR (ms, ms, level, ms)
    P_MA, P_MD, P_MS, P_MR,   // mod ADSR
    P_AMT_OSC, P_AMT_LPF,
    P_RELEASE, P_ACTIVE,
    P_ROW,                    // caller-visible voice index of this slot (output row, see s2_capi.cu "slots")
    P_COUNT
};

// Carried state, same layout: state[k * vpad + v].
enum StateIdx {
    S_PHASE = 0, S_HAS_PHASE, S_OFFSET, S_LAST, S_X1, S_X2, S_Y1, S_Y2,
    // Memo of the constants derived from the oscillator frequency and the filter cutoff (divisions,
    // exp / sin / cos in binary64), keyed by the exact input bits: a launch that finds its key reuses
    // them instead of re-deriving ~1000 instructions per voice.  Not part of s2_voice_state.
    S_FO_KEY, S_OSC_P, S_OSC_D, S_OSC_SLOPE, S_OSC_HALF, S_OSC_TS1, S_OSC_TS2,
    S_FL_KEY, S_DAMP_KEY, S_FC_C0, S_FC_C1, S_FC_C2,
    S_COUNT
};
constexpr uint32_t kNoKey = 0x7fc00001u;   // a NaN payload no frequency can have

enum TraceMode { TRACE_NONE = 0, TRACE_PHASE = 1 };

struct RenderArgs {
    const float* params;   // [P_COUNT][vpad]
    float* state;          // [S_COUNT][vpad]
    uint32_t n_voices;     // voices in the whole bank
    uint32_t slot_begin;   // this launch renders slots [slot_begin, slot_end): a sub-bank (multiple of 64) or the bank
    uint32_t slot_end;
    uint32_t vpad;
    float sample_rate;     // `sample_rate.0 as f32` (filters.rs:17, units.rs:21)
    uint32_t frames;       // frames to render this launch
    float* voice_out;      // [n_voices][row_stride] or nullptr
    size_t row_stride;
    float* bus_partials;   // [n_warps][frames] or nullptr
    uint32_t has_sine;     // any voice uses the table oscillator -> stage SIN_TABLE in smem
    float one;             // 1.0f, opaque to the compiler (s2_cutoff.h: vaddp)
    // A staged note-off table to apply first (pipelined banks): release offset of slot s = staged_release[caller's
    // voice index of s], also written to release_row[s] so that it persists.  nullptr = the parameter row as it is.
    const uint32_t* staged_release;
    uint32_t* release_row;
    uint32_t force_path;   // test hook (S2_FORCE_PATH): 0 = normal, 1 = moving-cutoff chunks one frame at a time,
                           // 2 = every chunk through the general per-frame path, 3 = packed moving-cutoff chunks
                           // without the lane-pair help (chunk_modcut_pk only)
};

constexpr int kWarpsPerBlock = 1;
constexpr int kChunk = 32;          // frames per compute chunk (classification granularity)
// The render kernel's tile is 32 voices x 64 frames = two chunks: a warp then writes 256 contiguous bytes per output
// row per visit.  128 bytes per row per visit tops out at 5.1 TB/s of write bandwidth on this part whatever the
// kernel does in between, 256 bytes at 5.9 TB/s (tools/ubench/write_bw.cu, profiles/r2_write_bw.txt).
constexpr int kTileFrames = 64;
constexpr int kTileStride = 68;     // floats per tile row: 16-B aligned rows, conflict-free STS.128/LDS.128
constexpr int kTsTileStride = 36;   // the time-split kernel's 32-frame tile

// Launchers (s2_kernels.cu).  Return the cudaError_t of the launch.  A warp renders 32 consecutive slots.
cudaError_t launch_render(const RenderArgs& a, uint32_t filter_kind, int trace, cudaStream_t stream);
uint32_t render_warps(uint32_t n_voices);
// time-split rendering of narrow banks (s2_kernel_ts.cu): the phase pre-pass writes, for each of the block's 32
// time segments, the phase of its first frame and of the two frames before it to seg_phase[n_voices][3][32];
// the render consumes it.
// frames must be a multiple of 1024.
cudaError_t launch_ts_phase(const RenderArgs& a, float* seg_phase, cudaStream_t stream);
// moving: some voice's cutoff follows a ramping mod envelope in this block (per-frame coefficients)
cudaError_t launch_ts_render(const RenderArgs& a, uint32_t filter_kind, bool moving, const float* seg_phase,
                             cudaStream_t stream);
// one kernel, fixed summation order; partial row w starts at partials + w * row_stride.  With `counters` (kBusCounters
// zeroed words owned by the bank, launches on one stream at a time) and 4 * frames floats of scratch, wide banks
// take the four-frames-per-lane form.
constexpr uint32_t kBusCounters = 1024;
cudaError_t launch_bus_reduce(const float* partials, uint32_t n_warps, size_t row_stride, uint32_t frames,
                              float* scratch, float* bus, cudaStream_t stream, unsigned int* counters = nullptr,
                              size_t scratch_floats = 0);
uint32_t bus_segments(uint32_t n_warps);
// release_row[slot] = staged[voice_of_slot[slot]]  (bulk note-off table given in voice order)
cudaError_t launch_gather_u32(const uint32_t* staged, const float* row_index_bits, uint32_t* dst_row,
                              uint32_t n, cudaStream_t stream);

}  // namespace s2
