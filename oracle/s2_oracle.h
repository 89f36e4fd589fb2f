/*
 * s2_oracle.h — CPU restatement of the brson/synth2 render hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (synth2_b200/, libs2cuda.so) never does.
 *
 * Every function restates one reference function, same operation order, one IEEE-754 binary32
 * round-to-nearest operation per reference operation, fused multiply-add only where the
 * reference calls `mul_add` under its default `fma` feature (s2_lib/Cargo.toml:6-8).
 * Compile with -ffp-contract=off -fno-fast-math (oracle/Makefile).
 *
 * PINNING STATUS
 *   pinned by the reference's own tests (tests/test_oracle_kats.py):
 *     hash_word            hashnoise.rs:70-98  (popcount sum 3,200,064; scalar==x16)
 *     table lookups        lookup.rs:250-310   (7 known answers x scalar/x16)
 *     ADSR shape, 2^(m*a)f Untitled.ipynb cells 2,4 (f64 prototype values, weak)
 *   PARITY UNPINNED (nothing in the reference pins them; the Rust toolchain is absent, so
 *   the reference cannot be run here): oscillators, f32 envelopes, one-pole filter,
 *   process_layer_x16, Synth::sample, and the bits of libm/sleef transcendentals
 *   (`sleef::pow` 0.3.2 at process.rs:244 is restated as glibc powf(2,x); `f32::exp` /
 *   `f32::powf` are glibc expf/powf, which is what Rust's std calls on linux-gnu).
 *   The 2nd-order low-pass (dsp_filters.rs:82-130) is dead code in the reference; its use as
 *   the voice filter (filter_kind = 1) is a composition of reference formulas.
 */
#ifndef S2_ORACLE_H
#define S2_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* static_config.rs:25-31 (declaration order) */
enum { S2O_OSC_SQUARE = 0, S2O_OSC_SAW = 1, S2O_OSC_TRIANGLE = 2, S2O_OSC_SINE = 3 };
/* 0: filters.rs:15-34 (live path)   1: dsp_filters.rs:82-130 (spec for "resonant biquad") */
enum { S2O_FILTER_ONE_POLE = 0, S2O_FILTER_BIQUAD_LP = 1, S2O_FILTER_BIQUAD_HP = 2, S2O_FILTER_BIQUAD_BP = 3,
       S2O_FILTER_FIRST_ORDER_LP = 4, S2O_FILTER_FIRST_ORDER_HP = 5 };

#define S2O_NO_RELEASE 0xFFFFFFFFu /* Option::None == unwrap_or(u32::MAX): simdtest.rs:283, envelopes.rs:35 */

/* static_config.rs:38-44 */
typedef struct {
    float attack_ms, decay_ms, sustain, release_ms;
} s2o_adsr;

/* static_config.rs:3-36 + the filter extension */
typedef struct {
    uint32_t osc_kind;
    float osc_gain;
    float noise;
    float lpf_freq;
    s2o_adsr amp_env;
    s2o_adsr mod_env;
    float mod_env_to_osc_freq;
    float mod_env_to_lpf_freq;
    uint32_t filter_kind;
    float damping;
} s2o_layer_config;

/* state.rs:8-21, oscillators.rs:402-406, filters.rs:3-7, dsp_filters.rs:74-80 */
typedef struct {
    uint32_t has_phase; /* Option<Unipolar<1>> discriminant */
    float phase;
    uint32_t noise_seed;
    float lpf_last;
    float x1, x2, y1, y2;
} s2o_layer_state;

/* Same memory layout as s2_voice_desc in include/s2_cuda.h (checked by tests). */
typedef struct {
    uint32_t osc_kind;
    uint32_t noise_seed;
    float pitch_hz;
    float osc_gain;
    float noise_amt;
    float lpf_freq_hz;
    float damping;
    float amp_attack_ms, amp_decay_ms, amp_sustain, amp_release_ms;
    float mod_attack_ms, mod_decay_ms, mod_sustain, mod_release_ms;
    float mod_env_to_osc_freq, mod_env_to_lpf_freq;
    uint32_t frame_offset;   /* synth.rs:27 current_frame_offset */
    uint32_t release_offset; /* synth.rs:28, S2O_NO_RELEASE = None */
    uint32_t active;         /* current_frame_offset.is_some() */
} s2o_voice_desc;

/* Same layout as s2_voice_state in include/s2_cuda.h. */
typedef struct {
    float phase;
    uint32_t has_phase;
    uint32_t frame_offset;
    float lpf_last;
    float x1, x2, y1, y2;
} s2o_voice_state;

/* ---- L0/L1 primitives ---- */
float s2o_ms_as_samples(float ms, uint32_t sample_rate);           /* units.rs:44-53 */
float s2o_hz_as_samples(float hz, uint32_t sample_rate);           /* units.rs:19-26, 32-41 */
float s2o_note_to_pitch(uint8_t note);                             /* synth.rs:208-212 */
uint32_t s2o_hash_word(uint32_t start, uint32_t word);             /* hashnoise.rs:53-55 */
void s2o_hash_word_x16(const uint32_t* start, const uint32_t* word, uint32_t* out); /* hashnoise.rs:57-68 */
float s2o_hash_noise(uint32_t seed, float offset);                 /* hashnoise.rs:15-26, 33-51 */
float s2o_noise_fast_form(uint32_t seed, uint32_t n);              /* kernel's division-free form, for a CPU proof */
float s2o_line_fma(float rise, float run, float x, float y0);      /* math.rs:11-19, 27-40 */
float s2o_line_nofma(float rise, float run, float x, float y0);    /* old/simdtest.rs:247-261 */
float s2o_adsr_x16_lane(float attack, float decay, float sustain, float release,
                        uint32_t offset, uint32_t release_offset); /* old/simdtest.rs:270-330 */
float s2o_adsr_scalar(float attack, float decay, float sustain, float release,
                      uint32_t offset, uint32_t release_offset);   /* envelopes.rs:22-149 */
float s2o_modulate_freq(float freq, float mod_sample, float amount); /* process.rs:221-250 */
float s2o_accum_phase(float phase, float period);                  /* oscillators.rs:377-381 */
/* x16 = 1: gather_or_default semantics (lookup.rs:72-73); 0: scalar (lookup.rs:31-32 would panic) */
float s2o_table_lookup_exclusive(const float* table, uint32_t len, float value, float range, int x16);
float s2o_table_lookup_inclusive(const float* table, uint32_t len, float value, float range, int x16);
float s2o_table_lookup_periodic(const float* table, uint32_t len, float value, float range, int x16);
/* basic oscillators behind the phased layer: phase in [0,1), offset 0 (oscillators.rs:47-239) */
float s2o_osc_sample(uint32_t kind, float period, float phase, int x16);
float s2o_lpf_coeff(float freq, uint32_t sample_rate);             /* filters.rs:17-21 */
float s2o_lpf_process(float* last, uint32_t sample_rate, float freq, float input); /* filters.rs:15-34 */
/* coefficients (alpha, beta, gamma) of dsp_filters.rs:99-109 */
void s2o_biquad_lp_coeffs(uint32_t sample_rate, float cutoff, float damping, float* abg);
/* dsp_filters.rs:132-230 and :12-80 (dead code in the reference, like the low-pass above) */
float s2o_biquad_hp_process(s2o_layer_state* st, uint32_t sample_rate, float cutoff, float damping, float input);
float s2o_biquad_bp_process(s2o_layer_state* st, uint32_t sample_rate, float center, float quality, float input);
float s2o_first_order_process(s2o_layer_state* st, uint32_t sample_rate, float cutoff, int high, float input);
float s2o_biquad_lp_process(s2o_layer_state* st, uint32_t sample_rate, float cutoff, float damping,
                            float input);                          /* dsp_filters.rs:91-130 */
const float* s2o_sin_table(void);                                  /* tables.rs:1-1026 */

/* ---- L2 block render (process.rs) ---- */
void s2o_process_layer_x16(const s2o_layer_config* cfg, s2o_layer_state* st, float pitch,
                           uint32_t sample_rate, uint32_t offset, uint32_t release_offset,
                           float out[16]);                         /* process.rs:88-99, 137-174, 306-379 */
float s2o_process_layer(const s2o_layer_config* cfg, s2o_layer_state* st, float pitch,
                        uint32_t sample_rate, uint32_t offset, uint32_t release_offset); /* process.rs:75-86 */
/* returns 0, or -1 on frame-offset overflow (the reference panics: process.rs:36,71) */
int s2o_process_layer_buf_simd(const s2o_layer_config* cfg, s2o_layer_state* st, float pitch,
                               uint32_t sample_rate, uint32_t offset, uint32_t release_offset,
                               float* buf, size_t len);            /* process.rs:14-49 */

/* debug taps for bit-exact parity of discrete quantities: phase before each sample, and the
   sine-table index (0xFFFFFFFF when the oscillator is not Sine).  x16 semantics. */
void s2o_trace_voice(const s2o_layer_config* cfg, float pitch, uint32_t sample_rate,
                     uint32_t offset, uint32_t release_offset, size_t frames /* multiple of 16 */,
                     s2o_layer_state* st, float* phases, uint32_t* table_idx, float* out);

/* ---- L3 Synth (synth.rs) ---- */
typedef struct s2o_synth s2o_synth;
s2o_synth* s2o_synth_new(void);                                    /* synth.rs:54-59 */
void s2o_synth_free(s2o_synth*);
void s2o_synth_note_on(s2o_synth*, uint8_t note, float velocity);  /* synth.rs:61-70 */
int s2o_synth_note_off(s2o_synth*, uint8_t note);                  /* synth.rs:72-80; 1 = "released twice" warn */
void s2o_synth_sample(s2o_synth*, float* buffer, size_t frames, uint32_t sample_rate); /* synth.rs:154-203 */
/* test hooks: voice slot inspection */
int s2o_synth_voice_info(const s2o_synth*, int slot, uint8_t* note, uint32_t* cur, uint32_t* rel,
                         s2o_layer_state* st);
void s2o_default_config(s2o_layer_config* out);                    /* synth.rs:125-152 */

/* ---- voice bank (the batched shape of BASELINE configs 2-5) ---- */
/* Renders `frames` frames of every active voice with process_layer_buf_simd semantics, advances
   `states`, writes voice_out[v*stride + i] (may be NULL) and bus[i] = ((0+v0[i])+v1[i])+...
   over active voices in index order (synth.rs:176-202; may be NULL).  nthreads > 1 partitions
   voices statically over threads (CPU baseline); the bus is then a sum of per-thread partial
   buses (order differs from the reference; use nthreads = 1 for parity). */
int s2o_bank_render(const s2o_voice_desc* voices, s2o_voice_state* states, size_t n_voices,
                    uint32_t sample_rate, uint32_t filter_kind, size_t frames,
                    float* voice_out, size_t stride, float* bus, int nthreads);
void s2o_bank_init_states(const s2o_voice_desc* voices, s2o_voice_state* states, size_t n_voices);

#ifdef __cplusplus
}
#endif
#endif
