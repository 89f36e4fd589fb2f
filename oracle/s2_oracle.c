/*
 * s2_oracle.c — CPU restatement of the brson/synth2 render hot path (see s2_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, loaded by, or called from the product path.
 *
 * Paths below are relative to /root/reference/components/.  Rust semantics restated in C:
 *   f32 `%`            -> fmodf (exact)                 `x as u32` -> saturating, NaN -> 0
 *   `a.mul_add(b, c)`  -> fmaf(a, b, c)                 `n as f32` -> round-to-nearest
 *   everything else    -> one binary32 op per source op (no contraction: -ffp-contract=off)
 */
#include "s2_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static const uint32_t SIN_TABLE_BITS[1024] = {
#include "sin_table_bits.inc"
};

const float* s2o_sin_table(void) { return (const float*)SIN_TABLE_BITS; }

/* Rust `f32 as u32`: saturating, NaN -> 0 */
static uint32_t f32_as_u32(float v) {
    if (!(v > 0.0f)) return 0u; /* NaN, negatives, zero */
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

/* ------------------------------------------------------------------ units.rs */

/* s2_lib/src/try3/units.rs:44-53: seconds = ms / 1000.0; samples = sample_rate * seconds */
float s2o_ms_as_samples(float ms, uint32_t sample_rate) {
    float sr = (float)sample_rate;
    float seconds = ms / 1000.0f;
    return sr * seconds;
}

/* s2_lib/src/try3/units.rs:19-26 and 32-41 */
float s2o_hz_as_samples(float hz, uint32_t sample_rate) {
    float sr = (float)sample_rate;
    return sr / hz;
}

/* s2_lib/src/try3/synth.rs:208-212 */
float s2o_note_to_pitch(uint8_t note) {
    float n = (float)note;
    return 440.0f * powf(2.0f, (n - 69.0f) / 12.0f);
}

/* ------------------------------------------------------------------ hashnoise.rs */

/* s2_lib/src/try3/hashnoise.rs:7,53-55 */
uint32_t s2o_hash_word(uint32_t start, uint32_t word) {
    uint32_t rot = (start << 5) | (start >> 27);
    return (rot ^ word) * 0x9e3779b9u;
}

/* s2_lib/src/try3/hashnoise.rs:57-68 (explicit shifts/or instead of rotate_left) */
void s2o_hash_word_x16(const uint32_t* start, const uint32_t* word, uint32_t* out) {
    for (int i = 0; i < 16; i++) {
        uint32_t l = start[i] << 5;
        uint32_t r = start[i] >> (32 - 5);
        out[i] = ((l | r) ^ word[i]) * 0x9e3779b9u;
    }
}

/* s2_lib/src/try3/hashnoise.rs:15-26 (scalar) == :33-51 (x16): value / 65535 * 2 - 1 */
float s2o_hash_noise(uint32_t seed, float offset) {
    uint32_t off = f32_as_u32(offset);
    uint32_t hash = s2o_hash_word(seed, off);
    float value = (float)(uint16_t)hash;
    float q = value / 65535.0f;
    float d = q * 2.0f;
    return d - 1.0f;
}

/* NOT a reference function: the division-free form of the noise map used by the CUDA kernel
   (synth2_b200/csrc/s2_device.cuh noise_fast).  Kept here so a CPU-only test can prove it equal to
   s2o_hash_noise for all 65,536 possible 16-bit hash values. */
float s2o_noise_fast_form(uint32_t seed, uint32_t n) {
    uint32_t hash = s2o_hash_word(seed, n);
    float v = (float)(hash & 0xffffu);
    /* v * 0x1.0001p-16 is exact inside the fma and lies less than 2^-32 below v / 65535 with no rounding boundary
       in between unless it sits on one (a tie): the tiny addend resolves those upwards, as the quotient would be */
    float q = fmaf(v, 0x1.0001p-16f, 0x1p-45f);
    return fmaf(q, 2.0f, -1.0f);
}

/* ------------------------------------------------------------------ math.rs */

/* s2_lib/src/try3/math.rs:11-19 / 27-40 with feature "fma": slope = rise / run; fma(slope, x, y0) */
float s2o_line_fma(float rise, float run, float x, float y0) {
    float slope = rise / run;
    return fmaf(slope, x, y0);
}

/* s2_lib/src/old/simdtest.rs:247-261: (rise / run) * x, then + y0 (never fused) */
float s2o_line_nofma(float rise, float run, float x, float y0) {
    float slope = rise / run;
    float y = slope * x;
    return y + y0;
}

/* ------------------------------------------------------------------ envelopes */

/* s2_lib/src/old/simdtest.rs:270-330, one lane.  Release starts no earlier than attack+decay
   (:285) and always from the sustain level. */
float s2o_adsr_x16_lane(float attack, float decay, float sustain, float release,
                        uint32_t offset_u, uint32_t release_offset_u) {
    float offset = (float)offset_u;
    float decay_offset = attack;
    float sustain_offset = attack + decay;
    float release_offset = (float)release_offset_u; /* None -> u32::MAX as f32 */
    release_offset = fmaxf(release_offset, sustain_offset); /* simd_max, simdtest.rs:285 */
    float end_offset = release_offset + release;

    int in_attack = offset < decay_offset;
    int in_decay = !in_attack && offset < sustain_offset;
    int in_sustain = !in_attack && !in_decay && offset < release_offset;
    int in_release = !in_attack && !in_decay && !in_sustain && offset < end_offset;
    int in_end = !in_attack && !in_decay && !in_sustain && !in_release;

    float attack_sample = s2o_line_nofma(1.0f, attack, offset, 0.0f);
    float decay_sample = s2o_line_nofma(sustain - 1.0f, decay, offset - decay_offset, 1.0f);
    float sustain_sample = sustain;
    float release_sample = s2o_line_nofma(-sustain, release, offset - release_offset, sustain);
    float end_sample = 0.0f;

    float sample = 0.0f;
    if (in_attack) sample = attack_sample;
    if (in_decay) sample = decay_sample;
    if (in_sustain) sample = sustain_sample;
    if (in_release) sample = release_sample;
    if (in_end) sample = end_sample;
    return sample;
}

/* s2_lib/src/try3/envelopes.rs:22-149 (tail frames only).  Release takes precedence and starts
   from the level reached at release time; FMA line (math.rs:11-19). */
float s2o_adsr_scalar(float attack, float decay, float sustain, float release,
                      uint32_t offset_u, uint32_t release_offset_u) {
    float offset = (float)offset_u;
    float decay_offset = attack;
    float sustain_offset = attack + decay;
    float release_offset = (float)release_offset_u;
    float end_offset = release_offset + release;

    int in_release = offset >= release_offset && offset < end_offset;
    int in_end = offset >= end_offset;
    int in_attack = !in_release && !in_end && offset < decay_offset;
    int in_decay = !in_release && !in_end && !in_attack && offset < sustain_offset;
    int in_sustain = !in_release && !in_end && !in_attack && !in_decay && offset < release_offset;

    float release_start_sample;
    if (release_offset < decay_offset)
        release_start_sample = s2o_line_fma(1.0f, attack, release_offset, 0.0f);
    else if (release_offset < sustain_offset)
        release_start_sample = s2o_line_fma(sustain - 1.0f, decay, release_offset - decay_offset, 1.0f);
    else
        release_start_sample = sustain;

    if (in_attack) return s2o_line_fma(1.0f, attack, offset, 0.0f);
    if (in_decay) return s2o_line_fma(sustain - 1.0f, decay, offset - decay_offset, 1.0f);
    if (in_sustain) return sustain;
    if (in_release)
        return s2o_line_fma(-release_start_sample, release, offset - release_offset, release_start_sample);
    return 0.0f;
}

/* s2_lib/src/try3/process.rs:221-229 (libm powf) and :231-250 (sleef pow, restated as powf:
   "parity unpinned", s2_oracle.h) */
float s2o_modulate_freq(float freq, float mod_sample, float amount) {
    float m = mod_sample * amount;
    return powf(2.0f, m) * freq;
}

/* ------------------------------------------------------------------ lookup.rs */

static float gather(const float* table, uint32_t len, uint32_t idx) {
    return idx < len ? table[idx] : 0.0f; /* gather_or_default, lookup.rs:72-73 */
}

/* s2_lib/src/try3/lookup.rs:10-44 (scalar) / 46-85 (x16) */
float s2o_table_lookup_exclusive(const float* table, uint32_t len, float value, float range, int x16) {
    (void)x16; /* out-of-range index: x16 gathers 0.0; the scalar path would panic (lookup.rs:31-32) */
    float table_length = (float)len;
    float table_value = value * table_length / range;
    uint32_t low = f32_as_u32(table_value);
    uint32_t idx1 = low;
    uint32_t idx2 = (idx1 + 1u) % len;
    float lowf = (float)low;
    float s1 = gather(table, len, idx1);
    float s2 = gather(table, len, idx2);
    return s2o_line_fma(s2 - s1, 1.0f, table_value - lowf, s1);
}

/* s2_lib/src/try3/lookup.rs:92-130 / 132-174 */
float s2o_table_lookup_inclusive(const float* table, uint32_t len, float value, float range, int x16) {
    (void)x16;
    float table_length = (float)(len > 0 ? len - 1u : 0u);
    float table_value = value * table_length / range;
    uint32_t low = f32_as_u32(table_value);
    uint32_t idx1 = low;
    uint32_t idx2 = (idx1 + 1u) % len;
    float lowf = (float)low;
    float s1 = gather(table, len, idx1);
    float s2 = gather(table, len, idx2);
    return s2o_line_fma(s2 - s1, 1.0f, table_value - lowf, s1);
}

/* s2_lib/src/try3/lookup.rs:181-199 */
float s2o_table_lookup_periodic(const float* table, uint32_t len, float value, float range, int x16) {
    return s2o_table_lookup_exclusive(table, len, fmodf(value, range), range, x16);
}

/* ------------------------------------------------------------------ oscillators.rs */

/* s2_lib/src/try3/oscillators.rs:377-381 */
float s2o_accum_phase(float phase, float period) {
    float delta = 1.0f / period;
    return fmodf(phase + delta, 1.0f);
}

static __thread uint32_t g_last_table_idx; /* debug tap for s2o_trace_voice */

/* phased layer (oscillators.rs:207-239: fma(period, phase, offset = 0)) then the basic
   oscillators (:47-199).  Scalar and x16 variants perform the same operations. */
float s2o_osc_sample(uint32_t kind, float period, float phase, int x16) {
    float offset = fmaf(period, phase, 0.0f);
    switch (kind) {
    case S2O_OSC_SQUARE: { /* oscillators.rs:47-80 */
        float x = fmodf(offset, period);
        float half = period / 2.0f;
        return x < half ? 1.0f : -1.0f;
    }
    case S2O_OSC_SAW: { /* oscillators.rs:82-119 */
        float x = fmodf(offset, period);
        return s2o_line_fma(-2.0f, period, x, 1.0f);
    }
    case S2O_OSC_TRIANGLE: { /* oscillators.rs:121-183 */
        float x = fmodf(offset, period);
        float half = period / 2.0f;
        if (x < half) return s2o_line_fma(-2.0f, half, x, 1.0f);
        return s2o_line_fma(2.0f, half, x - half, -1.0f);
    }
    default: { /* Sine: oscillators.rs:185-199, 592-623 -> lookup.rs:181-199 */
        float v = fmodf(offset, period);
        float tv = v * 1024.0f / period;
        g_last_table_idx = f32_as_u32(tv);
        return s2o_table_lookup_exclusive(s2o_sin_table(), 1024u, v, period, x16);
    }
    }
}

/* ------------------------------------------------------------------ filters */

/* s2_lib/src/try3/filters.rs:17-21: (-2.0 * pi * freq / sample_rate).exp() */
float s2o_lpf_coeff(float freq, uint32_t sample_rate) {
    float sr = (float)sample_rate;
    float pi = 3.14159274101257324219f; /* std::f32::consts::PI */
    float t = -2.0f * pi;
    t = t * freq;
    t = t / sr;
    return expf(t);
}

/* s2_lib/src/try3/filters.rs:15-34, feature "fma": a0.mul_add(input, -b1 * last), b1 = -x */
float s2o_lpf_process(float* last, uint32_t sample_rate, float freq, float input) {
    float x = s2o_lpf_coeff(freq, sample_rate);
    float a0 = 1.0f - x;
    float b1 = -x;
    float fb = -b1 * *last;
    float out = fmaf(a0, input, fb);
    *last = out;
    return out;
}

/* s2_lib/src/try3/dsp_filters.rs:99-109 */
void s2o_biquad_lp_coeffs(uint32_t sample_rate, float cutoff, float damping, float* abg) {
    float sr = (float)sample_rate;
    float pi = 3.14159274101257324219f;
    float theta = 2.0f * pi;
    theta = theta * cutoff;
    theta = theta / sr;
    float s = sinf(theta);
    float c = cosf(theta);
    float hd = damping / 2.0f;
    float num = 1.0f - hd * s;
    float den = 1.0f + hd * s;
    float beta = (1.0f / 2.0f) * (num / den);
    float gamma = (1.0f / 2.0f + beta) * c;
    float alpha = (1.0f / 2.0f + beta - gamma) / 4.0f;
    abg[0] = alpha;
    abg[1] = beta;
    abg[2] = gamma;
}

/* s2_lib/src/try3/dsp_filters.rs:91-130 (no mul_add in the source -> no fused ops) */
float s2o_biquad_lp_process(s2o_layer_state* st, uint32_t sample_rate, float cutoff, float damping,
                            float input) {
    float abg[3];
    s2o_biquad_lp_coeffs(sample_rate, cutoff, damping, abg);
    float alpha = abg[0], beta = abg[1], gamma = abg[2];
    float x1 = st->x1, x2 = st->x2, y1 = st->y1, y2 = st->y2;
    float x = input;
    float s = x + 2.0f * x1;
    s = s + x2;
    float t = alpha * s;
    t = t + gamma * y1;
    t = t - beta * y2;
    float y = 2.0f * t;
    st->x2 = x1;
    st->x1 = x;
    st->y2 = y1;
    st->y1 = y;
    return y;
}

/* s2_lib/src/try3/dsp_filters.rs:132-178 SecondOrderHighPassFilter::process */
float s2o_biquad_hp_process(s2o_layer_state* st, uint32_t sample_rate, float cutoff, float damping,
                            float input) {
    float sr = (float)sample_rate;
    float pi = 3.14159274101257324219f;
    float theta = 2.0f * pi;
    theta = theta * cutoff;
    theta = theta / sr;
    float sn = sinf(theta);
    float cs = cosf(theta);
    float hd = damping / 2.0f;
    float num = 1.0f - hd * sn;
    float den = 1.0f + hd * sn;
    float beta = (1.0f / 2.0f) * (num / den);
    float gamma = (1.0f / 2.0f + beta) * cs;
    float alpha = (1.0f / 2.0f + beta + gamma) / 4.0f;
    float x1 = st->x1, x2 = st->x2, y1 = st->y1, y2 = st->y2;
    float x = input;
    float s = x - 2.0f * x1;
    s = s + x2;
    float t = alpha * s;
    t = t + gamma * y1;
    t = t - beta * y2;
    float y = 2.0f * t;
    st->x2 = x1;
    st->x1 = x;
    st->y2 = y1;
    st->y1 = y;
    return y;
}

/* s2_lib/src/try3/dsp_filters.rs:180-230 SecondOrderBandPassFilter::process; `quality` = quality_factor */
float s2o_biquad_bp_process(s2o_layer_state* st, uint32_t sample_rate, float center, float quality,
                            float input) {
    float sr = (float)sample_rate;
    float pi = 3.14159274101257324219f;
    float theta = 2.0f * pi;
    theta = theta * center;
    theta = theta / sr;
    float tn = tanf(theta / (2.0f * quality));
    float beta = (1.0f / 2.0f) * ((1.0f - tn) / (1.0f + tn));
    float gamma = (1.0f / 2.0f + beta) * cosf(theta);
    float alpha = (1.0f / 2.0f - beta) / 2.0f;
    float x2 = st->x2, y1 = st->y1, y2 = st->y2, x1 = st->x1;
    float x = input;
    float t = alpha * (x - x2);
    t = t + gamma * y1;
    t = t - beta * y2;
    float y = 2.0f * t;
    st->x2 = x1;
    st->x1 = x;
    st->y2 = y1;
    st->y1 = y;
    return y;
}

/* s2_lib/src/try3/dsp_filters.rs:12-44 FirstOrderLowPassFilter::process (high = 0) and :46-80
   FirstOrderHighPassFilter::process (high = 1) */
float s2o_first_order_process(s2o_layer_state* st, uint32_t sample_rate, float cutoff, int high, float input) {
    float sr = (float)sample_rate;
    float pi = 3.14159274101257324219f;
    float theta = 2.0f * pi;
    theta = theta * cutoff;
    theta = theta / sr;
    float gamma = cosf(theta) / (1.0f + sinf(theta));
    float alpha = high ? (1.0f + gamma) / 2.0f : (1.0f - gamma) / 2.0f;
    float x1 = st->x1, y1 = st->y1;
    float x = input;
    float xs = high ? x - x1 : x + x1;
    float y = alpha * xs + gamma * y1;
    st->x1 = x;
    st->y1 = y;
    return y;
}

static float filter_process(const s2o_layer_config* cfg, s2o_layer_state* st, uint32_t sr, float freq,
                            float input) {
    switch (cfg->filter_kind) {
    case S2O_FILTER_BIQUAD_LP: return s2o_biquad_lp_process(st, sr, freq, cfg->damping, input);
    case S2O_FILTER_BIQUAD_HP: return s2o_biquad_hp_process(st, sr, freq, cfg->damping, input);
    case S2O_FILTER_BIQUAD_BP: return s2o_biquad_bp_process(st, sr, freq, cfg->damping, input);
    case S2O_FILTER_FIRST_ORDER_LP: return s2o_first_order_process(st, sr, freq, 0, input);
    case S2O_FILTER_FIRST_ORDER_HP: return s2o_first_order_process(st, sr, freq, 1, input);
    default: return s2o_lpf_process(&st->lpf_last, sr, freq, input);
    }
}

/* ------------------------------------------------------------------ process.rs */

static void block_x16(const s2o_layer_config* cfg, s2o_layer_state* st, float pitch, uint32_t sr,
                      uint32_t offset, uint32_t release_offset, float out[16], float* phases_out,
                      uint32_t* idx_out) {
    /* prepare_frame_x16, process.rs:137-174 */
    float gains[16], periods[16], lpf_freqs[16];
    float aA = s2o_ms_as_samples(cfg->amp_env.attack_ms, sr); /* process.rs:197-202 */
    float aD = s2o_ms_as_samples(cfg->amp_env.decay_ms, sr);
    float aR = s2o_ms_as_samples(cfg->amp_env.release_ms, sr);
    float mA = s2o_ms_as_samples(cfg->mod_env.attack_ms, sr);
    float mD = s2o_ms_as_samples(cfg->mod_env.decay_ms, sr);
    float mR = s2o_ms_as_samples(cfg->mod_env.release_ms, sr);
    for (int i = 0; i < 16; i++) {
        uint32_t off = offset + (uint32_t)i; /* offsets_x16, process.rs:213-219 (wrapping lanes) */
        gains[i] = s2o_adsr_x16_lane(aA, aD, cfg->amp_env.sustain, aR, off, release_offset);
        float m = s2o_adsr_x16_lane(mA, mD, cfg->mod_env.sustain, mR, off, release_offset);
        float fo = s2o_modulate_freq(pitch, m, cfg->mod_env_to_osc_freq);
        lpf_freqs[i] = s2o_modulate_freq(cfg->lpf_freq, m, cfg->mod_env_to_lpf_freq);
        periods[i] = s2o_hz_as_samples(fo, sr);
    }

    /* sample_voice_x16, process.rs:306-379 */
    /* accum_phase_x16, oscillators.rs:391-400; initial phase = osc.phase = 0.0 (process.rs:316) */
    float phase[16];
    float acc = st->has_phase ? st->phase : 0.0f;
    phase[0] = acc;
    for (int i = 1; i < 16; i++) {
        acc = s2o_accum_phase(acc, periods[i - 1]);
        phase[i] = acc;
    }
    acc = s2o_accum_phase(acc, periods[15]);

    float samples[16];
    for (int i = 0; i < 16; i++) {
        g_last_table_idx = 0xFFFFFFFFu;
        float osc = s2o_osc_sample(cfg->osc_kind, periods[i], phase[i], 1);
        if (phases_out) phases_out[i] = phase[i];
        if (idx_out) idx_out[i] = g_last_table_idx;
        float osc_g = osc + cfg->osc_gain; /* process.rs:341-345: ADD (x16 quirk) */
        uint32_t off = offset + (uint32_t)i;
        float nz = s2o_hash_noise(st->noise_seed, (float)off); /* process.rs:347-351 */
        float nz_g = nz + cfg->noise;                          /* process.rs:353-356: ADD */
        samples[i] = osc_g + nz_g;                             /* process.rs:358 */
    }
    st->has_phase = 1;
    st->phase = acc;

    for (int i = 0; i < 16; i++) /* process.rs:363-371, sequential */
        samples[i] = filter_process(cfg, st, sr, lpf_freqs[i], samples[i]);
    for (int i = 0; i < 16; i++) /* process.rs:373-378 */
        out[i] = samples[i] * gains[i];
}

void s2o_process_layer_x16(const s2o_layer_config* cfg, s2o_layer_state* st, float pitch,
                           uint32_t sample_rate, uint32_t offset, uint32_t release_offset,
                           float out[16]) {
    block_x16(cfg, st, pitch, sample_rate, offset, release_offset, out, NULL, NULL);
}

/* process.rs:75-86 -> prepare_frame :101-135 + sample_voice :252-304 */
float s2o_process_layer(const s2o_layer_config* cfg, s2o_layer_state* st, float pitch,
                        uint32_t sr, uint32_t offset, uint32_t release_offset) {
    /* sample_envelope, process.rs:176-189 */
    float amp = s2o_adsr_scalar(s2o_ms_as_samples(cfg->amp_env.attack_ms, sr),
                                s2o_ms_as_samples(cfg->amp_env.decay_ms, sr), cfg->amp_env.sustain,
                                s2o_ms_as_samples(cfg->amp_env.release_ms, sr), offset, release_offset);
    float m = s2o_adsr_scalar(s2o_ms_as_samples(cfg->mod_env.attack_ms, sr),
                              s2o_ms_as_samples(cfg->mod_env.decay_ms, sr), cfg->mod_env.sustain,
                              s2o_ms_as_samples(cfg->mod_env.release_ms, sr), offset, release_offset);
    float fo = s2o_modulate_freq(pitch, m, cfg->mod_env_to_osc_freq);
    float fl = s2o_modulate_freq(cfg->lpf_freq, m, cfg->mod_env_to_lpf_freq);
    float period = s2o_hz_as_samples(fo, sr);

    /* phase_accumulating::*Oscillator::sample, oscillators.rs:414-431 etc. */
    float phase = st->has_phase ? st->phase : 0.0f;
    float osc = s2o_osc_sample(cfg->osc_kind, period, phase, 0);
    st->has_phase = 1;
    st->phase = s2o_accum_phase(phase, period);

    float osc_g = osc * cfg->osc_gain;                               /* process.rs:287: MUL */
    float nz = s2o_hash_noise(st->noise_seed, (float)offset);        /* process.rs:289-291 */
    float nz_g = nz * cfg->noise;                                    /* process.rs:292: MUL */
    float sample = osc_g + nz_g;
    sample = filter_process(cfg, st, sr, fl, sample);                /* process.rs:296-301 */
    return sample * amp;                                             /* process.rs:302 */
}

/* process.rs:14-49 (+ :51-73) */
int s2o_process_layer_buf_simd(const s2o_layer_config* cfg, s2o_layer_state* st, float pitch,
                               uint32_t sr, uint32_t offset, uint32_t release_offset, float* buf,
                               size_t len) {
    size_t i = 0;
    for (; i + 16 <= len; i += 16) {
        block_x16(cfg, st, pitch, sr, offset, release_offset, buf + i, NULL, NULL);
        if (offset > 0xFFFFFFFFu - 16u) return -1; /* checked_add(16).expect("overflow") */
        offset += 16u;
    }
    for (; i < len; i++) {
        buf[i] = s2o_process_layer(cfg, st, pitch, sr, offset, release_offset);
        if (offset == 0xFFFFFFFFu) return -1;
        offset += 1u;
    }
    return 0;
}

void s2o_trace_voice(const s2o_layer_config* cfg, float pitch, uint32_t sr, uint32_t offset,
                     uint32_t release_offset, size_t frames, s2o_layer_state* st, float* phases,
                     uint32_t* table_idx, float* out) {
    float tmp[16];
    for (size_t i = 0; i + 16 <= frames; i += 16) {
        block_x16(cfg, st, pitch, sr, offset, release_offset, out ? out + i : tmp,
                  phases ? phases + i : NULL, table_idx ? table_idx + i : NULL);
        offset += 16u;
    }
}

/* ------------------------------------------------------------------ synth.rs */

#define NUM_VOICES 8 /* synth.rs:7 */

typedef struct {
    uint8_t note;
    float velocity;        /* stored, never read by the DSP (synth.rs:26) */
    int has_current;       /* Option<FrameOffset> */
    uint32_t current;
    int has_release;
    uint32_t release;
    s2o_layer_state state;
} voice_t;

struct s2o_synth {
    s2o_layer_config config;
    voice_t voices[NUM_VOICES];
};

/* synth.rs:125-152 */
void s2o_default_config(s2o_layer_config* c) {
    memset(c, 0, sizeof *c);
    c->osc_kind = S2O_OSC_SAW;
    c->osc_gain = 1.0f;
    c->noise = 0.0f;
    c->lpf_freq = 200.0f;
    c->amp_env = (s2o_adsr){100.0f, 100.0f, 0.5f, 100.0f};
    c->mod_env = (s2o_adsr){0.0f, 200.0f, 0.0f, 0.0f};
    c->mod_env_to_osc_freq = 0.0f;
    c->mod_env_to_lpf_freq = 10.0f;
    c->filter_kind = S2O_FILTER_ONE_POLE;
    c->damping = 1.41421356f;
}

s2o_synth* s2o_synth_new(void) {
    s2o_synth* s = (s2o_synth*)calloc(1, sizeof *s);
    if (!s) return NULL;
    s2o_default_config(&s->config);
    return s;
}

void s2o_synth_free(s2o_synth* s) { free(s); }

/* synth.rs:101-120: first voice with the strictly greatest offset; a free voice counts as u32::MAX */
static voice_t* next_voice(s2o_synth* s) {
    int oldest = 0;
    for (int i = 1; i < NUM_VOICES; i++) {
        uint32_t this_off = s->voices[i].has_current ? s->voices[i].current : 0xFFFFFFFFu;
        uint32_t old_off = s->voices[oldest].has_current ? s->voices[oldest].current : 0xFFFFFFFFu;
        if (this_off > old_off) oldest = i;
    }
    return &s->voices[oldest];
}

/* synth.rs:61-70 */
void s2o_synth_note_on(s2o_synth* s, uint8_t note, float velocity) {
    voice_t* v = next_voice(s);
    memset(v, 0, sizeof *v);
    v->note = note;
    v->velocity = velocity;
    v->has_current = 1;
    v->current = 0;
}

/* synth.rs:72-99: the LAST active voice with that note */
int s2o_synth_note_off(s2o_synth* s, uint8_t note) {
    int found = -1;
    for (int i = 0; i < NUM_VOICES; i++) {
        voice_t* v = &s->voices[i];
        if (v->note == note && v->has_current && !v->has_release) found = i;
    }
    if (found < 0) return 0;
    voice_t* v = &s->voices[found];
    if (!v->has_release) { /* always true for an active voice; kept for shape (synth.rs:74-78) */
        v->has_release = v->has_current;
        v->release = v->current;
        return 0;
    }
    return 1;
}

/* synth.rs:171-203 */
static void accumulate_frames(s2o_synth* s, float* buffer, size_t needed, uint32_t sr) {
    float accum[16];
    for (int i = 0; i < 16; i++) accum[i] = 0.0f;
    for (int vi = 0; vi < NUM_VOICES; vi++) {
        voice_t* v = &s->voices[vi];
        if (!v->has_current) continue;
        float pitch = s2o_note_to_pitch(v->note);
        float buf[16];
        for (int i = 0; i < 16; i++) buf[i] = 0.0f;
        s2o_process_layer_buf_simd(&s->config, &v->state, pitch, sr, v->current,
                                   v->has_release ? v->release : S2O_NO_RELEASE, buf, needed);
        for (int i = 0; i < 16; i++) accum[i] = accum[i] + buf[i];
        uint64_t nxt = (uint64_t)v->current + needed; /* saturating_add, synth.rs:197 */
        v->current = nxt > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)nxt;
    }
    memcpy(buffer, accum, needed * sizeof(float));
}

/* synth.rs:154-169 */
void s2o_synth_sample(s2o_synth* s, float* buffer, size_t frames, uint32_t sr) {
    size_t i = 0;
    for (; i + 16 <= frames; i += 16) accumulate_frames(s, buffer + i, 16, sr);
    if (frames - i > 0) accumulate_frames(s, buffer + i, frames - i, sr);
}

int s2o_synth_voice_info(const s2o_synth* s, int slot, uint8_t* note, uint32_t* cur, uint32_t* rel,
                         s2o_layer_state* st) {
    if (slot < 0 || slot >= NUM_VOICES) return -1;
    const voice_t* v = &s->voices[slot];
    if (note) *note = v->note;
    if (cur) *cur = v->has_current ? v->current : S2O_NO_RELEASE;
    if (rel) *rel = v->has_release ? v->release : S2O_NO_RELEASE;
    if (st) *st = v->state;
    return v->has_current;
}

/* ------------------------------------------------------------------ voice bank */

void s2o_bank_init_states(const s2o_voice_desc* voices, s2o_voice_state* states, size_t n) {
    for (size_t v = 0; v < n; v++) {
        memset(&states[v], 0, sizeof states[v]);
        states[v].frame_offset = voices[v].frame_offset;
    }
}

typedef struct {
    const s2o_voice_desc* voices;
    s2o_voice_state* states;
    size_t v0, v1;
    uint32_t sr, filter_kind;
    size_t frames;
    float* voice_out;
    size_t stride;
    float* bus; /* per-thread partial (or the real bus when single-threaded) */
    float* scratch;
    int rc;
} bank_job;

static void* bank_worker(void* arg) {
    bank_job* j = (bank_job*)arg;
    if (j->bus) memset(j->bus, 0, j->frames * sizeof(float));
    for (size_t v = j->v0; v < j->v1; v++) {
        const s2o_voice_desc* d = &j->voices[v];
        float* row = j->voice_out ? j->voice_out + v * j->stride : j->scratch;
        if (!d->active) {
            if (j->voice_out) memset(row, 0, j->frames * sizeof(float));
            continue;
        }
        s2o_layer_config cfg;
        cfg.osc_kind = d->osc_kind;
        cfg.osc_gain = d->osc_gain;
        cfg.noise = d->noise_amt;
        cfg.lpf_freq = d->lpf_freq_hz;
        cfg.amp_env = (s2o_adsr){d->amp_attack_ms, d->amp_decay_ms, d->amp_sustain, d->amp_release_ms};
        cfg.mod_env = (s2o_adsr){d->mod_attack_ms, d->mod_decay_ms, d->mod_sustain, d->mod_release_ms};
        cfg.mod_env_to_osc_freq = d->mod_env_to_osc_freq;
        cfg.mod_env_to_lpf_freq = d->mod_env_to_lpf_freq;
        cfg.filter_kind = j->filter_kind;
        cfg.damping = d->damping;
        s2o_voice_state* s = &j->states[v];
        s2o_layer_state st;
        st.has_phase = s->has_phase;
        st.phase = s->phase;
        st.noise_seed = d->noise_seed;
        st.lpf_last = s->lpf_last;
        st.x1 = s->x1; st.x2 = s->x2; st.y1 = s->y1; st.y2 = s->y2;
        if (s2o_process_layer_buf_simd(&cfg, &st, d->pitch_hz, j->sr, s->frame_offset,
                                       d->release_offset, row, j->frames) != 0)
            j->rc = -1;
        s->has_phase = st.has_phase;
        s->phase = st.phase;
        s->lpf_last = st.lpf_last;
        s->x1 = st.x1; s->x2 = st.x2; s->y1 = st.y1; s->y2 = st.y2;
        uint64_t nxt = (uint64_t)s->frame_offset + j->frames; /* synth.rs:197 */
        s->frame_offset = nxt > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)nxt;
        if (j->bus)
            for (size_t i = 0; i < j->frames; i++) j->bus[i] = j->bus[i] + row[i]; /* synth.rs:194-195 */
    }
    return NULL;
}

int s2o_bank_render(const s2o_voice_desc* voices, s2o_voice_state* states, size_t n_voices,
                    uint32_t sample_rate, uint32_t filter_kind, size_t frames, float* voice_out,
                    size_t stride, float* bus, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n_voices) nthreads = n_voices ? (int)n_voices : 1;
    bank_job* jobs = (bank_job*)calloc((size_t)nthreads, sizeof *jobs);
    pthread_t* tids = (pthread_t*)calloc((size_t)nthreads, sizeof *tids);
    int rc = 0;
    for (int t = 0; t < nthreads; t++) {
        bank_job* j = &jobs[t];
        j->voices = voices; j->states = states;
        j->v0 = n_voices * (size_t)t / (size_t)nthreads;
        j->v1 = n_voices * (size_t)(t + 1) / (size_t)nthreads;
        j->sr = sample_rate; j->filter_kind = filter_kind; j->frames = frames;
        j->voice_out = voice_out; j->stride = stride;
        j->scratch = voice_out ? NULL : (float*)malloc((frames ? frames : 1) * sizeof(float));
        j->bus = !bus ? NULL : (t == 0 ? bus : (float*)malloc((frames ? frames : 1) * sizeof(float)));
    }
    if (nthreads == 1) {
        bank_worker(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; t++) pthread_create(&tids[t], NULL, bank_worker, &jobs[t]);
        for (int t = 0; t < nthreads; t++) pthread_join(tids[t], NULL);
    }
    for (int t = 0; t < nthreads; t++) {
        if (jobs[t].rc) rc = jobs[t].rc;
        if (bus && t > 0) {
            for (size_t i = 0; i < frames; i++) bus[i] = bus[i] + jobs[t].bus[i];
            free(jobs[t].bus);
        }
        free(jobs[t].scratch);
    }
    free(jobs);
    free(tids);
    return rc;
}
