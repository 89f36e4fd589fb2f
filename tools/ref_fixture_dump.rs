//! ref_fixture_dump.rs — writes the reference's OWN outputs for the hot path, so that the CPU oracle of synth2-b200
//! (oracle/s2_oracle.c) can be pinned to them bit for bit.  The image that builds synth2-b200 has no Rust toolchain;
//! anyone with the reference's nightly toolchain turns "parity unpinned" into a green test with:
//!
//!   1. cp tools/ref_fixture_dump.rs  <synth2>/components/s2_lib/src/try3/ref_fixture_dump.rs
//!   2. add to <synth2>/components/s2_lib/src/try3/mod.rs:      #[cfg(test)] mod ref_fixture_dump;
//!      (an in-crate test module: `filters` and `dsp_filters` are private modules, try3/mod.rs:8-9)
//!   3. S2_REF_IN=<synth2-b200>/tests/golden/ref_in  S2_REF_OUT=<synth2-b200>/tests/golden/ref_out \
//!        cargo test -p s2_lib --release ref_fixture_dump -- --nocapture
//!   4. in synth2-b200:  python -m pytest tests/test_reference_fixtures.py
//!
//! Inputs (tests/golden/ref_in, written by tools/make_ref_inputs.py):
//!   bank16.desc   16 x 80-byte little-endian s2_voice_desc records (include/s2_cuda.h)
//!   signal.f32    1000 f32 samples fed to the dsp_filters.rs filters
//! Outputs (raw little-endian f32):
//!   synth_config1.f32      480,000 frames of Synth::sample for the scripted notes of BASELINE config 1
//!   bank16_onepole.f32     [16][1000] process_layer_buf_simd per voice (62 x16 chunks + 8 tail frames; one-pole)
//!   dsp_lp / dsp_hp / dsp_bp / dsp_fo_lp / dsp_fo_hp .f32   [16][1000] the dsp_filters.rs filters on signal.f32 with
//!                          each voice's (lpf_freq_hz, damping) held constant at 48 kHz
use super::dsp_filters::*;
use super::process::process_layer_buf_simd;
use super::state as st;
use super::static_config as sc;
use super::synth::{Note, Synth, Velocity};
use super::units::*;
use std::fs;
use std::path::PathBuf;

const SR: SampleRateKhz = SampleRateKhz(48000);
const FRAMES: usize = 1000;

fn dir(var: &str) -> PathBuf {
    PathBuf::from(std::env::var(var).unwrap_or_else(|_| panic!("set {var}")))
}

fn write_f32(name: &str, data: &[f32]) {
    let mut bytes = Vec::with_capacity(data.len() * 4);
    for x in data {
        bytes.extend_from_slice(&x.to_le_bytes());
    }
    let out = dir("S2_REF_OUT");
    fs::create_dir_all(&out).unwrap();
    fs::write(out.join(name), bytes).unwrap();
}

fn read_f32(name: &str) -> Vec<f32> {
    let bytes = fs::read(dir("S2_REF_IN").join(name)).unwrap();
    bytes.chunks_exact(4).map(|b| f32::from_le_bytes([b[0], b[1], b[2], b[3]])).collect()
}

/// One s2_voice_desc record (include/s2_cuda.h: 20 little-endian 32-bit words).
struct VoiceDesc {
    osc_kind: u32,
    noise_seed: u32,
    pitch_hz: f32,
    osc_gain: f32,
    noise_amt: f32,
    lpf_freq_hz: f32,
    damping: f32,
    amp: [f32; 4],
    mod_: [f32; 4],
    mod_to_osc: f32,
    mod_to_lpf: f32,
    frame_offset: u32,
    release_offset: u32,
    active: u32,
}

fn read_bank() -> Vec<VoiceDesc> {
    let bytes = fs::read(dir("S2_REF_IN").join("bank16.desc")).unwrap();
    bytes
        .chunks_exact(80)
        .map(|r| {
            let u = |i: usize| u32::from_le_bytes([r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]]);
            let f = |i: usize| f32::from_bits(u(i));
            VoiceDesc {
                osc_kind: u(0),
                noise_seed: u(1),
                pitch_hz: f(2),
                osc_gain: f(3),
                noise_amt: f(4),
                lpf_freq_hz: f(5),
                damping: f(6),
                amp: [f(7), f(8), f(9), f(10)],
                mod_: [f(11), f(12), f(13), f(14)],
                mod_to_osc: f(15),
                mod_to_lpf: f(16),
                frame_offset: u(17),
                release_offset: u(18),
                active: u(19),
            }
        })
        .collect()
}

fn layer_of(v: &VoiceDesc) -> sc::Layer {
    let adsr = |e: &[f32; 4]| sc::Adsr { attack: Ms(e[0]), decay: Ms(e[1]), sustain: Unipolar(e[2]), release: Ms(e[3]) };
    sc::Layer {
        osc: sc::Oscillator {
            kind: match v.osc_kind {
                0 => sc::OscillatorKind::Square,
                1 => sc::OscillatorKind::Saw,
                2 => sc::OscillatorKind::Triangle,
                _ => sc::OscillatorKind::Sine,
            },
            gain: Unipolar(v.osc_gain),
        },
        noise: Unipolar(v.noise_amt),
        lpf: sc::LowPassFilter { freq: Hz(v.lpf_freq_hz) },
        amp_env: adsr(&v.amp),
        mod_env: adsr(&v.mod_),
        modulations: sc::Modulations {
            mod_env_to_osc_freq: Bipolar(v.mod_to_osc),
            mod_env_to_lpf_freq: Bipolar(v.mod_to_lpf),
        },
    }
}

#[test]
fn ref_fixture_dump() {
    // ---- BASELINE config 1: Synth::sample over the scripted notes (tests/golden/make_golden.py: EVENTS, TOTAL)
    let events: [(usize, bool, u8); 6] =
        [(0, true, 69), (96000, true, 57), (192000, true, 76), (240000, false, 69), (336000, false, 57), (336000, false, 76)];
    let total = 480000usize;
    let mut synth = Synth::new();
    let mut buf = vec![0.0f32; total];
    let mut cuts: Vec<usize> = events.iter().map(|e| e.0).collect();
    cuts.push(0);
    cuts.push(total);
    cuts.sort();
    cuts.dedup();
    for w in cuts.windows(2) {
        for (frame, on, note) in events.iter() {
            if *frame == w[0] {
                if *on {
                    synth.note_on(Note(*note), Velocity(Unipolar(1.0)));
                } else {
                    synth.note_off(Note(*note));
                }
            }
        }
        synth.sample(&mut buf[w[0]..w[1]], SR);
    }
    write_f32("synth_config1.f32", &buf);

    // ---- the 16-voice bank through process_layer_buf_simd (one-pole low-pass: the only filter process.rs wires in)
    let bank = read_bank();
    let mut out = vec![0.0f32; bank.len() * FRAMES];
    for (i, v) in bank.iter().enumerate() {
        if v.active == 0 {
            continue; // inactive voices render silence (synth.rs:177-180 skips them)
        }
        let cfg = layer_of(v);
        let mut state = st::Layer::default();
        state.noise.seed = v.noise_seed;
        let release = if v.release_offset == u32::MAX { None } else { Some(v.release_offset) };
        process_layer_buf_simd(&cfg, &mut state, Hz(v.pitch_hz), SR, v.frame_offset, release, &mut out[i * FRAMES..(i + 1) * FRAMES]);
    }
    write_f32("bank16_onepole.f32", &out);

    // ---- dsp_filters.rs on a fixed signal, one (cutoff, damping) per voice
    let signal = read_f32("signal.f32");
    assert_eq!(signal.len(), FRAMES);
    let mut lp = vec![0.0f32; bank.len() * FRAMES];
    let mut hp = lp.clone();
    let mut bp = lp.clone();
    let mut fo_lp = lp.clone();
    let mut fo_hp = lp.clone();
    for (i, v) in bank.iter().enumerate() {
        let (f, d) = (Hz(v.lpf_freq_hz), Unipolar::<10>(v.damping));
        let mut s_lp = SecondOrderLowPassFilterState::default();
        let mut s_hp = SecondOrderHighPassFilterState::default();
        let mut s_bp = SecondOrderBandPassFilterState::default();
        let mut s_fl = FirstOrderLowPassFilterState::default();
        let mut s_fh = FirstOrderHighPassFilterState::default();
        for (n, x) in signal.iter().enumerate() {
            let k = i * FRAMES + n;
            lp[k] = SecondOrderLowPassFilter { state: &mut s_lp, sample_rate: SR, cutoff_freq: f, damping_factor: d }.process(*x);
            hp[k] = SecondOrderHighPassFilter { state: &mut s_hp, sample_rate: SR, cutoff_freq: f, damping_factor: d }.process(*x);
            // the band-pass reads the field as its quality factor; the fixture adds 2 as tests/golden/make_golden.py does
            bp[k] = SecondOrderBandPassFilter { state: &mut s_bp, sample_rate: SR, center_freq: f, quality_factor: Unipolar::<10>(v.damping + 2.0) }
                .process(*x);
            fo_lp[k] = FirstOrderLowPassFilter { state: &mut s_fl, sample_rate: SR, cutoff_freq: f }.process(*x);
            fo_hp[k] = FirstOrderHighPassFilter { state: &mut s_fh, sample_rate: SR, cutoff_freq: f }.process(*x);
        }
    }
    write_f32("dsp_lp.f32", &lp);
    write_f32("dsp_hp.f32", &hp);
    write_f32("dsp_bp.f32", &bp);
    write_f32("dsp_fo_lp.f32", &fo_lp);
    write_f32("dsp_fo_hp.f32", &fo_hp);
}
