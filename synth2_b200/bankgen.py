"""Synthetic voice banks of BASELINE.json's shapes (SURVEY.md section 8d "synthetic bank generator").

Counter-based and order-free: every field of voice v is a pure function of (v, field index), so a
rank can generate exactly its own voice range and the oracle and the GPU see identical bytes.

    r(v, k) = splitmix64(0x53594E32_00000000 ^ (v << 8) ^ k),   u(v, k) = (r >> 40) / 2^24 in [0, 1)
"""
import numpy as np

from ._lib import NO_RELEASE, OSC_SAW, OSC_SQUARE, VOICE_DESC
from .bank import note_to_pitch

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def _u(v: np.ndarray, k: int) -> np.ndarray:
    r = splitmix64(np.uint64(0x53594E3200000000) ^ (v.astype(np.uint64) << np.uint64(8)) ^ np.uint64(k))
    return (r >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def pitch_table() -> np.ndarray:
    """128-entry f32 table of `note_to_pitch` (synth.rs:208-212), computed once on the host."""
    return np.array([note_to_pitch(n) for n in range(128)], dtype=np.float32)


# mod-env -> cutoff amounts (octaves at full envelope).  The one-pole filter takes any cutoff
# (k = exp(-2*pi*f/sr) just underflows towards 0), so its banks use the reference default of 10
# (synth.rs:150).  The 2nd-order low-pass (dsp_filters.rs:99-109) is only stable below Nyquist:
# 8000 Hz * 2^1.5 = 22.6 kHz < 24 kHz, so biquad banks sweep 1.5 octaves.
MOD_TO_LPF_ONE_POLE = (0.0, 10.0)
MOD_TO_LPF_BIQUAD = (0.0, 1.5)


def make_bank(n_voices: int, render_frames: int, first_voice: int = 0, kinds=(OSC_SAW, OSC_SQUARE),
              mod_to_lpf_choices=MOD_TO_LPF_ONE_POLE, pitches: np.ndarray = None) -> np.ndarray:
    """Voices [first_voice, first_voice + n_voices) of the synthetic bank.

    note in [24, 108]; cutoff log-uniform [100, 8000] Hz; damping uniform [0.2, 1.414]; A, D, R
    uniform [5, 500] ms; S uniform [0.2, 0.9]; mod env 0/200/0/0 ms; mod->lpf drawn from `mod_to_lpf_choices`; mod->osc 0;
    osc gain 1, noise amount 0, noise seed = voice index; oscillator kind cycles through `kinds`
    by voice index; note-on at frame 0; release at 75 % of `render_frames`, rounded down to a
    multiple of 16.
    """
    if pitches is None:
        pitches = pitch_table()
    v = np.arange(first_voice, first_voice + n_voices, dtype=np.uint64)
    b = np.zeros(n_voices, dtype=VOICE_DESC)
    kinds = np.asarray(kinds, dtype=np.uint32)
    b["osc_kind"] = kinds[(v % np.uint64(len(kinds))).astype(np.int64)]
    b["noise_seed"] = (v & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    note = 24 + np.minimum((_u(v, 0) * 85.0).astype(np.int64), 84)
    b["pitch_hz"] = pitches[note]
    b["osc_gain"] = 1.0
    b["noise_amt"] = 0.0
    b["lpf_freq_hz"] = (100.0 * np.exp(_u(v, 1) * np.log(80.0))).astype(np.float32)
    b["damping"] = (0.2 + _u(v, 2) * (1.414 - 0.2)).astype(np.float32)
    b["amp_attack_ms"] = (5.0 + _u(v, 3) * 495.0).astype(np.float32)
    b["amp_decay_ms"] = (5.0 + _u(v, 4) * 495.0).astype(np.float32)
    b["amp_sustain"] = (0.2 + _u(v, 5) * 0.7).astype(np.float32)
    b["amp_release_ms"] = (5.0 + _u(v, 6) * 495.0).astype(np.float32)
    b["mod_attack_ms"] = 0.0
    b["mod_decay_ms"] = 200.0
    b["mod_sustain"] = 0.0
    b["mod_release_ms"] = 0.0
    b["mod_env_to_osc_freq"] = 0.0
    choices = np.asarray(mod_to_lpf_choices, dtype=np.float32)
    b["mod_env_to_lpf_freq"] = choices[np.minimum((_u(v, 7) * len(choices)).astype(np.int64), len(choices) - 1)]
    b["frame_offset"] = 0
    rel = (int(render_frames) * 3 // 4) & ~15
    b["release_offset"] = rel if rel > 0 else NO_RELEASE
    b["active"] = 1
    return b
