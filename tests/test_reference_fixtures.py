"""Pins the CPU oracle to the reference's OWN outputs, when someone has produced them.

The reference is nightly Rust and cannot be built in this image, so the oracle is pinned to the reference's unit
tests and notebook values only (tests/test_oracle_kats.py) and DESIGN.md says "parity unpinned" for the DSP chain.
tools/ref_fixture_dump.rs is an in-crate test for s2_lib that renders the inputs of tests/golden/ref_in with the
reference itself and writes raw f32 files; drop them into tests/golden/ref_out and these tests compare the oracle
with them.  Without the files the comparison tests skip — and say so.

What is checked here without the reference: that the committed inputs are the ones the golden fixtures use, and
that the oracle renders them (so the loader and the oracle side of the comparison cannot rot).
"""
import pathlib

import numpy as np
import pytest

import oracle
from synth2_b200 import VOICE_DESC

HERE = pathlib.Path(__file__).resolve().parent
REF_IN = HERE / "golden" / "ref_in"
REF_OUT = HERE / "golden" / "ref_out"
SR, FRAMES = 48000, 1000
EVENTS = [(0, "on", 69), (96000, "on", 57), (192000, "on", 76), (240000, "off", 69), (336000, "off", 57), (336000, "off", 76)]
TOTAL = 480000


def load_bank():
    return np.frombuffer((REF_IN / "bank16.desc").read_bytes(), dtype=VOICE_DESC).copy()


def load_signal():
    return np.frombuffer((REF_IN / "signal.f32").read_bytes(), dtype="<f4").copy()


def ref(name, shape):
    p = REF_OUT / name
    if not p.exists():
        pytest.skip(f"{p.relative_to(HERE.parent)} not present: produce it with tools/ref_fixture_dump.rs (needs the "
                    "reference's nightly Rust toolchain); until then the DSP chain is 'parity unpinned'")
    return np.frombuffer(p.read_bytes(), dtype="<f4").reshape(shape)


def oracle_synth_config1():
    syn = oracle.OracleSynth()
    buf = np.zeros(TOTAL, dtype=np.float32)
    cuts = sorted({0, TOTAL, *[f for f, _, _ in EVENTS]})
    for a, b in zip(cuts[:-1], cuts[1:]):
        for f, op, note in EVENTS:
            if f == a:
                syn.note_on(note, 1.0) if op == "on" else syn.note_off(note)
        syn.sample(buf[a:b], SR)
    return buf


def oracle_dsp(kind, bank, signal):
    """The oracle's restatement of one dsp_filters.rs filter on `signal`, one (cutoff, damping) per voice."""
    L = oracle.lib()
    out = np.zeros((bank.shape[0], signal.size), dtype=np.float32)
    for i, v in enumerate(bank):
        st = np.zeros(1, dtype=oracle.LAYER_STATE)
        p = st.ctypes.data
        f, d = float(v["lpf_freq_hz"]), float(v["damping"])
        for n, x in enumerate(signal):
            x = float(x)
            if kind == "lp":
                out[i, n] = L.s2o_biquad_lp_process(p, SR, f, d, x)
            elif kind == "hp":
                out[i, n] = L.s2o_biquad_hp_process(p, SR, f, d, x)
            elif kind == "bp":
                out[i, n] = L.s2o_biquad_bp_process(p, SR, f, d + 2.0, x)
            elif kind == "fo_lp":
                out[i, n] = L.s2o_first_order_process(p, SR, f, 0, x)
            else:
                out[i, n] = L.s2o_first_order_process(p, SR, f, 1, x)
    return out


def compare(name, got, want, bit_exact):
    assert got.shape == want.shape
    same = float(np.mean(got.view(np.uint32) == want.view(np.uint32)))
    err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))))
    print(f"{name}: {100 * same:.3f} % of samples bit-identical, max |err| {err:.3e}")
    if bit_exact:
        assert same == 1.0, f"{name}: oracle and reference differ (max |err| {err:.3e})"
    else:
        # sleef `pow` (x16 path) against libm `powf`: last-bit differences of the modulated cutoff while the mod
        # envelope moves; everything downstream of them is compared with the north-star tolerance
        assert err <= 1e-4 * max(1.0, float(np.max(np.abs(want))))


# ---- always: the inputs are the golden fixtures' and the oracle renders them --------------------------------

def test_inputs_are_the_golden_bank():
    z = np.load(HERE / "golden" / "bank_small_onepole.npz")
    bank = load_bank()
    assert bank.tobytes() == z["voices"].tobytes()
    st = oracle.bank_init_states(bank)
    out, _ = oracle.bank_render(bank, st, SR, 0, FRAMES)
    assert out.tobytes() == z["out"].tobytes()
    sig = load_signal()
    assert sig.shape == (FRAMES,) and np.all(np.isfinite(sig)) and 0.0 <= sig.min() and sig.max() <= 2.0


def test_oracle_renders_every_reference_item():
    bank, sig = load_bank(), load_signal()
    for kind in ("lp", "hp", "bp", "fo_lp", "fo_hp"):
        out = oracle_dsp(kind, bank[:2], sig[:64])
        assert np.all(np.isfinite(out))


# ---- with tests/golden/ref_out: the pin ------------------------------------------------------------------

def test_synth_config1_against_the_reference():
    want = ref("synth_config1.f32", (TOTAL,))
    compare("synth_config1", oracle_synth_config1(), want, bit_exact=False)


def test_process_layer_against_the_reference():
    want = ref("bank16_onepole.f32", (16, FRAMES))
    bank = load_bank()
    st = oracle.bank_init_states(bank)
    got, _ = oracle.bank_render(bank, st, SR, 0, FRAMES)
    compare("bank16_onepole", got, want, bit_exact=False)
    still = bank["mod_env_to_lpf_freq"] == 0                       # no pow on the path: bit for bit
    if np.any(still):
        compare("bank16_onepole (cutoff not modulated)", got[still], want[still], bit_exact=True)


@pytest.mark.parametrize("kind", ["lp", "hp", "bp", "fo_lp", "fo_hp"])
def test_dsp_filters_against_the_reference(kind):
    want = ref(f"dsp_{kind}.f32", (16, FRAMES))
    got = oracle_dsp(kind, load_bank(), load_signal())
    # sinf / cosf / tanf are libm's on both sides (Rust's f32::sin lowers to the platform libm): bit for bit
    compare(f"dsp_{kind}", got, want, bit_exact=True)
