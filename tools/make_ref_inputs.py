#!/usr/bin/env python3
"""Writes tests/golden/ref_in/: the inputs tools/ref_fixture_dump.rs (run inside the reference crate) renders.

    bank16.desc   the 16-voice one-pole bank of tests/golden/make_golden.py as raw s2_voice_desc records
    signal.f32    1000 samples: a 110 Hz saw plus its +1 offset (what the x16 path feeds its filter), as f32

    PYTHONPATH=. python tools/make_ref_inputs.py
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle  # noqa: E402
from synth2_b200 import bankgen  # noqa: E402


def bank16():
    """Identical to the one-pole fixture of tests/golden/make_golden.py (pitches from the oracle's table, so this
    script does not need the CUDA library)."""
    v = bankgen.make_bank(16, 2048, kinds=(0, 1, 2, 3), mod_to_lpf_choices=bankgen.MOD_TO_LPF_ONE_POLE,
                          pitches=oracle.pitch_table())
    v["noise_amt"] = (np.arange(16) % 3) * 0.25
    v["osc_gain"] = 0.5 + (np.arange(16) % 4) * 0.125
    v["release_offset"] = 640
    v["active"][5] = 0
    return v


def signal():
    n = np.arange(1000, dtype=np.float64)
    period = 48000.0 / 110.0
    saw = 1.0 - 2.0 * ((n / period) % 1.0)
    return (saw + 1.0).astype(np.float32)


def main():
    out = ROOT / "tests" / "golden" / "ref_in"
    out.mkdir(parents=True, exist_ok=True)
    v = bank16()
    assert v.dtype.itemsize == 80
    (out / "bank16.desc").write_bytes(v.tobytes())
    (out / "signal.f32").write_bytes(signal().astype("<f4").tobytes())
    print("wrote", sorted(p.name for p in out.iterdir()))


if __name__ == "__main__":
    main()
