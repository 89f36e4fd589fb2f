#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/s2_oracle.c).

The reference itself cannot be run here (Rust toolchain absent), so these are
*derived* fixtures: they freeze the oracle's output — itself pinned to the reference's known answers by
tests/test_oracle_kats.py — so that (a) an accidental change of the oracle is caught on the CPU, and
(b) the CUDA path is also compared with bytes that were committed before it ran.

    python tests/golden/make_golden.py
"""
import pathlib
import sys
import zlib

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import oracle  # noqa: E402
from synth2_b200 import bankgen  # noqa: E402

SR = 48000
EVENTS = [(0, "on", 69), (96000, "on", 57), (192000, "on", 76), (240000, "off", 69), (336000, "off", 57), (336000, "off", 76)]
TOTAL = 480000
# filter kinds: one-pole, then dsp_filters.rs low-pass / high-pass / band-pass / first-order low- and high-pass
BANK_FIXTURES = ((0, "bank_small_onepole"), (1, "bank_small_biquad"), (2, "bank_small_biquad_hp"),
                 (3, "bank_small_biquad_bp"), (4, "bank_small_first_lp"), (5, "bank_small_first_hp"))
WINDOWS = [0, 4800, 9600, 96000, 192000, 240000, 244800, 336000, 340800, 479936]   # 64-frame windows


def render_config1():
    syn = oracle.OracleSynth()
    buf = np.zeros(TOTAL, dtype=np.float32)
    cuts = sorted({0, TOTAL, *[f for f, _, _ in EVENTS]})
    for a, b in zip(cuts[:-1], cuts[1:]):
        for f, op, note in EVENTS:
            if f == a:
                syn.note_on(note, 1.0) if op == "on" else syn.note_off(note)
        syn.sample(buf[a:b], SR)
    return buf


def main():
    buf = render_config1()
    np.savez_compressed(
        HERE / "synth_config1.npz",
        windows=np.array(WINDOWS), window_data=np.stack([buf[w:w + 64] for w in WINDOWS]),
        stride_samples=buf[::4801].copy(), crc32=np.uint32(zlib.crc32(buf.tobytes())),
        sum=np.float64(buf.astype(np.float64).sum()), sumsq=np.float64((buf.astype(np.float64) ** 2).sum()),
        peak=np.float32(np.abs(buf).max()))
    for fk, name in BANK_FIXTURES:
        sweep = bankgen.MOD_TO_LPF_BIQUAD if fk else bankgen.MOD_TO_LPF_ONE_POLE
        v = bankgen.make_bank(16, 2048, kinds=(0, 1, 2, 3), mod_to_lpf_choices=sweep)
        if fk == 3:
            v["damping"] += 2.0        # the band-pass reads the field as its quality factor
        v["noise_amt"] = (np.arange(16) % 3) * 0.25
        v["osc_gain"] = 0.5 + (np.arange(16) % 4) * 0.125
        v["release_offset"] = 640
        v["active"][5] = 0
        st = oracle.bank_init_states(v)
        out, bus = oracle.bank_render(v, st, SR, fk, 1000)      # 62 x16 blocks + 8 tail frames
        np.savez_compressed(HERE / f"{name}.npz", voices=v, out=out, bus=bus, state=st, filter_kind=np.uint32(fk))
    print("wrote", sorted(p.name for p in HERE.glob("*.npz")))


if __name__ == "__main__":
    main()
