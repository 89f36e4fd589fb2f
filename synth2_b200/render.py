"""Offline render of a `.synth2` patch: `python -m synth2_b200.render example.synth2 --seconds 10 --rate 48000 -o out.f32`.

BASELINE.json's first configuration names this entry ("render example.synth2 offline via s2_bin on CPU to a
10 s 48 kHz buffer"); the reference itself only plays live (s2_bin/src/main.rs:15-19 has the `midi` and
`build-tables` commands).  The render is the s2_bin loop — note events applied between 16-frame chunks, then
`Synth::sample` (main.rs:138-147) — run through `s2_synth_render_score` on the GPU.

Notes come from the file's `score { }` block; without one, `--note N` is held from frame 0 for `--hold`
seconds (default: 3/4 of the render).  Output: raw little-endian f32 (`.f32`, anything else) or a mono
IEEE-float WAV (`.wav`).
"""
import argparse
import struct
import sys

import numpy as np

from . import patch as patchmod
from .synth import Synth


def write_wav_f32(path, samples: np.ndarray, rate: int):
    data = np.ascontiguousarray(samples, dtype="<f4").tobytes()
    fmt = struct.pack("<HHIIHH", 3, 1, rate, rate * 4, 4, 32)             # WAVE_FORMAT_IEEE_FLOAT, mono
    fact = struct.pack("<I", samples.size)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact \
        + b"data" + struct.pack("<I", len(data)) + data
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def render_patch(p: patchmod.Patch, frames: int, rate: int, events: np.ndarray = None, device: int = 0) -> np.ndarray:
    synth = Synth(device)
    try:
        synth.set_patch(p)
        return synth.render_score(p.events if events is None else events, frames, rate)
    finally:
        synth.close()


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m synth2_b200.render", description=__doc__.split("\n\n")[0])
    ap.add_argument("patch", help=".synth2 file")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--rate", type=int, default=48000)
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("--note", type=int, default=69, help="MIDI note played when the file has no score block")
    ap.add_argument("--hold", type=float, default=None, help="seconds before that note's note_off")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)

    p = patchmod.load(args.patch, args.rate)
    frames = int(round(args.seconds * args.rate))
    events = p.events
    if events.size == 0:
        hold = args.seconds * 0.75 if args.hold is None else args.hold
        events = patchmod.make_events([(0, "on", args.note), (int(round(hold * args.rate)), "off", args.note)])
    out = render_patch(p, frames, args.rate, events, args.device)
    if args.output.lower().endswith(".wav"):
        write_wav_f32(args.output, out, args.rate)
    else:
        out.astype("<f4").tofile(args.output)
    peak = float(np.max(np.abs(out))) if out.size else 0.0
    print(f"{p.name}: {frames} frames at {args.rate} Hz, {events.size} events, peak {peak:.4f} -> {args.output}",
          file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
