"""`.synth2` patch files: `synth NAME { ... }` plus an optional `score { ... }` of note events.

The reader is in the library (csrc/s2_patch.cpp, C ABI `s2_patch_parse`); the grammar and the field names —
those of `static_config::Layer` (s2_lib/src/try3/static_config.rs:3-44) — are documented there.  The
reference's own `example.synth2` (an empty block) parses to `Synth::default_config()`.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import NOTE_EVENT, PATCH, check, lib, ptr


@dataclass
class Patch:
    record: np.ndarray            # one PATCH record (voice template, filter kind, name)
    events: np.ndarray            # NOTE_EVENT records of the score block (may be empty)

    @property
    def name(self) -> str:
        return self.record["name"][0].decode("utf-8", "replace")

    @property
    def voice(self) -> np.ndarray:
        return self.record["voice"][0]

    @property
    def filter_kind(self) -> int:
        return int(self.record["filter_kind"][0])


def default_patch() -> Patch:
    rec = np.zeros(1, dtype=PATCH)
    lib().s2_default_patch(ptr(rec))
    return Patch(rec, np.zeros(0, dtype=NOTE_EVENT))


def parse(text: str, sample_rate: int = 48000) -> Patch:
    """Parses patch text; times written in s / ms become frames at `sample_rate`.  Raises S2Error with the
    line number on a malformed file."""
    data = text.encode("utf-8")
    rec = np.zeros(1, dtype=PATCH)
    n = C.c_size_t(0)
    check(lib().s2_patch_parse(data, int(sample_rate), ptr(rec), None, 0, C.byref(n)))
    events = np.zeros(n.value, dtype=NOTE_EVENT)
    if n.value:
        check(lib().s2_patch_parse(data, int(sample_rate), ptr(rec), ptr(events), n.value, C.byref(n)))
    return Patch(rec, events)


def load(path, sample_rate: int = 48000) -> Patch:
    with open(path, "r", encoding="utf-8") as f:
        return parse(f.read(), sample_rate)


_OSC_NAMES = {0: "square", 1: "saw", 2: "triangle", 3: "sine"}
_FILTER_NAMES = {0: "one_pole", 1: "biquad", 2: "biquad_hp", 3: "biquad_bp", 4: "first_order", 5: "first_order_hp"}


def dumps(p: Patch) -> str:
    """The patch (and its score, in frames) as `.synth2` text that `parse` reads back to the same records."""
    v = p.voice

    def num(x):
        return repr(float(np.float32(x)))          # shortest text that reads back to the same binary32

    lines = [f"synth {p.name or 'patch'} {{",
             f"    osc {{ kind {_OSC_NAMES[int(v['osc_kind'])]}; gain {num(v['osc_gain'])} }}",
             f"    noise {num(v['noise_amt'])}",
             f"    lpf {{ freq {num(v['lpf_freq_hz'])}; kind {_FILTER_NAMES[p.filter_kind]}; damping {num(v['damping'])} }}"]
    for env in ("amp", "mod"):
        lines.append(f"    {env}_env {{ attack {num(v[env + '_attack_ms'])}; decay {num(v[env + '_decay_ms'])}; "
                     f"sustain {num(v[env + '_sustain'])}; release {num(v[env + '_release_ms'])} }}")
    lines.append(f"    modulations {{ mod_env_to_osc_freq {num(v['mod_env_to_osc_freq'])}; "
                 f"mod_env_to_lpf_freq {num(v['mod_env_to_lpf_freq'])} }}")
    lines.append("}")
    if p.events.size:
        lines.append("score {")
        for e in p.events:
            lines.append(f"    on {int(e['frame'])} {int(e['note'])} {num(e['velocity'])}" if e["on"]
                         else f"    off {int(e['frame'])} {int(e['note'])}")
        lines.append("}")
    return "\n".join(lines) + "\n"


def make_events(items) -> np.ndarray:
    """[(frame, "on" | "off", note[, velocity]), ...] -> NOTE_EVENT records (kept in the given order)."""
    ev = np.zeros(len(items), dtype=NOTE_EVENT)
    for i, it in enumerate(items):
        ev["frame"][i] = int(it[0])
        ev["on"][i] = 1 if it[1] == "on" else 0
        ev["note"][i] = int(it[2])
        ev["velocity"][i] = float(it[3]) if len(it) > 3 else 1.0
    return ev
