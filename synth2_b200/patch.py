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


def make_events(items) -> np.ndarray:
    """[(frame, "on" | "off", note[, velocity]), ...] -> NOTE_EVENT records (kept in the given order)."""
    ev = np.zeros(len(items), dtype=NOTE_EVENT)
    for i, it in enumerate(items):
        ev["frame"][i] = int(it[0])
        ev["on"][i] = 1 if it[1] == "on" else 0
        ev["note"][i] = int(it[2])
        ev["velocity"][i] = float(it[3]) if len(it) > 3 else 1.0
    return ev
