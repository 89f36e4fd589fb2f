"""Synth — host-side mirror of `s2_lib::try3::synth::Synth` (s2_lib/src/try3/synth.rs:9-203).

Same names, argument meaning and behaviour as the reference: eight voice slots, the hard-coded
default patch, oldest-voice allocation, `note_on` / `note_off` / `sample`.  `sample` overwrites a
host buffer with the mono mix rendered on the GPU; there is no CPU rendering path.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import VOICE_STATE, check, lib, ptr

NUM_VOICES = 8  # synth.rs:7


@dataclass(frozen=True)
class Note:          # synth.rs:14-16
    value: int


@dataclass(frozen=True)
class Velocity:      # synth.rs:17-18 (Unipolar<1>)
    value: float


@dataclass(frozen=True, order=True)
class FrameOffset:   # synth.rs:19-21
    value: int


class Synth:
    def __init__(self, device: int = 0):          # Synth::new(), synth.rs:54-59
        self._h = C.c_void_p()
        check(lib().s2_synth_new(int(device), C.byref(self._h)))

    @classmethod
    def new(cls, device: int = 0):
        return cls(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().s2_synth_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def note_on(self, note, velocity=Velocity(1.0)):      # synth.rs:61-70
        n = note.value if isinstance(note, Note) else int(note)
        v = velocity.value if isinstance(velocity, Velocity) else float(velocity)
        check(lib().s2_synth_note_on(self._h, n & 0xFF, v))

    def note_off(self, note) -> bool:                     # synth.rs:72-80; True = "released twice"
        n = note.value if isinstance(note, Note) else int(note)
        return check(lib().s2_synth_note_off(self._h, n & 0xFF)) == 1

    def sample(self, buffer: np.ndarray, sample_rate: int):   # synth.rs:154-169
        """Overwrites `buffer` (float32, contiguous, host) with `len(buffer)` mono frames."""
        assert isinstance(buffer, np.ndarray) and buffer.dtype == np.float32 and buffer.flags.c_contiguous
        check(lib().s2_synth_sample(self._h, ptr(buffer), buffer.size, int(sample_rate)))

    # -- beyond the reference's API: a patch other than default_config(), and the
    #    s2_bin loop (apply the MIDI messages that arrived, then sample 16 frames: main.rs:138-147) as one call
    def set_patch(self, patch):
        """Notes started from now on play `patch` (a synth2_b200.patch.Patch)."""
        check(lib().s2_synth_set_patch(self._h, ptr(patch.record)))

    def render_score(self, events: np.ndarray, frames: int, sample_rate: int, out: np.ndarray = None) -> np.ndarray:
        """Renders `frames` mono frames, applying each NOTE_EVENT before the first 16-frame chunk that starts
        at or after its arrival frame."""
        from ._lib import NOTE_EVENT
        events = np.ascontiguousarray(events, dtype=NOTE_EVENT)
        if out is None:
            out = np.empty(int(frames), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size >= frames
        check(lib().s2_synth_render_score(self._h, ptr(events) if events.size else None, events.size,
                                          int(sample_rate), ptr(out), int(frames)))
        return out[:frames]

    def voice_info(self, slot: int):
        """Test hook: (in_use, note, current_frame_offset|None, release_frame_offset|None, state)."""
        note, cur, rel = C.c_uint8(), C.c_uint32(), C.c_uint32()
        st = np.zeros(1, dtype=VOICE_STATE)
        used = check(lib().s2_synth_voice_info(self._h, int(slot), C.byref(note), C.byref(cur), C.byref(rel), ptr(st)))
        return (bool(used), note.value, cur.value if used else None,
                None if rel.value == 0xFFFFFFFF else rel.value, st[0])
