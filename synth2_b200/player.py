"""Player — the renderer-to-audio-callback hand-off of `s2_bin` (s2_bin/src/audio_player.rs, main.rs:120-160).

Two 2048-frame mono buffers circulate between an internal synth thread (GPU render + copy to pinned host
memory) and the caller's audio callback.  `fill` is the reference's `fill_buffer`: it never blocks, duplicates
the mono signal over the output channels and pads with zeros when no rendered buffer is ready.
"""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

BUFFER_FRAMES = 2048            # audio_player.rs:21


class Player:
    def __init__(self, sample_rate: int = 48000, device: int = 0, patch=None, start: bool = True):
        self._h = C.c_void_p()
        check(lib().s2_player_new(int(device), int(sample_rate), C.byref(self._h)))
        if patch is not None:
            check(lib().s2_player_set_patch(self._h, ptr(patch.record)))
        if start:
            self.start()

    def start(self):
        check(lib().s2_player_start(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().s2_player_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def note_on(self, note: int, velocity: float = 1.0):
        check(lib().s2_player_note_on(self._h, int(note) & 0xFF, float(velocity)))

    def note_off(self, note: int):
        check(lib().s2_player_note_off(self._h, int(note) & 0xFF))

    def fill(self, out: np.ndarray) -> int:
        """The audio callback: `out` is float32 [frames, channels] (or [frames]); returns the frames that
        came from rendered buffers, the rest of `out` is zeros."""
        assert out.dtype == np.float32 and out.flags.c_contiguous
        frames = out.shape[0]
        channels = out.shape[1] if out.ndim == 2 else 1
        return int(check(lib().s2_player_fill(self._h, ptr(out), frames, channels)))

    def wait_buffers(self, n_rendered: int, timeout_ms: int = 5000) -> bool:
        return check(lib().s2_player_wait_buffers(self._h, int(n_rendered), int(timeout_ms))) == 0

    def stats(self):
        r, u, f = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().s2_player_stats(self._h, C.byref(r), C.byref(u), C.byref(f)))
        return {"buffers_rendered": r.value, "underruns": u.value, "frames_played": f.value}
