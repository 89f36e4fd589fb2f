"""VoiceBank — host-side handle of s2_bank: V independent voices resident on one GPU.

Mirrors `process::process_layer_buf_simd(&sc::Layer, &mut st::Layer, pitch, sample_rate, offset,
release_offset, buf)` (s2_lib/src/try3/process.rs:14-22) batched over voices: each render call
fills `frames` frames per voice (x16 blocks then the `frames % 16` scalar tail) and carries
`st::Layer` + the frame offset to the next call.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import VOICE_DESC, VOICE_STATE, check, lib, ptr


def note_to_pitch(note: int) -> float:
    """synth.rs:208-212 (host libm powf inside libs2cuda.so)."""
    return float(lib().s2_note_to_pitch(int(note) & 0xFF))


def default_voice(n: int = 1) -> np.ndarray:
    """`Synth::default_config()` (synth.rs:125-152) as n inactive voice descriptions."""
    one = np.zeros(1, dtype=VOICE_DESC)
    lib().s2_default_voice(ptr(one))
    return np.repeat(one, n)


class VoiceBank:
    def __init__(self, voices: np.ndarray, sample_rate: int = 48000,
                 filter_kind: int = _lib.FILTER_ONE_POLE, device: int = 0, stream=None):
        voices = np.ascontiguousarray(voices, dtype=VOICE_DESC)
        self._h = C.c_void_p()
        self.n_voices = int(voices.shape[0])
        self.sample_rate = int(sample_rate)
        self.filter_kind = int(filter_kind)
        self.device = int(device)
        stream_ptr = None
        if stream is not None:
            stream_ptr = C.c_void_p(getattr(stream, "cuda_stream", stream))
        check(lib().s2_bank_create(self.device, self.sample_rate, self.filter_kind, self.n_voices,
                                   ptr(voices), stream_ptr, C.byref(self._h)))

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().s2_bank_destroy(self._h)
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

    # -- voices
    def set_voice(self, index: int, voice: np.ndarray):
        voice = np.ascontiguousarray(voice, dtype=VOICE_DESC).reshape(-1)[:1]
        check(lib().s2_bank_set_voice(self._h, int(index), ptr(voice)))

    def release_voice(self, index: int):
        check(lib().s2_bank_release_voice(self._h, int(index)))

    def set_releases(self, release_offsets):
        """Bulk note_off table: u32 per voice (NO_RELEASE = held); numpy array or pinned torch tensor."""
        check(lib().s2_bank_set_releases(self._h, ptr(release_offsets)))

    # -- rendering (device buffers: anything with .data_ptr(), or a raw address)
    def render(self, frames: int, voice_out=None, row_stride: int = 0, bus_out=None):
        if voice_out is not None and not row_stride:
            row_stride = int(voice_out.stride(0)) if hasattr(voice_out, "stride") else int(frames)
        check(lib().s2_bank_render(self._h, int(frames), ptr(voice_out), int(row_stride), ptr(bus_out)))

    def render_bus_host(self, frames: int, voice_out=None, row_stride: int = 0, out: np.ndarray = None):
        """Render and return the mono mix in host memory (the `Synth::sample` shape)."""
        if out is None:
            out = np.empty(int(frames), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size >= frames
        if voice_out is not None and not row_stride:
            row_stride = int(voice_out.stride(0)) if hasattr(voice_out, "stride") else int(frames)
        check(lib().s2_bank_render_bus_host(self._h, int(frames), ptr(voice_out), int(row_stride), ptr(out)))
        return out

    def render_bus_host_async(self, frames: int, pinned_out, voice_out=None, row_stride: int = 0):
        """Streaming form: enqueue render + mix + D2H of the mix into pinned host memory, no sync."""
        if voice_out is not None and not row_stride:
            row_stride = int(voice_out.stride(0)) if hasattr(voice_out, "stride") else int(frames)
        check(lib().s2_bank_render_bus_host_async(self._h, int(frames), ptr(voice_out), int(row_stride), ptr(pinned_out)))

    def trace_phase(self, frames: int, phase_out, row_stride: int = 0):
        if not row_stride:
            row_stride = int(phase_out.stride(0)) if hasattr(phase_out, "stride") else int(frames)
        check(lib().s2_bank_trace_phase(self._h, int(frames), ptr(phase_out), int(row_stride)))

    def sync(self):
        check(lib().s2_bank_sync(self._h))

    def set_pipeline(self, n_sub: int):
        """n_sub > 1: render contiguous voice ranges on n_sub internal streams (see include/s2_cuda.h)."""
        check(lib().s2_bank_set_pipeline(self._h, int(n_sub)))

    def set_time_split(self, enable: bool = True):
        """Narrow banks: render qualifying blocks as 32 time segments per voice (s2_cuda.h)."""
        check(lib().s2_bank_set_time_split(self._h, 1 if enable else 0))

    @property
    def time_split_blocks(self) -> int:
        """Blocks rendered through the time-split kernels so far."""
        n = C.c_uint64(0)
        check(lib().s2_bank_time_split_blocks(self._h, C.byref(n)))
        return int(n.value)

    def join(self, stream=None):
        """Make `stream` (a torch stream, a raw cudaStream_t, or None = default) wait for the bank's work."""
        sp = None if stream is None else C.c_void_p(getattr(stream, "cuda_stream", stream))
        check(lib().s2_bank_join(self._h, sp))

    # -- carried state
    def get_state(self) -> np.ndarray:
        st = np.zeros(self.n_voices, dtype=VOICE_STATE)
        check(lib().s2_bank_get_state(self._h, ptr(st)))
        return st

    def set_state(self, state: np.ndarray):
        state = np.ascontiguousarray(state, dtype=VOICE_STATE)
        assert state.shape[0] == self.n_voices
        check(lib().s2_bank_set_state(self._h, ptr(state)))
