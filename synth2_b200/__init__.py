"""synth2_b200 — B200-native renderer for the hot path of brson/synth2.

Only what the path needs: the CUDA library (csrc/ -> libs2cuda.so, C ABI in include/s2_cuda.h) and
the host-side mirror of the reference interface:

    Synth        synth::Synth          s2_lib/src/try3/synth.rs:9-203
    VoiceBank    the batched form of   process::process_layer_buf_simd (process.rs:14-49)
    bankgen      synthetic voice banks of BASELINE.json's shapes
    Player       the two-buffer hand-off to an audio callback   s2_bin/src/audio_player.rs
    patch        `.synth2` patch + score files (example.synth2); `python -m synth2_b200.render` renders one

Importing the package does not need a GPU; creating a Synth / VoiceBank does, and fails loudly
without one (no CPU or PyTorch fallback exists).
"""
from ._lib import (FILTER_BIQUAD_BP, FILTER_BIQUAD_HP, FILTER_BIQUAD_LP, FILTER_FIRST_ORDER_HP, FILTER_FIRST_ORDER_LP,
                   FILTER_ONE_POLE, NO_RELEASE, NOTE_EVENT, OSC_SAW, OSC_SINE, OSC_SQUARE,
                   OSC_TRIANGLE, PATCH, VOICE_DESC, VOICE_STATE, S2Error, lib)
from . import patch
from .bank import VoiceBank, default_voice, note_to_pitch
from .player import Player
from .synth import FrameOffset, Note, Synth, Velocity

__all__ = [
    "Synth", "Player", "Note", "Velocity", "FrameOffset", "VoiceBank", "default_voice", "note_to_pitch",
    "VOICE_DESC", "VOICE_STATE", "PATCH", "NOTE_EVENT", "patch", "S2Error", "lib", "NO_RELEASE",
    "OSC_SQUARE", "OSC_SAW", "OSC_TRIANGLE", "OSC_SINE", "FILTER_ONE_POLE", "FILTER_BIQUAD_LP", "FILTER_BIQUAD_HP", "FILTER_BIQUAD_BP", "FILTER_FIRST_ORDER_LP",
    "FILTER_FIRST_ORDER_HP",
]
