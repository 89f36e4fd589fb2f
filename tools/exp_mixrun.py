"""A few sustain blocks with per-voice rows + the mono mix, one launch per block (for ncu):
    ncu --set full --import-source on -k regex:render_kernel --launch-skip 26 -c 1 -o out python tools/exp_mixrun.py [nomix]"""
import sys
import torch
import synth2_b200 as s2
from synth2_b200 import bankgen

SR, V, T = 48000, 65536, 4096
voices = bankgen.make_bank(V, 60 * SR, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
bus = None if "nomix" in sys.argv else torch.empty(T, device="cuda")
bank = s2.VoiceBank(voices, SR, 1, device=0, stream=stream)
for i in range(28):
    bank.render(T, ring[i & 1], T, bus)
bank.sync()
bank.close()
