"""Time of the moving-cutoff blocks (the first 200 ms of every note) at the bench shape: first three 4,096-frame
blocks from note-on, then a sustain block for scale.  `S2_EXP_LIB=path PYTHONPATH=. python tools/exp_modcut.py`"""
import os
import pathlib

import numpy as np
import torch

import synth2_b200._lib as _s2lib
if os.environ.get("S2_EXP_LIB"):
    _s2lib.LIB_PATH = pathlib.Path(os.environ["S2_EXP_LIB"]).resolve()
import synth2_b200 as s2
from synth2_b200 import bankgen

V, T, SR = 65536, 4096, 48000
voices = bankgen.make_bank(V, 2_880_000, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
stream = torch.cuda.current_stream()
bank = s2.VoiceBank(voices, SR, s2.FILTER_BIQUAD_LP, device=0, stream=stream)
bank.set_pipeline(4)
ring = [torch.empty((V, T), device="cuda", dtype=torch.float32) for _ in range(2)]
st0 = bank.get_state()
for rep in range(2):
    bank.set_state(st0)
    torch.cuda.synchronize()
    times = []
    for i in range(16):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bank.render(T, ring[i & 1], T, None)
        bank.join(stream)
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
print("ms per block from note-on:", " ".join(f"{t:.2f}" for t in times))
bank.close()
