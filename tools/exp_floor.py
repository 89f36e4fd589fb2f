"""Where does a sustain step go?  Same bank, same state: rows written / nothing written / mix only, pipeline 1 and 4."""
import sys, json
import numpy as np, torch
import synth2_b200 as s2
from synth2_b200 import bankgen

SR, V, T = 48000, 65536, 4096
voices = bankgen.make_bank(V, 60 * SR, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
bus = torch.empty(T, device="cuda")
for pipe in (1, 4):
    bank = s2.VoiceBank(voices, SR, 1, device=0, stream=stream)
    if pipe > 1:
        bank.set_pipeline(pipe)
    for i in range(16):
        bank.render(T, ring[i & 1], T, None)       # past the ramps: sustain
    bank.sync()
    st = bank.get_state()
    for name, rows, mix in (("rows", True, False), ("nothing", False, False), ("mix only", False, True), ("rows + mix", True, True)):
        bank.set_state(st)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            bank.render(T, ring[i & 1] if rows else None, T if rows else 0, bus if mix else None)
        bank.join(stream); torch.cuda.synchronize()
        ev0.record(stream)
        K = 60
        for i in range(K):
            bank.render(T, ring[i & 1] if rows else None, T if rows else 0, bus if mix else None)
        bank.join(stream)
        ev1.record(stream); torch.cuda.synchronize()
        print(f"pipeline {pipe}  {name:12s} {ev0.elapsed_time(ev1) / K * 1e3:8.1f} us/step", flush=True)
    bank.close()
