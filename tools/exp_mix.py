"""Why does rows + mix cost 22 % more than rows?  Host enqueue time vs device time, pipeline 1 / 4."""
import time
import numpy as np, torch
import synth2_b200 as s2
from synth2_b200 import bankgen

SR, V, T = 48000, 65536, 4096
voices = bankgen.make_bank(V, 60 * SR, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
bus = torch.empty(T, device="cuda")
for pipe in (1, 2, 4, 8):
    bank = s2.VoiceBank(voices, SR, 1, device=0, stream=stream)
    if pipe > 1:
        bank.set_pipeline(pipe)
    for i in range(16):
        bank.render(T, ring[i & 1], T, None)
    bank.sync()
    st = bank.get_state()
    for name, rows, mix in (("rows", True, False), ("rows + mix", True, True), ("mix only", False, True)):
        bank.set_state(st)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            bank.render(T, ring[i & 1] if rows else None, T if rows else 0, bus if mix else None)
        bank.join(stream); torch.cuda.synchronize()
        K = 100
        t0 = time.perf_counter()
        ev0.record(stream)
        for i in range(K):
            bank.render(T, ring[i & 1] if rows else None, T if rows else 0, bus if mix else None)
        t1 = time.perf_counter()
        bank.join(stream)
        ev1.record(stream); torch.cuda.synchronize()
        print(f"pipeline {pipe}  {name:12s} device {ev0.elapsed_time(ev1) / K * 1e3:7.1f} us/step   host enqueue {(t1 - t0) / K * 1e6:7.1f} us/step", flush=True)
    bank.close()
