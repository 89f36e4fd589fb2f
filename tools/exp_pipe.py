"""First 20 blocks of the bench bank with 1 / 2 / 4 / 8 voice ranges (s2_bank_set_pipeline), interleaved."""
import torch
import synth2_b200 as s2
from synth2_b200 import bankgen
SR, V, T = 48000, 65536, 4096
voices = bankgen.make_bank(V, 60 * SR, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
banks = {}
for pipe in (1, 2, 4, 8):
    b = s2.VoiceBank(voices, SR, 1, device=0, stream=stream)
    if pipe > 1: b.set_pipeline(pipe)
    banks[pipe] = (b, b.get_state())
for rep in range(3):
    for pipe, (bank, st0) in banks.items():
        bank.set_state(st0)
        for i in range(3): bank.render(T, ring[i & 1], T, None)
        bank.set_state(st0); bank.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        ev[0].record(stream)
        for i in range(20):
            bank.render(T, ring[i & 1], T, None); bank.join(stream); ev[i + 1].record(stream)
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(20)]
        print(f"ranges {pipe}: 20 blocks {sum(ms):.3f} ms; sweep {sum(ms[:3]):.3f}; ramps 3-11 {sum(ms[3:12]):.3f}; last 8 {sum(ms[12:]):.3f}", flush=True)
