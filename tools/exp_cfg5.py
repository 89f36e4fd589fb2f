"""Per-block times of BASELINE config 5 (32,768 patch variants, biquad, first 2 s from note-on)."""
import torch
import synth2_b200 as s2
from synth2_b200 import bankgen
SR, T = 48000, 4096
V = bankgen.SWEEP_VARIANTS
voices = bankgen.make_sweep_bank(0, 10 * SR)
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
bank = s2.VoiceBank(voices, SR, s2.FILTER_BIQUAD_LP, device=0, stream=stream)
bank.set_pipeline(4)
st0 = bank.get_state()
for rep in range(2):
    bank.set_state(st0); bank.sync()
    K = 24
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record(stream)
    for i in range(K):
        bank.render(T, ring[i & 1], T, None); bank.join(stream); ev[i + 1].record(stream)
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
print("ms per block:", " ".join(f"{m:.3f}" for m in ms), " total", f"{sum(ms):.3f}")
bank.close()
