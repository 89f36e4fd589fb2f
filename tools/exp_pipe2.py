"""Hybrid: the sweep blocks as one full-bank launch each, then 4 voice ranges (a host sync at the switch)."""
import torch
import synth2_b200 as s2
from synth2_b200 import bankgen
SR, V, T = 48000, 65536, 4096
voices = bankgen.make_bank(V, 60 * SR, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
bank = s2.VoiceBank(voices, SR, 1, device=0, stream=stream)
st0 = bank.get_state()
def run(first, switch_at):
    bank.set_pipeline(first)
    bank.set_state(st0)
    for i in range(3): bank.render(T, ring[i & 1], T, None)
    bank.set_state(st0); bank.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(20):
        if i == switch_at: bank.set_pipeline(4)
        bank.render(T, ring[i & 1], T, None)
    bank.join(stream); e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for rep in range(3):
    print("4 ranges throughout        %.3f ms" % run(4, -1))
    print("1 range for 2 blocks, then 4 %.3f ms" % run(1, 2))
    print("1 range for 3 blocks, then 4 %.3f ms" % run(1, 3))
    print("2 ranges for 3 blocks, then 4 %.3f ms" % run(2, 3), flush=True)
