"""Does the sustain step time drift with the frame offset or with the time the GPU has been writing?
Two passes of 520 blocks from note-on, back to back; mean step time over blocks 64-128 and 448-512 of each."""
import torch
import synth2_b200 as s2
from synth2_b200 import bankgen
SR, V, T = 48000, 65536, 4096
voices = bankgen.make_bank(V, 60 * SR, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
bank = s2.VoiceBank(voices, SR, 1, device=0, stream=stream)
bank.set_pipeline(4)
st0 = bank.get_state()
for p in range(3):
    bank.set_state(st0); bank.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(521)]
    ev[0].record(stream)
    for i in range(520):
        bank.render(T, ring[i & 1], T, None); bank.join(stream); ev[i + 1].record(stream)
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(520)]
    print(f"pass {p}: blocks 64-128 {sum(ms[64:128]) / 64:.4f} ms   256-320 {sum(ms[256:320]) / 64:.4f}   448-512 {sum(ms[448:512]) / 64:.4f}", flush=True)
bank.close()
