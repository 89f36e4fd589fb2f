"""Do the few warps whose lane pairs hold two followers pace the sweep blocks?  The bench bank (random 50/50
followers: 10 such warps of 2,048, some in every voice range) against the same bank with the followers dealt exactly
half and half inside each oscillator kind (none).  First 20 blocks, ms."""
import numpy as np, torch
import synth2_b200 as s2
from synth2_b200 import bankgen
SR, V, T = 48000, 65536, 4096
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
def run(voices, label):
    bank = s2.VoiceBank(voices, SR, 1, device=0, stream=stream)
    bank.set_pipeline(4)
    st0 = bank.get_state()
    best = None
    for rep in range(3):
        bank.set_state(st0)
        for i in range(3): bank.render(T, ring[i & 1], T, None)
        bank.set_state(st0); bank.sync()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        ev[0].record(stream)
        for i in range(20):
            bank.render(T, ring[i & 1], T, None); bank.join(stream); ev[i + 1].record(stream)
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(20)]
        if best is None or sum(ms) < sum(best): best = ms
    print(f"{label:28s} 20 blocks {sum(best):.3f} ms; sweep blocks {best[0]:.3f} {best[1]:.3f} {best[2]:.3f}", flush=True)
    bank.close()
base = bankgen.make_bank(V, 60 * SR, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
bal = base.copy()
idx = np.arange(V)
bal["mod_env_to_lpf_freq"] = np.where((idx >> 1) & 1, 1.5, 0.0).astype(np.float32)     # kind = idx & 1: exactly half per kind
for _ in range(2):
    run(base, "bench bank (random 50/50)")
    run(bal, "exactly half per kind")
