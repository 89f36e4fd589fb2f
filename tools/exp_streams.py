"""Experiment: render the 65,536-voice bank as K independent sub-banks on K CUDA streams (no barrier
between blocks across sub-banks) and time the steady state."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import synth2_b200 as s2
from synth2_b200 import bankgen

V, T = 65536, 4096
voices = bankgen.make_bank(V, 2880000, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
ring = [torch.empty((V, T), device="cuda", dtype=torch.float32) for _ in range(2)]
for K in (1, 2, 4, 8, 16):
    streams = [torch.cuda.Stream() for _ in range(K)]
    per = V // K
    banks = [s2.VoiceBank(voices[k * per:(k + 1) * per], 48000, 1, stream=streams[k]) for k in range(K)]
    def run(nblocks):
        for i in range(nblocks):
            for k in range(K):
                banks[k].render(T, ring[i & 1][k * per:(k + 1) * per], T, None)
    run(40)     # past the envelope transients
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev0.record(torch.cuda.current_stream())
    for s in streams:
        s.wait_event(ev0)
    N = 200
    run(N)
    for k in range(K):
        ends[k].record(streams[k])
    torch.cuda.synchronize()
    ms = max(ev0.elapsed_time(e) for e in ends)
    print(f"K={K}: {ms / N * 1e3:.1f} us/block  {V * T * N / (ms * 1e-3):.4e} voice-samples/s  {V*T*4*N/(ms*1e-3)/6525.2e9*100:.1f}% of measured HBM")
    for b in banks:
        b.close()
