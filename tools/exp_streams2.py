"""What does an event record per step cost a pipelined render, and on which stream?  Rows only, sustain, warm GPU,
variants interleaved: none / on the bank's (caller's) stream / on an idle stream / on an idle stream with a wait."""
import sys, time
import torch
import synth2_b200 as s2
from synth2_b200 import bankgen
V, T, SR = 65536, 4096, 48000
voices = bankgen.make_bank(V, 2_880_000, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
idle_before = torch.cuda.Stream()
main = torch.cuda.Stream() if "own" in sys.argv else torch.cuda.current_stream()
torch.cuda.set_stream(main)
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
bank = s2.VoiceBank(voices, SR, s2.FILTER_BIQUAD_LP, device=0, stream=main)
bank.set_pipeline(4)
idle_after = torch.cuda.Stream()
ev = torch.cuda.Event()
for i in range(16): bank.render(T, ring[i & 1], T, None)
bank.join(main); torch.cuda.synchronize()
st = bank.get_state()
def run(kind, K=200):
    bank.set_state(st); bank.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(main)
    for i in range(K):
        bank.render(T, ring[i & 1], T, None)
        if kind == "bank stream": ev.record(main)
        elif kind == "idle stream (created before the bank)": ev.record(idle_before)
        elif kind == "idle stream (created after)": ev.record(idle_after)
        elif kind == "idle stream waits for the bank's stream event": ev.record(main); idle_after.wait_event(ev)
    host = (time.perf_counter() - t0) / K * 1e6
    bank.join(main); e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3, host
for _ in range(6): run("none")
kinds = ["none", "bank stream", "idle stream (created before the bank)", "idle stream (created after)", "idle stream waits for the bank's stream event"]
for rep in range(3):
    for k in kinds:
        d, h = run(k)
        print(f"{k:50s} {d:7.1f} us/step   host loop {h:6.1f} us/step", flush=True)
bank.close()
