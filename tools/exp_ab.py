"""A/B of library variants on one box: python tools/exp_ab.py lib1.so lib2.so ...  (each in its own process)."""
import os, subprocess, sys, json
CHILD = r'''
import json, sys, numpy as np, torch
import synth2_b200 as s2
from synth2_b200 import bankgen
SR, V, T = 48000, 65536, 4096
voices = bankgen.make_bank(V, 60 * SR, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
stream = torch.cuda.current_stream()
ring = [torch.empty((V, T), device="cuda") for _ in range(2)]
bank = s2.VoiceBank(voices, SR, 1, device=0, stream=stream)
bank.set_pipeline(4)
st0 = bank.get_state()
res = {}
for rep in range(2):
    bank.set_state(st0)
    for i in range(3): bank.render(T, ring[i & 1], T, None)
    bank.set_state(st0); bank.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    ev[0].record(stream)
    for i in range(40):
        bank.render(T, ring[i & 1], T, None); bank.join(stream); ev[i + 1].record(stream)
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(40)]
    res = {"first20_ms": sum(ms[:20]), "modcut3": sum(ms[:3]), "ramp_3_12": sum(ms[3:12]), "sustain_avg": sum(ms[24:40]) / 16}
bank.close()
# the latency-bound regime: half the voices (1.7 warps per scheduler), sustain
vh = voices[:V // 2]
bank = s2.VoiceBank(vh, SR, 1, device=0, stream=stream)
bank.set_pipeline(4)
for i in range(16): bank.render(T, ring[i & 1][:V // 2], T, None)
bank.join(stream); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for i in range(40): bank.render(T, ring[i & 1][:V // 2], T, None)
bank.join(stream); e1.record(stream); torch.cuda.synchronize()
res["half_bank_sustain"] = e0.elapsed_time(e1) / 40
print(json.dumps({k: round(v, 4) for k, v in res.items()}))
'''
for lib in sys.argv[1:]:
    lib, _, extra = lib.partition(":")                   # lib.so:VAR=value sets an environment variable for that run
    env = dict(os.environ, S2_LIB=os.path.abspath(lib), PYTHONPATH=".")
    if extra:
        k, _, v = extra.partition("=")
        env[k] = v
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
    print(f"{os.path.basename(lib) + ' ' + extra:40s} {line}", flush=True)
