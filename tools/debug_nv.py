import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle, synth2_b200 as s2
from synth2_b200 import bankgen
fk = 0
frames = 4096 * 3
v = bankgen.make_bank(160, 11240, kinds=(0, 1, 2, 3))
v["noise_amt"] = (np.arange(160) % 2) * 0.5
v["release_offset"] = 6000
res = {}
for nv in ("1", "2"):
    os.environ["S2_FORCE_NV"] = nv
    out = torch.zeros((160, frames), device="cuda")
    with s2.VoiceBank(v, 48000, fk) as b:
        b.render(frames, out, frames, None); b.sync()
    res[nv] = out.cpu().numpy()
st = oracle.bank_init_states(v)
ref, _ = oracle.bank_render(v, st, 48000, fk, frames, want_bus=False, nthreads=8)
for vi in (1, 2, 3, 5):
    d12 = np.flatnonzero(res["1"][vi] != res["2"][vi])
    d1r = np.flatnonzero(res["1"][vi] != ref[vi])
    d2r = np.flatnonzero(res["2"][vi] != ref[vi])
    print("voice", vi, "kind", v["osc_kind"][vi], "namt", v["noise_amt"][vi], "A,D,S,R", v["amp_attack_ms"][vi], v["amp_decay_ms"][vi], v["amp_sustain"][vi], v["amp_release_ms"][vi])
    print("  nv1!=nv2:", d12[:8], len(d12), " nv1!=ref:", d1r[:8], len(d1r), " nv2!=ref:", d2r[:8], len(d2r))
    if len(d12):
        i = d12[0]
        print("  at", i, res["1"][vi][i].view(np.uint32) if False else float(res["1"][vi][i]), float(res["2"][vi][i]), float(ref[vi][i]))
