import os, sys
import numpy as np, torch
sys.path.insert(0, "tests")
import synth2_b200 as s2
from synth2_b200 import bankgen
from test_gpu_parity import gpu_bank_render, bank_for

for fk in (0, 1):
    v = bank_for(fk, 64, 8192, kinds=(1, 0))
    res = {}
    for path in ("0", "1", "2"):
        os.environ["S2_FORCE_PATH"] = path
        res[path + "a"] = gpu_bank_render(v, fk, [8192])[0]
        res[path + "b"] = gpu_bank_render(v, fk, [4096, 2048, 16, 2032])[0]
    ref = res["2a"]
    for k, x in res.items():
        bad = np.argwhere(x != ref)
        print("fk", fk, k, "differs from 2a at", len(bad), "first", bad[:3].tolist(), flush=True)
        if len(bad):
            vv, ff = bad[0]
            print("   voice", vv, {n: v[n][vv] for n in ("osc_kind", "lpf_freq_hz", "damping", "amp_attack_ms", "amp_decay_ms", "mod_env_to_lpf_freq")},
                  "vals", x[vv, ff], ref[vv, ff])
