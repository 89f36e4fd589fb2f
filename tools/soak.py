"""Soak: many more seeds of the random-bank parity test than the suite runs, plus random banks through the
time-split mode (all six filter kinds on the default path, the two supported ones time-split).
`PYTHONPATH=.:tests python tools/soak.py [first_seed] [count]` on a GPU box; prints the worst margins."""
import sys

import numpy as np

sys.path.insert(0, "tests")
import test_gpu_parity as T     # noqa: E402
import synth2_b200 as s2        # noqa: E402
from synth2_b200 import bankgen # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
worst = 0.0
for seed in range(first, first + count):
    T.test_random_banks_against_oracle(seed)
print(f"default / one-frame-at-a-time moving cutoff / general / pipelined paths: seeds {first}..{first + count - 1} within tolerance", flush=True)

for seed in range(first, first + count // 2):
    rng = np.random.default_rng(7000 + seed)
    fk = int(rng.integers(0, 6))
    V = int(rng.choice([40, 96, 200]))
    blocks = [int(x) for x in rng.choice([1024, 2048, 1000, 4096], size=3)]
    v = bankgen.make_bank(V, sum(blocks) + 4096, kinds=(0, 1, 2, 3), mod_to_lpf_choices=(0.0, 1.0))
    v["pitch_hz"] *= np.float32(0.5)
    if fk == 3:
        v["damping"] += 2.0
    v["amp_attack_ms"] = rng.choice([1.0, 20.0, 60.0], V).astype(np.float32)
    v["release_offset"] = rng.integers(100, sum(blocks), V).astype(np.uint32)
    ref, _, rst = T.oracle_bank_render(v, fk, blocks)
    got, _, st = T.gpu_bank_render(v, fk, blocks, want_bus=False)
    e, snr = T.assert_parity(ref, got, f"filter {fk} seed {seed}")
    worst = max(worst, e / max(1.0, float(np.max(np.abs(ref)))))
    if fk <= 1:
        got2, st2, n_ts = T._render_blocks(v, fk, blocks, True)
        e2, _ = T.assert_parity(ref, got2, f"time-split filter {fk} seed {seed}")
        worst = max(worst, e2 / max(1.0, float(np.max(np.abs(ref)))))
        assert st2["phase"].tobytes() == rst["phase"].tobytes()
print(f"six filter kinds + time-split: worst max|err| / full scale = {worst:.2e} (bar 1e-4)", flush=True)
