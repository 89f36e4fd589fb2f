"""Soak at other sample rates: test_random_banks_against_oracle with the test module's rate replaced.
theta_of (s2_cutoff.h) is the IEEE quotient at the usual rates and within an ulp at any other."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_parity as T
for sr in (44100, 22050, 96000, 37123):
    T.SR = sr
    T.gpu_bank_render.__defaults__ = (True, sr)            # (want_bus, sr)
    T.oracle_bank_render.__defaults__ = (sr, 8)            # (sr, nthreads)
    fails = []
    for seed in range(2000, 2120):
        try:
            T.test_random_banks_against_oracle(seed)
        except AssertionError as e:
            fails.append((seed, str(e)[:120]))
    print(sr, len(fails), "failures of 120", fails[:3], flush=True)
