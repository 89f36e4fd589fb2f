import sys, traceback
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_parity as T
fails = []
for seed in range(100, 400):
    try:
        T.test_random_banks_against_oracle(seed)
    except AssertionError as e:
        fails.append((seed, str(e)[:160]))
print(len(fails), "failures of 300")
for f in fails: print(f)
