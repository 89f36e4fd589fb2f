"""Soak: test_random_banks_against_oracle over many seeds.  PYTHONPATH=. python tools/soak_all.py [first] [count]"""
import sys, traceback
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_parity as T
FIRST = int(sys.argv[1]) if len(sys.argv) > 1 else 100
COUNT = int(sys.argv[2]) if len(sys.argv) > 2 else 300
fails = []
for seed in range(FIRST, FIRST + COUNT):
    try:
        T.test_random_banks_against_oracle(seed)
    except AssertionError as e:
        fails.append((seed, str(e)[:160]))
print(len(fails), "failures of", COUNT, "seeds from", FIRST)
for f in fails: print(f)
