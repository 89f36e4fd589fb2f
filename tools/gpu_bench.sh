#!/bin/bash
# Bench lines only: the driver's 20-step window and the whole 60 s render.
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench20.json 2> gpurun_out/${tag}_bench20.err; echo "bench20 rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench20.json"))
print("20 steps: value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
print("step_ms", d.get("step_ms"))
print("parity", d.get("parity"))
print("extra", json.dumps(d.get("extra"))[:1500])
PY
timeout 900 python bench.py --gpus 1 --no-extra > gpurun_out/${tag}_bench704.json 2> gpurun_out/${tag}_bench704.err; echo "bench704 rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench704.json"))
print("704 steps: value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
s = d.get("step_ms") or []
print("step_ms first 30", s[:30], "median", sorted(s)[len(s)//2] if s else None)
print("parity", d.get("parity"))
PY
tail -n 3 gpurun_out/${tag}_bench20.err gpurun_out/${tag}_bench704.err
