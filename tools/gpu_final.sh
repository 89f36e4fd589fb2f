#!/bin/bash
# The round's evidence in one visit: tests, both bench windows, the launch list, ncu --set full of a sustain launch
# and of a moving-cutoff launch.
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1_20steps.json 2> gpurun_out/${tag}_bench20.err; echo "bench20 rc=$?"
timeout 900 python bench.py --gpus 1 > gpurun_out/${tag}_bench_n1_704steps.json 2> gpurun_out/${tag}_bench704.err; echo "bench704 rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_ref.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra --no-parity > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?"
common="--set full --clock-control none --import-source on -k regex:render_kernel -c 1 -f"
timeout 600 ncu $common --launch-skip 23 -o gpurun_out/${tag}_sustain python bench.py --steps 24 --warmup 3 --pipeline 1 --no-e2e --no-cpu-baseline --no-extra --no-parity > gpurun_out/${tag}_ncu_sustain.log 2>&1; echo "ncu sustain rc=$?"
timeout 600 ncu $common --launch-skip 3 -o gpurun_out/${tag}_modcut python bench.py --steps 4 --warmup 3 --pipeline 1 --no-e2e --no-cpu-baseline --no-extra --no-parity > gpurun_out/${tag}_ncu_modcut.log 2>&1; echo "ncu modcut rc=$?"
python - <<PY
import json
for k in ("20steps", "704steps"):
    d = json.load(open("gpurun_out/${tag}_bench_n1_%s.json" % k))
    print(k, "value %.3e frac %.3f e2e %.3e" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
