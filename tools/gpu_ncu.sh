#!/bin/bash
# ncu --set full of one moving-cutoff launch and one sustain launch (full-bank launches: --pipeline 1)
tag=${1:-x}
mkdir -p gpurun_out
common="--set full --clock-control none --import-source on -k regex:render_kernel"
timeout 600 ncu $common --launch-skip 3 -c 1 -o gpurun_out/${tag}_modcut -f python bench.py --steps 4 --warmup 3 --pipeline 1 --no-e2e --no-cpu-baseline --no-extra --no-parity > gpurun_out/${tag}_ncu_modcut.log 2>&1; echo "ncu modcut rc=$?"
timeout 600 ncu $common --launch-skip 23 -c 1 -o gpurun_out/${tag}_sustain -f python bench.py --steps 24 --warmup 3 --pipeline 1 --no-e2e --no-cpu-baseline --no-extra --no-parity > gpurun_out/${tag}_ncu_sustain.log 2>&1; echo "ncu sustain rc=$?"
ls -la gpurun_out/${tag}_*.ncu-rep
