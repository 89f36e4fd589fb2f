#!/bin/bash
# ncu --set full of ONE launch: $2 = launches to skip, $3 = bench steps
tag=${1:-x}; skip=${2:-3}; steps=${3:-4}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip $skip -c 1 -o gpurun_out/${tag} -f \
  python bench.py --steps $steps --warmup 3 --pipeline 1 --no-e2e --no-cpu-baseline --no-extra --no-parity > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
