#!/bin/bash
# One GPU-box visit: parity tests, the driver's 20-step bench line, the launch list of the same command.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG [pytest-args]'
tag=${1:-x}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q "$@" > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench20.json 2> gpurun_out/${tag}_bench20.err; echo "bench20 rc=$?"
cat gpurun_out/${tag}_bench20.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
