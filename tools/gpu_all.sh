#!/bin/bash
# tests (stop at first failure), then both bench windows
tag=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_tests.log
bash tools/gpu_bench.sh $tag
