#!/bin/bash
# One GPU-box call: the whole `-m gpu` test suite (per-file logs), then the default bench line.
# usage: tools/gpu_check.sh <tag> [extra bench args]      -> gpurun_out/<tag>_*
tag=${1:-R2}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
t0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu --no-header --tb=short -p no:cacheprovider --durations=15 > gpurun_out/${tag}_tests.txt 2>&1
echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -25 gpurun_out/${tag}_tests.txt
t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
