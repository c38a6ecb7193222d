#!/bin/bash
# First-contact battery: hardware probe, then every GPU test file in its own process.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 180 tools/bin/probe_umma > gpurun_out/probe.txt 2>&1; echo "probe rc=$?"
for t in "$@"; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --no-header --tb=short -p no:cacheprovider > gpurun_out/test_$t.txt 2>&1
  echo "$t rc=$?"
  tail -3 gpurun_out/test_$t.txt
done
grep -c "mismatches=0" gpurun_out/probe.txt; grep "T5" gpurun_out/probe.txt | head -30
