#!/bin/bash
# Quick GPU check after a conv kernel change: conv / UNet parity tests, then the sampling leg of the bench twice.
# usage: tools/gpu_quick.sh <tag>
tag=${1:-Q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv3d.py tests/test_gpu_unet.py -x -q -m gpu --no-header --tb=short -p no:cacheprovider 2>&1 | tail -3
for r in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-gpu-baseline --no-other-configs --no-train --no-cpu-baseline > gpurun_out/${tag}_q$r.json 2> gpurun_out/${tag}_q$r.err || { echo "bench failed"; tail -5 gpurun_out/${tag}_q$r.err; continue; }
  python -c "
import json
d=json.load(open('gpurun_out/${tag}_q$r.json')); print('run $r ms_per_step %.3f' % d['ms_per_step'], 'conv_ms %.3f' % d['roofline']['conv_ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'])"
  grep "^\[conv\] 32->32 taps=27\|^\[conv\] 32->1 \|^\[conv\] 16->32" gpurun_out/${tag}_q$r.err | head -3
done
