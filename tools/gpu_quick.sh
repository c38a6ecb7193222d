#!/bin/bash
# quick loop: conv tests + transform ablation timings + short bench (no baselines)
tag=${1:-R2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv3d.py tests/test_gpu_unet.py -q -m gpu --no-header --tb=short -p no:cacheprovider 2>&1 | tail -6
python tools/bench_epilogue.py --taps 27 2>&1 | grep -E "nothing off|fused input|hand the stage" | tee gpurun_out/${tag}_ablation.txt
python bench.py --steps 20 --warmup 5 --no-torch-gpu-baseline --no-other-configs --no-train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print('ms_per_step', d['ms_per_step'], 'conv_ms', d['roofline']['conv_ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d.get('parity'))"
grep "^\[conv\]" gpurun_out/${tag}_bench.err | head -12
