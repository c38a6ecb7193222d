#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_fullsize.py -x -q -m gpu --no-header --tb=short -p no:cacheprovider 2>&1 | tail -5
for m in 1 1; do
  VDM4CDM_FUSE_GN_NARROW=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-gpu-baseline --no-other-configs --no-train > gpurun_out/ab2_$m.json 2> gpurun_out/ab2_$m.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab2_$m.json")); print("fuse_gn_narrow", $m, "ms_per_step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "conv_ms %.3f" % d["roofline"]["conv_ms_per_step"], "parity", d.get("parity",{}).get("denoiser_rel_l2"))
PY
done
