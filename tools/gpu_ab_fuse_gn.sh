#!/bin/bash
# A/B of the fused GroupNorm + SiLU input transform on the narrow (marching-schedule) layers, same box, same command.
mkdir -p gpurun_out
for m in 0 1 0 1; do
  VDM4CDM_FUSE_GN_NARROW=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-gpu-baseline --no-other-configs --no-train --no-cpu-baseline > gpurun_out/ab_fuse_$m.json 2> gpurun_out/ab_fuse_$m.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_fuse_$m.json")); print("fuse_gn_narrow", $m, "ms_per_step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "conv_ms %.3f" % d["roofline"]["conv_ms_per_step"], "sm_mhz", d["clocks"]["sm_mhz"])
PY
done
