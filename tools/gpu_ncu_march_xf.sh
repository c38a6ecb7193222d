#!/bin/bash
# ncu --set full capture (with SASS/source counters) of the marching conv kernel WITH the fused GroupNorm + SiLU input transform.
tag=${1:-R5}
mkdir -p gpurun_out
CMD="python tools/ncu_conv_case.py --cin 32 --cout 32 --grid 128 --batch 8 --taps 27 --in-norm"
$CMD > gpurun_out/${tag}_case.log 2>&1 || { echo "case failed"; tail -5 gpurun_out/${tag}_case.log; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_march -s 3 -c 1 -f \
    -o gpurun_out/${tag}_march_xf $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/${tag}_march_xf.ncu-rep
