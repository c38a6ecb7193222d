#!/bin/bash
# ncu --set full capture (with SASS/source counters) of the conv kernel on epilogue-bound cases; the report comes back whole.
tag=${1:-R2}
mkdir -p gpurun_out
for taps in 1 27; do
  CMD="python tools/ncu_conv_case.py --cin 32 --cout 32 --grid 128 --batch 2 --taps $taps"
  $CMD > gpurun_out/${tag}_case${taps}.log 2>&1 || { echo "case failed"; tail -5 gpurun_out/${tag}_case${taps}.log; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_planar -s 3 -c 2 -f \
      -o gpurun_out/${tag}_epi${taps} $CMD > gpurun_out/${tag}_ncu${taps}.log 2>&1
  echo "ncu taps=$taps rc=$?"; ls -la gpurun_out/${tag}_epi${taps}.ncu-rep
done
