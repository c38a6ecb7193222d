#!/bin/bash
# ncu --set full capture (with SASS/source counters) of the generic conv kernel on a wide layer; the report comes back whole.
tag=${1:-R3}
mkdir -p gpurun_out
CMD="python tools/ncu_conv_case.py --cin ${CIN:-128} --cout ${COUT:-128} --grid ${GRID:-32} --batch 8 --taps 27"
$CMD > gpurun_out/${tag}_case.log 2>&1 || { echo "case failed"; tail -5 gpurun_out/${tag}_case.log; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_planar -s 3 -c 2 -f \
    -o gpurun_out/${tag}_wide $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/${tag}_wide.ncu-rep
