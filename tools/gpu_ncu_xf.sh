#!/bin/bash
# epilogue / transform ablation timings (bring-up library) + ncu --set full captures of the conv kernel with SASS counters
tag=${1:-R2}
mkdir -p gpurun_out
python tools/bench_epilogue.py --taps 27 > gpurun_out/${tag}_ablation.txt 2>&1; python tools/bench_epilogue.py --taps 1 >> gpurun_out/${tag}_ablation.txt 2>&1
cat gpurun_out/${tag}_ablation.txt
CMD="python tools/ncu_conv_case.py --cin 32 --cout 32 --grid 128 --batch 2 --taps 27 --in-norm"
$CMD > gpurun_out/${tag}_case.log 2>&1 || { echo "case failed"; tail -5 gpurun_out/${tag}_case.log; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_planar -s 4 -c 4 -f \
    -o gpurun_out/${tag}_xf $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/${tag}_xf.ncu-rep
