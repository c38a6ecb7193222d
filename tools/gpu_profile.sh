#!/bin/bash
# bench line + ncu launch list + one full ncu capture of the conv kernel (B200_PROFILING.md recipe).
# usage: tools/gpu_profile.sh <tag>     -> gpurun_out/<tag>_*
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3d_planar -s 1 -c 4 -o gpurun_out/${tag}_conv $CMD > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
