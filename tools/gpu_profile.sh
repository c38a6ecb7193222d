#!/bin/bash
# bench line + ncu launch list + full ncu captures of the conv / wgrad kernels (B200_PROFILING.md recipe).
# usage: tools/gpu_profile.sh <tag>     -> gpurun_out/<tag>_*
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
# launch list of the sampling leg of the same command (kernel share of a step; times are cold-cache and serialised)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 330 --csv \
    --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu launches rc=$?"
# full capture of the conv kernels of one sampler step (first eager step of the session: 28 launches)
$CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3d_planar -s 0 -c 28 -f -o /tmp/${tag}_conv $CMD > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu full rc=$?"
# the report of 29 launches with source counters is ~75 MB (gpurun_out/ carries 64 MB back): summarise it here
python tools/ncu_summary.py /tmp/${tag}_conv.ncu-rep > gpurun_out/${tag}_conv_ncu.txt 2> gpurun_out/${tag}_summary.err
python tools/ncu_stalls.py /tmp/${tag}_conv.ncu-rep 14 > gpurun_out/${tag}_conv_stalls.txt 2>> gpurun_out/${tag}_summary.err
python tools/launch_share.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_share.txt 2>> gpurun_out/${tag}_summary.err
# training step: launch list of one eager step (profile_step.py drives sampler + trainer; skip to the trainer part)
ls -la gpurun_out | tail -8
