#!/bin/bash
# bench line + ncu launch list + full ncu capture of the conv kernels of one sampler step (B200_PROFILING.md recipe).
# usage: tools/gpu_profile.sh <tag>     -> gpurun_out/<tag>_*     (summaries are written on the box: the reports are too big to bring back)
tag=${1:-R2}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-torch-gpu-baseline --no-other-configs --pipeline-steps 0 --e2e-chain 3"
# 1. the command exits 0 without ncu
$CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "plain rc=$?"
# 2. launch list (kernel share of a step; times are cold-cache and serialised), DRAM bytes per launch
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu launches rc=$?"
python tools/launch_share.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_share.txt 2> gpurun_out/${tag}_summary.err
head -30 gpurun_out/${tag}_launch_share.txt
# 3. full capture of the conv kernels of one sampler step (the first eager step of the session)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv3d_(planar|march)" -s 0 -c 38 -f -o /tmp/${tag}_conv $CMD > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu full rc=$?"
timeout 300 python tools/ncu_summary.py /tmp/${tag}_conv.ncu-rep > gpurun_out/${tag}_conv_ncu.txt 2>> gpurun_out/${tag}_summary.err
# (bounded: exporting the source page of a 38-kernel report took longer than the rest of the call in R5n and ate the GPU budget)
timeout 240 python tools/ncu_stalls.py /tmp/${tag}_conv.ncu-rep 12 > gpurun_out/${tag}_conv_stalls.txt 2>> gpurun_out/${tag}_summary.err
ls -la gpurun_out | tail -8
