#!/bin/bash
# DRAM bytes + duration of every kernel of ONE graph-replayed... no: of the eager sampler step the bench's roofline leg profiles
# (3 cheap metrics, one pass per kernel).  usage: tools/gpu_traffic.sh <tag>  -> gpurun_out/<tag>_step_launches.csv + _launch_share.txt
tag=${1:-R2}
mkdir -p gpurun_out
CMD="python tools/one_step.py"
$CMD > gpurun_out/${tag}_one_step.log 2>&1 || { echo "one_step failed"; tail -5 gpurun_out/${tag}_one_step.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "profiled_step/" --csv \
    --log-file gpurun_out/${tag}_step_launches.csv $CMD > gpurun_out/${tag}_ncu_step.log 2>&1
echo "ncu rc=$?"
python tools/launch_share.py gpurun_out/${tag}_step_launches.csv > gpurun_out/${tag}_step_launch_share.txt
cat gpurun_out/${tag}_step_launch_share.txt
