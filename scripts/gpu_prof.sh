#!/bin/bash
# usage: scripts/gpu_prof.sh TAG kernel1 [kernel2 ...]   -- one ncu --set full capture per kernel regex (3rd launch), after a plain run
TAG=$1; shift
mkdir -p gpurun_out
python scripts/prof_step.py 1000000 4 > gpurun_out/${TAG}_plain.log 2>&1 || { cat gpurun_out/${TAG}_plain.log; exit 1; }
cat gpurun_out/${TAG}_plain.log
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/${TAG}_$K python scripts/prof_step.py 1000000 4 > gpurun_out/${TAG}_ncu_$K.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_$K.log
done
