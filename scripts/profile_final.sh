#!/bin/bash
# profiling recipe of B200_PROFILING.md on the bench command (plain run first, then ncu)
set -x
TAG=${1:-r1h}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 3 -c 1 -o gpurun_out/prof_walk_$TAG $CMD > gpurun_out/ncu_walk_$TAG.log 2>&1
$CMD > gpurun_out/plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pass1_group -s 3 -c 1 -o gpurun_out/prof_pass1_$TAG $CMD > gpurun_out/ncu_pass1_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-200
