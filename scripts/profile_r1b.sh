#!/bin/bash
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 3 -c 1 -o gpurun_out/prof_walk_r1b $CMD > gpurun_out/ncu_walk.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pass1 -s 0 -c 1 -o gpurun_out/prof_pass1_r1b $CMD > gpurun_out/ncu_pass1.log 2>&1
