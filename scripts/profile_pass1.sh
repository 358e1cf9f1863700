#!/bin/bash
set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plainA.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pass1_group -s 5 -c 1 -o gpurun_out/prof_pass1group_r1 $CMD > gpurun_out/ncu_p1g.log 2>&1
B200_PASS1_THREAD=1 $CMD > gpurun_out/plainB.log 2>&1 &&
B200_PASS1_THREAD=1 ncu --set full --clock-control none --import-source on -k regex:k_pass1 -s 0 -c 1 -o gpurun_out/prof_pass1thread_r1 $CMD > gpurun_out/ncu_p1t.log 2>&1
grep -o '"phases".*' gpurun_out/plainA.log; grep -o '"phases".*' gpurun_out/plainB.log
