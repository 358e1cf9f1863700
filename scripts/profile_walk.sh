#!/bin/bash
# plain run, then one full ncu capture of the walk kernel of the same command (B200_PROFILING.md)
TAG=${1:-r1d}
KERN=${2:-k_walk}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERN -s 3 -c 1 -o gpurun_out/prof_walk_$TAG $CMD > gpurun_out/ncu_walk_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
