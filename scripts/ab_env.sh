#!/bin/bash
# scripts/ab_env.sh <gpus> "<ENV=..>" ... : bench under different environment settings, one line each
G=$1; shift
for e in "$@"; do
  if [ "$G" = "1" ]; then CMD="python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 4 --warmup 3 --no-cpu --no-e2e"; fi
  env $e $CMD 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases']; r=d['roofline']
print('$G gpus [$e]', 'step %.2f walk %.2f (overlapped span %.2f) sidm %.2f ensure %.2f build %.2f' % (d['ms_per_step'], p['walk_ms'], r['ms_per_launch_while_sidm_overlaps'], p['sidm_ms'], p['ensure_ms'], p['build_ms']))"
done
