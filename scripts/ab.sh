#!/bin/bash
# A/B of library build variants: scripts/ab.sh <lib.so> ... ; prints step / walk / sidm ms of the bench
for lib in "$@"; do
  B200_LIB=$PWD/sidm-nbody_b200/$lib python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases']
print('$lib', 'step %.2f walk %.2f sidm %.2f ensure %.2f build %.2f  I_n %.1f' % (d['ms_per_step'], p['walk_ms'], p['sidm_ms'], p['ensure_ms'], p['build_ms'], d['roofline']['I_n_per_list']))"
done
