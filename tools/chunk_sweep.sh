#!/bin/bash
# statistics chunk-length sweep (MCOV_STAT_CHUNK tuning hook), prints per-kernel times
for c in "$@"; do
  MCOV_STAT_CHUNK=$c python bench.py --no-cpu --no-e2e --steps 20 --warmup 3 $BENCH_ARGS | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk', $c, 'ms/step %.4f' % d['ms_per_step'], {k: round(v['ms_per_launch']*1000,1) for k,v in d['roofline']['kernels'].items()})"
done
