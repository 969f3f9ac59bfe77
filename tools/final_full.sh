#!/bin/bash
# tools/final_full.sh TAG: BASELINE configs 3, 4, 5 at full size on one GPU (device-resident leg + e2e, no CPU leg)
TAG=${1:-r02}
O=gpurun_out/final_$TAG
mkdir -p $O
for w in c3 c4 c5; do
  timeout 600 python bench.py --workload $w --scale 1.0 --steps 10 --warmup 3 --no-cpu --no-strong --no-bam --e2e-steps 5 > $O/bench_${w}_full.json 2> $O/bench_${w}_full.err || tail -5 $O/bench_${w}_full.err
  tail -c 300 $O/bench_${w}_full.json
done
