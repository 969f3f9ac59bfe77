#!/bin/bash
# tools/gpu_ab.sh TAG [pytest -k expr]: GPU parity tests, then the C2 bench with the kernel variants selected through
# the environment (MCOV_*_LEGACY tuning hooks), one summary line per variant.  Run under gpurun.
TAG=${1:-x}
KEXPR=${2:-not full_size}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for v in ${VARIANTS:-new legacy}; do
  (
  if [ $v = legacy ]; then export MCOV_PREP_LEGACY=1 MCOV_TILE_LEGACY=1 MCOV_STATS_LEGACY=1; fi
  if [ $v = prep_legacy ]; then export MCOV_PREP_LEGACY=1; fi
  if [ $v = tile_legacy ]; then export MCOV_TILE_LEGACY=1; fi
  if [ $v = stats_legacy ]; then export MCOV_STATS_LEGACY=1; fi
  for wl in ${WORKLOADS:-c2:1.0}; do
    w=${wl%%:*}; sc=${wl##*:}
    python bench.py --workload $w --scale $sc --steps ${STEPS:-30} --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_${v}_$w.json 2> gpurun_out/bench_${TAG}_${v}_$w.err || tail -5 gpurun_out/bench_${TAG}_${v}_$w.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_${v}_$w.json"))
    print("$v $w", round(d["ms_per_step"], 4), "unpiped", round(d["unpipelined_ms_per_step"], 4),
          {k: (round(x["ms_per_launch"] * 1e3, 1), round(x.get("frac", 0), 3)) for k, x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$v $w failed", e)
PY
  done
  )
done
