#!/bin/bash
# tools/final_profile.sh TAG: the measurements that go to profiles/ -- GPU suite, default bench line, reference arm, ncu
# launch list of the bench command, one ncu --set full capture of the top kernels.  Run under gpurun (one GPU).
TAG=${1:-r02}
O=gpurun_out/final_$TAG
mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python bench.py > $O/bench_c2.json 2> $O/bench_c2.err || tail -5 $O/bench_c2.err
python bench.py --impl reference > $O/reference_arm_c2.json 2> $O/reference_arm_c2.err || tail -5 $O/reference_arm_c2.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong > $O/b_small.json 2> $O/b_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong > $O/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_fused_prep_tma|k_fused_tile_tma|k_stats_stream" --launch-skip 12 -c 3 \
    -o $O/full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-strong > $O/ncu_full.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_block_expand" -c 1 -o $O/full_blk python tools/e2e_probe.py > $O/ncu_blk.log 2>&1
python tools/bench_bam.py --scale 0.2 --reps 3 > $O/bench_bam.json 2> $O/bench_bam.err
ls -la $O
