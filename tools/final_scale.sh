#!/bin/bash
# tools/final_scale.sh N TAG: the N-GPU bench line and reference arm exactly as the driver launches them
N=${1:-2}; TAG=${2:-r02}
O=gpurun_out/final_$TAG
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 3 > $O/scale_n$N.json 2> $O/scale_n$N.err || tail -5 $O/scale_n$N.err
tail -c 600 $O/scale_n$N.json
