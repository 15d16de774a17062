#!/bin/bash
# bench.py under torchrun at N = $1 GPUs of one box (both arms), as the driver launches it.
set -u
N=${1:-2}
O=gpurun_out
T=${TAG:-r2}
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/${T}_scale_n$N.json 2> $O/${T}_scale_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/${T}_scale_n${N}_ref.json 2>> $O/${T}_scale_n$N.err
wc -l $O/${T}_scale_n$N.json $O/${T}_scale_n${N}_ref.json
cut -c1-300 $O/${T}_scale_n$N.json
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" $O/${T}_scale_n$N.err | tail -5
