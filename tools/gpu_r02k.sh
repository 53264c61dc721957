#!/bin/bash
N=${1:-2}; K=${2:-21}
mkdir -p gpurun_out
P=$((29500 + RANDOM % 300))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps $K --warmup 3 > gpurun_out/r02k_bench_c4_n${N}_k$K.json 2> gpurun_out/r02k_bench_c4_n${N}_k$K.err; echo "bench c4 n$N k$K rc=$?"
tail -c 1500 gpurun_out/r02k_bench_c4_n${N}_k$K.err
