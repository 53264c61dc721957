#!/bin/bash
# round 2, GPU call G (N GPUs): strong scaling of one C4 frame, reference arm under torchrun, movie (C5) frames dealt to the ranks
N=${1:-2}
mkdir -p gpurun_out
P=$((29500 + RANDOM % 300))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02g_bench_c4_n$N.json 2> gpurun_out/r02g_bench_c4_n$N.err; echo "bench c4 n$N rc=$?"
tail -c 600 gpurun_out/r02g_bench_c4_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02g_bench_c4_ref_n$N.json 2> gpurun_out/r02g_bench_c4_ref_n$N.err; echo "ref n$N rc=$?"
FR=${FRAMES:-120}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+2)) -m clive2_b200.movie movie --scene bench-spheres --width 1280 --height 720 --samples 32 --movie-frames $FR --out-dir /tmp/clv_movie > gpurun_out/r02g_movie_n$N.log 2>&1; echo "movie n$N rc=$?"
grep "frames in" gpurun_out/r02g_movie_n$N.log | head -8
cp /tmp/clv_movie/test-movie/frame_0007.png gpurun_out/r02g_movie_frame_0007_n$N.png 2>/dev/null
nproc; free -g | head -2
