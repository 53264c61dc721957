#!/bin/bash
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 300 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q -k "wavefront or c3_terrain" > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02n_pytest.log; tail -2 gpurun_out/r02n_pytest.log
timeout 400 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=2;mode=3,slices=1;mode=3,slices=1,pipe_shade_every=6;mode=3,slices=1,pipe_shade_every=8;mode=3,slices=1,pipe_shade_every=0;mode=3,slices=2,walk_grid_divisor=2;mode=3,slices=1,pipe_shade_every=6,walk_refill_below=28" > gpurun_out/r02n_sweep_c4.jsonl 2> gpurun_out/r02n_sweep_c4.err; echo "c4 rc=$?"
timeout 300 python tools/sweep_walk.py --config c4 --passes 8 --tile-shard 2 --settings "mode=2;mode=3,slices=1;mode=3,slices=1,pipe_shade_every=6" > gpurun_out/r02n_sweep_c4_shard2.jsonl 2> gpurun_out/r02n_sweep_c4_shard2.err; echo "c4 shard2 rc=$?"
timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=2;mode=3,slices=1" > gpurun_out/r02n_sweep_c3.jsonl 2> gpurun_out/r02n_sweep_c3.err; echo "c3 rc=$?"
