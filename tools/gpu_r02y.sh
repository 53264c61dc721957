#!/bin/bash
# vote weights below 8 (leaf steps win earlier) with cooperative leaves
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
S=""
for ws in 1 2 4 6 8; do S="$S;mode=0,walk_schedule=$ws,connect_schedule=$ws,walk_refill_below=24,connect_refill_below=16"; done
for cr in 12 20; do S="$S;mode=0,walk_schedule=8,connect_schedule=8,walk_refill_below=24,connect_refill_below=$cr"; done
for wr in 20; do S="$S;mode=0,walk_schedule=8,connect_schedule=8,walk_refill_below=$wr,connect_refill_below=16"; done
timeout 900 python tools/sweep_walk.py --config c4 --passes 5 --settings "${S:1}" > gpurun_out/r02y_c4.jsonl 2> gpurun_out/r02y_c4.err; echo "c4 rc=$?"
S=""
for ws in 2 4 8 16; do S="$S;mode=0,walk_schedule=$ws,connect_schedule=$ws,walk_refill_below=24,connect_refill_below=16"; done
timeout 900 python tools/sweep_walk.py --config c3 --passes 8 --settings "${S:1}" > gpurun_out/r02y_c3.jsonl 2> gpurun_out/r02y_c3.err; echo "c3 rc=$?"
