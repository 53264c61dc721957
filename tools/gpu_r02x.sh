#!/bin/bash
# vote weights / refill thresholds once the leaves are tested by the whole warp
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
S=""
for ws in 8 16 32 1000; do for cs in 8 16 32 1000; do S="$S;mode=0,walk_schedule=$ws,connect_schedule=$cs,walk_refill_below=24,connect_refill_below=16"; done; done
for wr in 16 28 32; do S="$S;mode=0,walk_schedule=8,connect_schedule=8,walk_refill_below=$wr,connect_refill_below=16"; done
for cr in 8 24 32; do S="$S;mode=0,walk_schedule=8,connect_schedule=8,walk_refill_below=24,connect_refill_below=$cr"; done
timeout 900 python tools/sweep_walk.py --config c4 --passes 5 --settings "${S:1}" > gpurun_out/r02x_c4.jsonl 2> gpurun_out/r02x_c4.err; echo "c4 rc=$?"
