#!/bin/bash
# round 2, GPU call L: L2 eviction hints, k_connect_rays register budgets, binary occlusion nodes; refill thresholds
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
: > gpurun_out/r02l_sweep.jsonl
for v in base evict rays64 rays40 occbin; do
  for c in c4 c3; do
    settings="mode=0"
    if [ "$v" = "base" ]; then settings="mode=0;mode=0,walk_refill_below=20;mode=0,walk_refill_below=28;mode=0,connect_refill_below=12;mode=0,connect_refill_below=20;mode=0,connect_schedule=6;mode=0,connect_schedule=12"; fi
    CLV_LIBRARY=clive2_b200/_lib/variants/$v.so timeout 400 python tools/sweep_walk.py --config $c --passes 6 --settings "$settings" 2>> gpurun_out/r02l_sweep.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/r02l_sweep.jsonl
  done
done
CLV_LIBRARY=clive2_b200/_lib/variants/evict.so timeout 300 python tools/sweep_walk.py --config c2 --passes 20 --settings "mode=0" 2>> gpurun_out/r02l_sweep.err | sed "s/^{/{\"variant\": \"evict\", /" >> gpurun_out/r02l_sweep.jsonl
CLV_LIBRARY=clive2_b200/_lib/variants/base.so timeout 300 python tools/sweep_walk.py --config c2 --passes 20 --settings "mode=0" 2>> gpurun_out/r02l_sweep.err | sed "s/^{/{\"variant\": \"base\", /" >> gpurun_out/r02l_sweep.jsonl
echo done
