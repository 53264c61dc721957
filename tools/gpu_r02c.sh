#!/bin/bash
# round 2, GPU call C: build variants of the wavefront trace / shade kernels on c4 and c3 (one process per variant, scene cached)
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
: > gpurun_out/r02c_sweep.jsonl
for v in base r64 r40 r32 dfs dfs_r40 dfs_r32 wide wide_r64 wide_r40 sh96 sh80 sh64; do
  for c in c4 c3; do
    settings="mode=2"
    if [ "$v" = "dfs" ] || [ "$v" = "base" ]; then settings="mode=2;mode=1"; fi
    CLV_LIBRARY=clive2_b200/_lib/variants/$v.so python tools/sweep_walk.py --config $c --passes 6 --settings "$settings" 2>> gpurun_out/r02c_sweep.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/r02c_sweep.jsonl
  done
done
echo "done rc=$?"
