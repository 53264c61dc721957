#!/bin/bash
# round 2, GPU call D: prefetch options and shade register caps
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
: > gpurun_out/r02d_sweep.jsonl
for v in ${VARIANTS:-pfleaf pfnext wide_pfleaf wide_pfnext sh48 sh40 sh32}; do
  for c in ${CONFIGS:-c4 c3}; do
    CLV_LIBRARY=clive2_b200/_lib/variants/$v.so python tools/sweep_walk.py --config $c --passes 6 --settings "${SETTINGS:-mode=2}" 2>> gpurun_out/r02d_sweep.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/r02d_sweep.jsonl
  done
done
echo "done"
