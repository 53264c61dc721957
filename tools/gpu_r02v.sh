#!/bin/bash
# k_connect_rays after the nearest-child selection stopped indexing its arrays at run time (no more local-memory arrays),
# and its register cap revisited
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
V=clive2_b200/_lib/variants
for v in main before rays56 rays64 rays40; do
  lib=$V/$v.so; [ $v = main ] && lib=clive2_b200/_lib/libclive_b200.so
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0" > gpurun_out/r02v_c4_$v.jsonl 2> gpurun_out/r02v_c4_$v.err; echo "c4 $v rc=$?"
done
for v in main before; do
  lib=$V/$v.so; [ $v = main ] && lib=clive2_b200/_lib/libclive_b200.so
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0" > gpurun_out/r02v_c3_$v.jsonl 2> gpurun_out/r02v_c3_$v.err; echo "c3 $v rc=$?"
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c2 --passes 16 --settings "mode=0" > gpurun_out/r02v_c2_$v.jsonl 2> gpurun_out/r02v_c2_$v.err; echo "c2 $v rc=$?"
done
