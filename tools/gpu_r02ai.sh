#!/bin/bash
# k_walk_shade with its independent loads issued first (walk state, prefetch of the hit triangle's shading record): parity, A/B
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q -k "suspend or c3_terrain or wavefront" > gpurun_out/r02ai_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02ai_pytest.log
V=clive2_b200/_lib/variants
for v in main lateloads; do
  lib=$V/$v.so; [ $v = main ] && lib=clive2_b200/_lib/libclive_b200.so
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0" > gpurun_out/r02ai_c4_$v.jsonl 2> gpurun_out/r02ai_c4_$v.err; echo "c4 $v rc=$?"
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0" > gpurun_out/r02ai_c3_$v.jsonl 2> gpurun_out/r02ai_c3_$v.err; echo "c3 $v rc=$?"
done
