#!/bin/bash
# leaves tested by the whole warp (CLV_COOP_LEAVES): parity, then A/B against the per-lane leaf loops
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02w_pytest.log
V=clive2_b200/_lib/variants
for v in main nocoop rays64 rays48; do
  lib=$V/$v.so; [ $v = main ] && lib=clive2_b200/_lib/libclive_b200.so
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0" > gpurun_out/r02w_c4_$v.jsonl 2> gpurun_out/r02w_c4_$v.err; echo "c4 $v rc=$?"
done
for v in main nocoop; do
  lib=$V/$v.so; [ $v = main ] && lib=clive2_b200/_lib/libclive_b200.so
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0" > gpurun_out/r02w_c3_$v.jsonl 2> gpurun_out/r02w_c3_$v.err; echo "c3 $v rc=$?"
done
