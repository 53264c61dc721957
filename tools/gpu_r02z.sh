#!/bin/bash
# full GPU suite at HEAD (cooperative leaves, vote weight 4), then the cache-resident flavour with cooperative leaves (C1 / C2)
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02z_pytest.log
V=clive2_b200/_lib/variants
for v in main resident; do
  lib=$V/$v.so; [ $v = main ] && lib=clive2_b200/_lib/libclive_b200.so
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c2 --passes 16 --settings "mode=0;mode=0,connect_schedule=4;mode=0,connect_schedule=8" > gpurun_out/r02z_c2_$v.jsonl 2> gpurun_out/r02z_c2_$v.err; echo "c2 $v rc=$?"
  CLV_LIBRARY=$lib timeout 300 python tools/sweep_walk.py --config c1 --passes 16 --settings "mode=0;mode=0,connect_schedule=4;mode=0,connect_schedule=8" > gpurun_out/r02z_c1_$v.jsonl 2> gpurun_out/r02z_c1_$v.err; echo "c1 $v rc=$?"
done
timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0;mode=0,walk_tail_budget=0;mode=0,walk_tail_budget=32;mode=0,walk_tail_budget=64,slices=2;mode=0,slices=1" > gpurun_out/r02z_c4.jsonl 2> gpurun_out/r02z_c4.err; echo "c4 rc=$?"
