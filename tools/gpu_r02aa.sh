#!/bin/bash
# why does k_walk_shade take longer when searches are suspended?  instruction and DRAM counts per launch, budget 0 vs 16
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1 CLV_PASS_SLICES=1
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes.sum,lts__t_bytes.sum,smsp__thread_inst_executed.sum
CLV_WALK_TAIL_BUDGET=0 python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02aa_plain.log 2>&1 || exit 1
CLV_WALK_TAIL_BUDGET=0 timeout 600 ncu --metrics $M --clock-control none -k regex:'k_walk_shade|k_walk_init' -s 26 -c 13 --csv --log-file gpurun_out/r02aa_shade_b0.csv python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02aa_ncu_b0.log 2>&1; echo "b0 rc=$?"
CLV_WALK_TAIL_BUDGET=16 timeout 600 ncu --metrics $M --clock-control none -k regex:'k_walk_shade|k_walk_init' -s 34 -c 17 --csv --log-file gpurun_out/r02aa_shade_b16.csv python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02aa_ncu_b16.log 2>&1; echo "b16 rc=$?"
