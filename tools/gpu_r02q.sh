#!/bin/bash
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1 CLV_PASS_SLICES=1 CLV_WALK_MODE=2
python tools/profile_pass.py --config c2 --passes 3 > gpurun_out/r02q_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_walk_trace|k_walk_shade' -s 48 -c 6 -o gpurun_out/r02q_c2_walk \
    python tools/profile_pass.py --config c2 --passes 3 > gpurun_out/r02q_ncu.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_walk' -s 50 -c 25 --csv --log-file gpurun_out/r02q_c2_rounds.csv python tools/profile_pass.py --config c2 --passes 3 > gpurun_out/r02q_ncu2.log 2>&1; echo "rounds rc=$?"
du -sh gpurun_out
