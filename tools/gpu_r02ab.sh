#!/bin/bash
# walk state as one 96-byte record per sample (RNG words inside): parity, timing, DRAM bytes of the shade rounds
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02ab_pytest.log
timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0;mode=0,walk_tail_budget=0;mode=0,walk_tail_budget=32;mode=0,walk_tail_budget=64" > gpurun_out/r02ab_c4.jsonl 2> gpurun_out/r02ab_c4.err; echo "c4 rc=$?"
timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0;mode=0,walk_tail_budget=0;mode=0,walk_tail_budget=32;mode=0,walk_tail_budget=64" > gpurun_out/r02ab_c3.jsonl 2> gpurun_out/r02ab_c3.err; echo "c3 rc=$?"
export CLV_PASS_SLICES=1
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes.sum,lts__t_bytes.sum
python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02ab_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_walk_shade|k_walk_init' -s 34 -c 17 --csv --log-file gpurun_out/r02ab_shade_b64.csv python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02ab_ncu.log 2>&1; echo "ncu rc=$?"
