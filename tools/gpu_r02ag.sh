#!/bin/bash
# final state of round 2: all GPU tests, smoke, default bench (driver's flags), ncu DRAM traffic per pass at C4
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02ag_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ag_pytest.log; tail -2 gpurun_out/r02ag_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ag_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02ag_bench_c4.json 2> gpurun_out/r02ag_bench_c4.err; echo "bench c4 rc=$?"
timeout 600 python bench.py --config c3 --steps 20 --warmup 5 > gpurun_out/r02ag_bench_c3.json 2> gpurun_out/r02ag_bench_c3.err; echo "bench c3 rc=$?"
export CLV_PASS_SLICES=1
timeout 300 python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02ag_plain_c4.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_walk|k_fused|k_connect' -s 74 -c 37 --csv --log-file gpurun_out/r02ag_traffic_c4.csv \
    python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02ag_ncu_traffic_c4.log 2>&1; echo "ncu traffic c4 rc=$?"
