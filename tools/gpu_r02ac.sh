#!/bin/bash
# round 2, validation at HEAD after the cooperative leaves: all GPU tests, smoke, benches c4 (default) + c1..c3, reference arm,
# ncu DRAM traffic per pass (c4 / c3 / c2), ncu full set of the two ray kernels, launch list
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_pytest.log; tail -2 gpurun_out/r02ac_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ac_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ac_bench_c4.json 2> gpurun_out/r02ac_bench_c4.err; echo "bench c4 rc=$?"
for c in c3 c2 c1; do timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r02ac_bench_$c.json 2> gpurun_out/r02ac_bench_$c.err; echo "bench $c rc=$?"; done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ac_bench_c4_ref.json 2> gpurun_out/r02ac_bench_c4_ref.err; echo "ref rc=$?"
export CLV_PASS_SLICES=1
for c in c4 c3 c2; do
  n=37; if [ $c = c2 ]; then n=5; fi
  timeout 300 python tools/profile_pass.py --config $c --passes 3 > gpurun_out/r02ac_plain_$c.log 2>&1 &&
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_walk|k_fused|k_connect' -s $((2*n)) -c $n --csv --log-file gpurun_out/r02ac_traffic_$c.csv \
      python tools/profile_pass.py --config $c --passes 3 > gpurun_out/r02ac_ncu_traffic_$c.log 2>&1; echo "ncu traffic $c rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_walk_trace' -s 16 -c 3 -o gpurun_out/r02ac_c4_trace \
    python tools/profile_pass.py --config c4 --passes 2 > gpurun_out/r02ac_ncu_full1.log 2>&1; echo "ncu full trace rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_connect_rays' -s 1 -c 1 -o gpurun_out/r02ac_c4_rays \
    python tools/profile_pass.py --config c4 --passes 2 > gpurun_out/r02ac_ncu_full2.log 2>&1; echo "ncu full rays rc=$?"
unset CLV_PASS_SLICES
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02ac_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02ac_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02ac_ncu_launches.log 2>&1; echo "launch list rc=$?"
du -sh gpurun_out
