#!/bin/bash
# round 2, GPU call F: validation at HEAD — all GPU tests, default bench (c4) + c1..c3, reference arm, ncu traffic + launch list + full set
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log; tail -2 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench_c4.json 2> gpurun_out/r02f_bench_c4.err; echo "bench c4 rc=$?"
for c in c3 c2 c1; do python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r02f_bench_$c.json 2> gpurun_out/r02f_bench_$c.err; echo "bench $c rc=$?"; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_c4_ref.json 2> gpurun_out/r02f_bench_c4_ref.err; echo "ref rc=$?"
export CLV_PASS_SLICES=1
for c in c4 c3 c2; do
  n=29; if [ $c = c2 ]; then n=5; fi
  python tools/profile_pass.py --config $c --passes 3 > gpurun_out/r02f_plain_$c.log 2>&1 &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_walk|k_fused|k_connect' -s $((2*n)) -c $n --csv --log-file gpurun_out/r02f_traffic_$c.csv \
      python tools/profile_pass.py --config $c --passes 3 > gpurun_out/r02f_ncu_traffic_$c.log 2>&1; echo "ncu traffic $c rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:'k_walk_trace|k_walk_shade|k_connect' -s 27 -c 27 -o gpurun_out/r02f_c4_full \
    python tools/profile_pass.py --config c4 --passes 2 > gpurun_out/r02f_ncu_full.log 2>&1; echo "ncu full rc=$?"
unset CLV_PASS_SLICES
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_launches.log 2>&1; echo "launch list rc=$?"
