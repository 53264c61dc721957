#!/bin/bash
# round 2, GPU call F2: benches at HEAD, traffic, ncu (small), cooperative finish A/B + tests under it
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0;mode=0,finish_cooperative=1" > gpurun_out/r02f_coop_c4.jsonl 2> gpurun_out/r02f_coop_c4.err; echo "coop c4 rc=$?"
python tools/sweep_walk.py --config c2 --passes 20 --settings "mode=0;mode=0,finish_cooperative=1" > gpurun_out/r02f_coop_c2.jsonl 2> gpurun_out/r02f_coop_c2.err; echo "coop c2 rc=$?"
CLV_FINISH_COOPERATIVE=1 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_coop.log 2>&1; echo "pytest(coop) rc=$?" >> gpurun_out/r02f_pytest_coop.log; tail -2 gpurun_out/r02f_pytest_coop.log
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
ncu --set full --clock-control none --import-source on -k regex:'k_walk_trace|k_walk_shade' -s 24 -c 4 -o gpurun_out/r02f_c4_walk \
    python tools/profile_pass.py --config c4 --passes 2 > gpurun_out/r02f_ncu_full1.log 2>&1; echo "ncu full walk rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_connect' -s 3 -c 3 -o gpurun_out/r02f_c4_connect \
    python tools/profile_pass.py --config c4 --passes 2 > gpurun_out/r02f_ncu_full2.log 2>&1; echo "ncu full connect rc=$?"
unset CLV_PASS_SLICES
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_launches.log 2>&1; echo "launch list rc=$?"
du -sh gpurun_out
