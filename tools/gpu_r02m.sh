#!/bin/bash
# round 2, GPU call M: final single-GPU validation at HEAD
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02m_pytest.log; tail -3 gpurun_out/r02m_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02m_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench_c4.json 2> gpurun_out/r02m_bench_c4.err; echo "bench c4 rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02m_bench_c4_ref.json 2> gpurun_out/r02m_bench_c4_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02m_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_ncu_launches.log 2>&1; echo "launch list rc=$?"
du -sh gpurun_out
