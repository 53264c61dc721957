#!/bin/bash
# round 2, GPU call A: all GPU tests, bench at HEAD on c4 / c2 / c3, reference arm, ncu of one c4 pass
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -3 gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_c4.json 2> gpurun_out/r02a_bench_c4.err; echo "bench c4 rc=$?"
python bench.py --config c2 --steps 20 --warmup 3 > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --config c3 --steps 20 --warmup 3 > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err; echo "bench c3 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_c4_ref.json 2> gpurun_out/r02a_bench_c4_ref.err; echo "ref rc=$?"
export CLV_PASS_SLICES=1
python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_fused|k_connect' -s 10 -c 5 -o gpurun_out/r02a_c4 \
    python tools/profile_pass.py --config c4 --passes 3 > gpurun_out/r02a_ncu.log 2>&1; echo "ncu c4 rc=$?"
python tools/profile_pass.py --config c2 --passes 3 > gpurun_out/r02a_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_fused|k_connect' -s 10 -c 5 -o gpurun_out/r02a_c2 \
    python tools/profile_pass.py --config c2 --passes 3 > gpurun_out/r02a_ncu2.log 2>&1; echo "ncu c2 rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_gpu.txt
nproc >> gpurun_out/r02a_gpu.txt
