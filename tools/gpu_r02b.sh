#!/bin/bash
# round 2, GPU call B: wavefront walks — parity tests, megakernel vs wavefront per configuration, ncu of a wavefront pass
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -3 gpurun_out/r02b_pytest.log
python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=1;mode=2;mode=2,walk_refill_below=16;mode=2,walk_refill_below=32;mode=2,walk_schedule=4;mode=2,walk_schedule=16;mode=2,slices=1;mode=2,slices=2;mode=2,slices=6" > gpurun_out/r02b_sweep_c4.jsonl 2> gpurun_out/r02b_sweep_c4.err; echo "sweep c4 rc=$?"
python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=1;mode=2;mode=2,walk_refill_below=16;mode=2,walk_refill_below=32" > gpurun_out/r02b_sweep_c3.jsonl 2> gpurun_out/r02b_sweep_c3.err; echo "sweep c3 rc=$?"
python tools/sweep_walk.py --config c2 --passes 20 --settings "mode=1;mode=2;mode=2,walk_refill_below=16;mode=2,walk_refill_below=32;mode=2,slices=1" > gpurun_out/r02b_sweep_c2.jsonl 2> gpurun_out/r02b_sweep_c2.err; echo "sweep c2 rc=$?"
python tools/sweep_walk.py --config c1 --passes 20 --settings "mode=1;mode=2" > gpurun_out/r02b_sweep_c1.jsonl 2> gpurun_out/r02b_sweep_c1.err; echo "sweep c1 rc=$?"
export CLV_PASS_SLICES=1
python tools/profile_pass.py --config c4 --passes 2 > gpurun_out/r02b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_walk' -s 25 -c 9 -o gpurun_out/r02b_c4_walk \
    python tools/profile_pass.py --config c4 --passes 2 > gpurun_out/r02b_ncu.log 2>&1; echo "ncu c4 rc=$?"
