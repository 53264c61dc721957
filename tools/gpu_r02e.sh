#!/bin/bash
# round 2, GPU call E: adopted wavefront flavours; dynamic fetch in k_connect_rays; trace grid divisor x slices; tests
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=2;mode=2,connect_dynamic_fetch=1;mode=2,connect_dynamic_fetch=1,connect_refill_below=24;mode=2,connect_dynamic_fetch=1,walk_grid_divisor=2,slices=3;mode=2,connect_dynamic_fetch=1,walk_grid_divisor=3,slices=3;mode=2,connect_dynamic_fetch=1,walk_grid_divisor=2,slices=4;mode=2,connect_dynamic_fetch=1,walk_grid_divisor=1,slices=4;mode=2,connect_dynamic_fetch=1,walk_grid_divisor=1,slices=2;mode=1,connect_dynamic_fetch=0,slices=0" > gpurun_out/r02e_sweep_c4.jsonl 2> gpurun_out/r02e_sweep_c4.err; echo "c4 rc=$?"
python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=2;mode=2,connect_dynamic_fetch=1;mode=2,connect_dynamic_fetch=1,walk_grid_divisor=2;mode=1,connect_dynamic_fetch=1;mode=1,connect_dynamic_fetch=0" > gpurun_out/r02e_sweep_c3.jsonl 2> gpurun_out/r02e_sweep_c3.err; echo "c3 rc=$?"
python tools/sweep_walk.py --config c2 --passes 20 --settings "mode=1;mode=1,connect_dynamic_fetch=1;mode=1,connect_dynamic_fetch=1,connect_refill_below=24;mode=2,connect_dynamic_fetch=1" > gpurun_out/r02e_sweep_c2.jsonl 2> gpurun_out/r02e_sweep_c2.err; echo "c2 rc=$?"
python tools/sweep_walk.py --config c1 --passes 20 --settings "mode=1;mode=1,connect_dynamic_fetch=1" > gpurun_out/r02e_sweep_c1.jsonl 2> gpurun_out/r02e_sweep_c1.err; echo "c1 rc=$?"
CLV_CONNECT_DYNAMIC=1 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_dyn.log 2>&1; echo "pytest(dyn) rc=$?" >> gpurun_out/r02e_pytest_dyn.log; tail -2 gpurun_out/r02e_pytest_dyn.log
