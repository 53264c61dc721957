#!/bin/bash
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02p_pytest.log; tail -2 gpurun_out/r02p_pytest.log
timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0" > gpurun_out/r02p_sweep_c4.jsonl 2> gpurun_out/r02p_sweep_c4.err; echo "c4 rc=$?"
timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0" > gpurun_out/r02p_sweep_c3.jsonl 2> gpurun_out/r02p_sweep_c3.err; echo "c3 rc=$?"
timeout 300 python tools/sweep_walk.py --config c2 --passes 20 --settings "mode=0" > gpurun_out/r02p_sweep_c2.jsonl 2> gpurun_out/r02p_sweep_c2.err; echo "c2 rc=$?"
timeout 300 python tools/sweep_walk.py --config c1 --passes 20 --settings "mode=0" > gpurun_out/r02p_sweep_c1.jsonl 2> gpurun_out/r02p_sweep_c1.err; echo "c1 rc=$?"
