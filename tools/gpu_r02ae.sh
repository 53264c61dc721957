#!/bin/bash
# leaf_deal by ballots (no scan shuffles): parity (incl. the new tile-shard test), timing
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02ae_pytest.log
timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0" > gpurun_out/r02ae_c4.jsonl 2> gpurun_out/r02ae_c4.err; echo "c4 rc=$?"
timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0" > gpurun_out/r02ae_c3.jsonl 2> gpurun_out/r02ae_c3.err; echo "c3 rc=$?"
