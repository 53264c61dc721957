#!/bin/bash
# after moving the cooperative leaf steps behind a warp policy (same SASS up to scheduling): all GPU tests + timing
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02aj_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02aj_pytest.log
timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0" > gpurun_out/r02aj_c4.jsonl 2> gpurun_out/r02aj_c4.err; echo "c4 rc=$?"
