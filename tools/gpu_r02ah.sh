#!/bin/bash
# new edge-case test (15-triangle leaves through the large-scene flavours) + the whole configs file
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r02ah_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02ah_pytest.log
