#!/bin/bash
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ak_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q -k "c3_terrain or suspend or 15_triangles" > gpurun_out/r02ak_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02ak_pytest.log
