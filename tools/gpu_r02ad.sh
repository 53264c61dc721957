#!/bin/bash
# triangles of the cooperative leaf steps through the texture pipe: parity, then A/B (CLV_NO_TRIS_TEX=1 = plain loads)
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q -k "suspend or c4_terrain or c3_terrain or wavefront" > gpurun_out/r02ad_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02ad_pytest.log
for t in tex plain; do
  if [ $t = plain ]; then export CLV_NO_TRIS_TEX=1; fi
  timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0" > gpurun_out/r02ad_c4_$t.jsonl 2> gpurun_out/r02ad_c4_$t.err; echo "c4 $t rc=$?"
  timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0" > gpurun_out/r02ad_c3_$t.jsonl 2> gpurun_out/r02ad_c3_$t.err; echo "c3 $t rc=$?"
done
