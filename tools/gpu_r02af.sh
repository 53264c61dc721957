#!/bin/bash
# part-major copy of the triangle records for the cooperative leaf steps: parity, then A/B (CLV_NO_TRIS_PARTS=1 = 48-byte records)
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q -k "suspend or c4_terrain or c3_terrain or tile_shards" > gpurun_out/r02af_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02af_pytest.log
for t in parts plain; do
  if [ $t = plain ]; then export CLV_NO_TRIS_PARTS=1; fi
  timeout 300 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0" > gpurun_out/r02af_c4_$t.jsonl 2> gpurun_out/r02af_c4_$t.err; echo "c4 $t rc=$?"
  timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0" > gpurun_out/r02af_c3_$t.jsonl 2> gpurun_out/r02af_c3_$t.err; echo "c3 $t rc=$?"
done
