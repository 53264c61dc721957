#!/bin/bash
# C2 (or $CFG) throughput against the number of slices a pass is split into (CLV_PASS_SLICES)
CFG=${CFG:-c2}
for s in ${SLICES:-1 2 3 4 6 8}; do
  CLV_PASS_SLICES=$s python bench.py --config $CFG --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('slices=$s', 'value=%.1f'%d['value'], 'ms=%.3f'%d['ms_per_step'], 'e2e=%.1f'%d['e2e']['value'], 'kernels_sum=%.3f'%sum(d['roofline']['kernel_ms_per_pass'].values()))"
done
