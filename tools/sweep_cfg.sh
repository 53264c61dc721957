#!/bin/bash
# usage: CFG=c3 STEPS=10 tools/sweep_cfg.sh "VAR=val" ...  -> one bench line per setting for the given configuration
for setting in "$@"; do
  out=$(env $setting python bench.py --config ${CFG:-c2} --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$setting" "$out" <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    k = d["roofline"]["kernel_ms_per_pass"]
    print(f"{sys.argv[1][-40:]:40s} value={d['value']:.1f} ms/step={d['ms_per_step']:.3f} kernels={ {a[2:]: round(b, 3) for a, b in k.items()} } e2e={d['e2e']['value']:.1f}")
except Exception as e:
    print(sys.argv[1], "FAILED", sys.argv[2][-300:])
PY
done
