for flags in "-DCLV_WIDE_CLOSEST=0 -DCLV_WIDE_OCCLUSION=0" "-DCLV_WIDE_CLOSEST=0 -DCLV_WIDE_OCCLUSION=1" "-DCLV_WIDE_CLOSEST=1 -DCLV_WIDE_OCCLUSION=1"; do
  CLV_NVCC_EXTRA="$flags" python -m clive2_b200.build --force > /dev/null 2>&1
  for cfg in c2 c3; do
  python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$flags', '$cfg', round(d['value'],2), round(d['ms_per_step'],3), {k[2:]:round(v,3) for k,v in d['roofline']['kernel_ms_per_pass'].items()})"
  done
done
python -m clive2_b200.build --force > /dev/null 2>&1
