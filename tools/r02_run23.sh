#!/bin/bash
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_23_%s.json'%n).read().strip().split('\n')[-1])
    k=d['config']['rank0_kernels_us_per_step']
    t=d['config']['rank0_device_timeline_us']
    print(n, 'value %.1f GVox/s step %.1f us' % (d['value'], 1000*d['ms_per_step']), 'events: hiz %.1f cand %.1f bricks %.1f' % (k['hiz'],k['candidates'],k['bricks']), 'timeline: step %.1f hiz %.1f cand %.1f bricks %.1f' % (t['step'],t['hiz'],t['candidates'],t['bricks']), 'parity', d['parity_check'].get('counters_equal'), d['parity_check'].get('state_bit_exact'))
except Exception as e: print(n, 'parse failed', e); print(open('gpurun_out/r02_23_%s.err'%n).read()[-1500:])
PY
}
for v in c8 c6 c5; do
  export CHS_LIB_PATH=$PWD/cvids_b200/_ab_$v.so
  timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_23_$v.json 2> gpurun_out/r02_23_$v.err; show $v
done
