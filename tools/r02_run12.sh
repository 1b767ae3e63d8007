#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_batch_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r02_12_pytest.log 2>&1
tail -4 gpurun_out/r02_12_pytest.log
show() {
python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_12_%s.json'%n).read().strip().split('\n')[-1])
    k=d['config']['rank0_kernels_us_per_step']
    print(n, 'value %.1f GVox/s step %.1f us host %.1f us' % (d['value'], 1000*d['ms_per_step'], d['config'].get('host_enqueue_us_per_step',0)), 'hiz %.1f cand %.1f bricks %.1f span %.1f' % (k['hiz'],k['candidates'],k['bricks'],k['bricks_first_cta_to_last_cta']), 'frac %.3f' % d['roofline']['frac'], 'parity', d['parity_check'].get('counters_equal'), d['parity_check'].get('state_bit_exact'), d['config']['timing'][-60:])
except Exception as e: print(n, 'parse failed', e); print(open('gpurun_out/r02_12_%s.err'%n).read()[-1500:])
PY
}
for v in main v8 v8t384 v8t128c3 v8t128c4; do
  if [ $v = main ]; then unset CHS_LIB_PATH; else export CHS_LIB_PATH=$PWD/cvids_b200/_ab_$v.so; fi
  timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_12_$v.json 2> gpurun_out/r02_12_$v.err; show $v
done
