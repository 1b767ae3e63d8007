#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_batch_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/r02_22_pytest.log 2>&1
tail -4 gpurun_out/r02_22_pytest.log
show() {
python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_22_%s.json'%n).read().strip().split('\n')[-1])
    k=d['config']['rank0_kernels_us_per_step']
    print(n, 'value %.1f GVox/s step %.1f us host %.1f us e2e %.1f' % (d['value'], 1000*d['ms_per_step'], d['config'].get('host_enqueue_us_per_step',0), d['e2e']['value']), 'hiz %.1f cand %.1f bricks %.1f span %.1f' % (k['hiz'],k['candidates'],k['bricks'],k['bricks_first_cta_to_last_cta']), 'frac %.3f' % d['roofline']['frac'], 'parity', d['parity_check'].get('counters_equal'), d['parity_check'].get('state_bit_exact'), d['config']['timing'][-60:])
    print('   timeline', d['config']['rank0_device_timeline_us'])
except Exception as e: print(n, 'parse failed', e); print(open('gpurun_out/r02_22_%s.err'%n).read()[-1500:])
PY
}
timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_22_n1.json 2> gpurun_out/r02_22_n1.err; show n1
