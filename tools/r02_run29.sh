#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_batch_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r02_29_pytest.log 2>&1
tail -3 gpurun_out/r02_29_pytest.log
export CHS_HOST_PROFILE=1
timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_29_n1.json 2> gpurun_out/r02_29_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_29_n1.json').read().strip().split('\n')[-1])
print('n1 value %.1f step %.1f us host %.1f' % (d['value'], 1000*d['ms_per_step'], d['config']['host_enqueue_us_per_step']), d['parity_check'].get('counters_equal'), d['parity_check'].get('state_bit_exact'))
print(d['config']['rank0_device_timeline_us'])
PY
grep -A1 "host profile" gpurun_out/r02_29_n1.err
