#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_16_n1.json 2> gpurun_out/r02_16_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_16_n1.json').read().strip().split('\n')[-1])
print('value %.1f step %.1f us' % (d['value'], 1000*d['ms_per_step']), d['config']['timing'][-70:])
print(d['config']['rank0_device_timeline_us'])
PY
