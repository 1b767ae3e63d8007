#!/bin/bash
mkdir -p gpurun_out
timeout 100 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_36_n1.json 2> gpurun_out/r02_36_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_36_n1.json').read().strip().split('\n')[-1])
print('value %.1f e2e %.2f' % (d['value'], d['e2e']['value']), 'e2e_depth_mm', d['e2e_depth_mm'])
PY
tail -3 gpurun_out/r02_36_n1.err
