#!/bin/bash
mkdir -p gpurun_out
for mode in off nvml smi; do
CHS_BENCH_SAMPLER=$mode timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_17_$mode.json 2> gpurun_out/r02_17_$mode.err
python - $mode <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02_17_%s.json'%sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], 'value %.1f step %.1f us host %.1f' % (d['value'], 1000*d['ms_per_step'], d['config']['host_enqueue_us_per_step']), d['config']['timing'][-70:], d['clocks'])
print('  ', d['config']['rank0_device_timeline_us'])
PY
done
