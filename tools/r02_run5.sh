#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu --no-side-lines --steps2 8 > gpurun_out/r02_5_n1.json 2> gpurun_out/r02_5_n1.err
echo "n1 rc=$?"; tail -c 3000 gpurun_out/r02_5_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_5_n1.json').read().strip().split('\n')[-1])
    print('N=1 value %.1f GVox/s step %.1f us e2e %.1f' % (d['value'], 1000*d['ms_per_step'], d['e2e']['value']), d['config']['rank0_kernels_us_per_step'], 'frac %.3f' % d['roofline']['frac'], d['parity_check'], d['mesh'])
    s=d.get('single_agent_color',{})
    print('single', {k:s.get(k) for k in ('value','ms_per_step','l2_flushed_per_step','kernels_us_per_step_cold_l2','parity_check')}, s.get('roofline',{}).get('frac'), s.get('e2e'), s.get('error'))
except Exception as e: print('parse failed', e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_5_n2.json 2> gpurun_out/r02_5_n2.err
echo "n2 rc=$?"; tail -c 2000 gpurun_out/r02_5_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_5_n2.json').read().strip().split('\n')[-1])
    print('N=2 value %.1f GVox/s step %.1f us e2e %.1f' % (d['value'], 1000*d['ms_per_step'], d['e2e']['value']), d['config']['rank0_kernels_us_per_step'], d['parity_check'], d['mesh'], d['config']['timing'])
except Exception as e: print('parse failed', e)
PY
