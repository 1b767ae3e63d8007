#!/bin/bash
# final check of the shipped build on 8 GPUs: the multi-GPU parity test and one N = 8 bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 350 -x -s > gpurun_out/r02_35_mgpu_n8.log 2>&1; tail -3 gpurun_out/r02_35_mgpu_n8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --no-cpu > gpurun_out/r02_35_n8.json 2> gpurun_out/r02_35_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_35_n8.json').read().strip().split('\n')[-1])
print('n8 value %.1f GVox/s step %.1f us host %.1f e2e %.1f' % (d['value'], 1000*d['ms_per_step'], d['config']['host_enqueue_us_per_step'], d['e2e']['value']), d['parity_check'].get('counters_equal'), 'mesh %.2f ms' % d['mesh']['wall_ms'], d['config']['timing'][-60:])
print(d['config']['rank0_device_timeline_us'])
PY
