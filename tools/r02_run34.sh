#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 500 -x -s > gpurun_out/r02_34_mgpu.log 2>&1; tail -3 gpurun_out/r02_34_mgpu.log
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r02_34_n2.json 2> gpurun_out/r02_34_n2.err; echo "bench rc=$? after $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_34_n2.json').read().strip().split('\n')[-1])
print('n2 value %.1f GVox/s step %.1f us host %.1f e2e %.1f' % (d['value'], 1000*d['ms_per_step'], d['config']['host_enqueue_us_per_step'], d['e2e']['value']), d['parity_check'].get('counters_equal'), 'mesh %.2f ms' % d['mesh']['wall_ms'], d['config']['timing'][-60:])
print(d['config']['rank0_device_timeline_us'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r02_34_ref_n2.json 2> gpurun_out/r02_34_ref_n2.err; echo "ref arm rc=$?"; tail -c 300 gpurun_out/r02_34_ref_n2.json
