#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/r02_30_nproc.txt; taskset -p $$ >> gpurun_out/r02_30_nproc.txt 2>&1; cat gpurun_out/r02_30_nproc.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 500 -x -s > gpurun_out/r02_30_mgpu.log 2>&1; tail -3 gpurun_out/r02_30_mgpu.log
show() {
python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_30_%s.json'%n).read().strip().split('\n')[-1])
    k=d['config']['rank0_kernels_us_per_step']
    print(n, 'value %.1f GVox/s step %.1f us host %.1f us e2e %.1f' % (d['value'], 1000*d['ms_per_step'], d['config'].get('host_enqueue_us_per_step',0), d['e2e']['value']), 'hiz %.1f cand %.1f bricks %.1f span %.1f' % (k['hiz'],k['candidates'],k['bricks'],k['bricks_first_cta_to_last_cta']), 'parity', d['parity_check'].get('counters_equal'), d['parity_check'].get('state_bit_exact'), d['config']['timing'][-60:])
    print('   timeline', d['config']['rank0_device_timeline_us'])
except Exception as e: print(n, 'parse failed', e); print(open('gpurun_out/r02_30_%s.err'%n).read()[-1500:])
PY
}
export CHS_HOST_PROFILE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu --quick > gpurun_out/r02_30_n2.json 2> gpurun_out/r02_30_n2.err; show n2; grep -A1 "host profile" gpurun_out/r02_30_n2.err | head -2
CHS_NO_SHARD_HIZ=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-cpu --quick > gpurun_out/r02_30_n2ns.json 2> gpurun_out/r02_30_n2ns.err; show n2ns
