#!/bin/bash
# 8-GPU box: the multi-GPU parity test at world 8, then the scaling series 8, 4, 2, 1 as the driver runs it
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_28_topo.txt 2>&1
CHS_TEST_WORLD=8 timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 800 -x -s > gpurun_out/r02_28_mgpu_n8.log 2>&1
tail -3 gpurun_out/r02_28_mgpu_n8.log; grep -h "MULTI_GPU" gpurun_out/r02_28_mgpu_n8.log | head
show() {
python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_28_%s.json'%n).read().strip().split('\n')[-1])
    k=d['config']['rank0_kernels_us_per_step']
    print(n, 'value %.1f GVox/s step %.1f us host %.1f us e2e %.1f' % (d['value'], 1000*d['ms_per_step'], d['config'].get('host_enqueue_us_per_step',0), d['e2e']['value']), 'hiz %.1f cand %.1f bricks %.1f span %.1f' % (k['hiz'],k['candidates'],k['bricks'],k['bricks_first_cta_to_last_cta']), 'parity', d['parity_check'].get('counters_equal'), d['parity_check'].get('state_bit_exact'), 'mesh wall %.2f ms' % d['mesh']['wall_ms'], d['config']['timing'][-60:])
    print('   timeline', d['config']['rank0_device_timeline_us'])
except Exception as e: print(n, 'parse failed', e); print(open('gpurun_out/r02_28_%s.err'%n).read()[-1500:])
PY
}
export CHS_HOST_PROFILE=1
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu > gpurun_out/r02_28_n$n.json 2> gpurun_out/r02_28_n$n.err; show n$n; grep -A1 "host profile" gpurun_out/r02_28_n$n.err | head -2
done
timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_28_n1.json 2> gpurun_out/r02_28_n1.err; show n1
