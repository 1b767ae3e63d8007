#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_batch_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_facade.py -m gpu -q --timeout 900 > gpurun_out/r02_33_pytest.log 2>&1; tail -3 gpurun_out/r02_33_pytest.log
S=$(date +%s)
timeout 900 python bench.py --no-cpu > gpurun_out/r02_33_bench_n1.json 2> gpurun_out/r02_33_bench_n1.err; echo "bench rc=$? after $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_33_bench_n1.json').read().strip().split('\n')[-1])
print('value %.2f step %.1f us roofline %.3f traffic %s' % (d['value'], 1000*d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic']))
sa=d['single_agent_color']
print('single agent value %.2f step %.1f us' % (sa['value'], 1000*sa['ms_per_step']), sa['kernels_us_per_step_cold_l2'], 'frac %.3f' % sa['roofline']['frac'], 'e2e %.2f' % sa['e2e']['value'], 'flushed', sa['l2_flushed_per_step'])
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"7quarter.*batch_bricks_fast" --launch-skip 6 --launch-count 1 -o gpurun_out/r02_33_color python bench.py --steps 4 --warmup 3 --steps2 6 --passes 1 --no-cpu --no-side-lines --parity-steps 0 > gpurun_out/r02_33_ncu.log 2>&1
tail -2 gpurun_out/r02_33_ncu.log
