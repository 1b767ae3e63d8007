#!/bin/bash
# final single-GPU verification of round 2: whole GPU suite, smoke, the default bench line (timed), the reference arm, ncu evidence
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_32_pytest.log 2>&1; echo "pytest rc=$? after $(( $(date +%s) - S )) s" | tee -a gpurun_out/r02_32_pytest.log; tail -3 gpurun_out/r02_32_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r02_32_smoke.log 2>&1; tail -1 gpurun_out/r02_32_smoke.log
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02_32_bench_n1.json 2> gpurun_out/r02_32_bench_n1.err; echo "bench rc=$? after $(( $(date +%s) - S )) s"
python - <<'PY'
import json
for ln in open('gpurun_out/r02_32_bench_n1.json').read().strip().split('\n'):
    try: d=json.loads(ln)
    except Exception: continue
    print(d.get('metric'), 'value %.2f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), 'roofline', d.get('roofline',{}).get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('config',{}).get('workload','')[:60])
PY
S=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r02_32_bench_ref.json 2> gpurun_out/r02_32_bench_ref.err; echo "reference arm rc=$? after $(( $(date +%s) - S )) s"; tail -c 600 gpurun_out/r02_32_bench_ref.json
B="python bench.py --steps 6 --warmup 3 --passes 1 --no-cpu --quick --no-side-lines --parity-steps 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"batch_candidates|batch_bricks_fast|batch_hiz_tma" --launch-skip 30 --launch-count 3 -o gpurun_out/r02_32_kernels $B > gpurun_out/r02_32_ncu.log 2>&1
tail -2 gpurun_out/r02_32_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --launch-skip 40 --csv --log-file gpurun_out/r02_32_launches.csv $B > /dev/null 2>&1
tail -3 gpurun_out/r02_32_launches.csv | cut -c1-160
