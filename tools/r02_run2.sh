#!/bin/bash
# round 2, GPU call 1: full GPU test suite, A/B of the brick kernel variants, one ncu capture of the fast brick kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_2_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_2_pytest.log
tail -5 gpurun_out/r02_2_pytest.log
ab() {  # name, lib, extra env
  local n=$1 lib=$2; shift 2
  ( if [ -n "$lib" ]; then export CHS_LIB_PATH=$PWD/cvids_b200/_ab_$lib.so; fi
    env "$@" timeout 600 python bench.py --no-cpu --quick --no-side-lines --parity-steps 1 2>gpurun_out/r02_2_$n.err | tail -1 > gpurun_out/r02_2_$n.json
    python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_2_%s.json'%n).read()); r=d['roofline']
    print(n, 'step %.1f us' % (1000*d['ms_per_step']), 'prep %.1f cand %.1f bricks %.1f' % (1000*r['prepare_ms_per_launch'], 1000*r['candidates_ms_per_launch'], 1000*r['bricks_ms_per_launch']), 'GVox/s %.1f' % d['value'], 'frac %.3f' % r['frac'], 'parity', d.get('parity_check'))
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/r02_2_%s.err'%n).read()[-2000:])
PY
  )
}
ab base "" X=1
ab nofast "" CHS_NO_FAST_BRICKS=1
ab t128 t128 X=1
ab t128c5 t128c5 X=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batch_bricks_fast --launch-skip 5 --launch-count 1 -o gpurun_out/r02_2_fast python bench.py --steps 4 --warmup 3 --no-cpu --quick --no-side-lines --parity-steps 0 > gpurun_out/r02_2_ncu.log 2>&1
tail -3 gpurun_out/r02_2_ncu.log
