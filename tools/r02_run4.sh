#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_4_smi.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 800 -s > gpurun_out/r02_4_multi.log 2>&1
echo "multi rc=$?" >> gpurun_out/r02_4_multi.log
tail -12 gpurun_out/r02_4_multi.log
timeout 900 python -m pytest tests -m gpu -q --timeout 800 -x --deselect tests/test_multi_gpu.py > gpurun_out/r02_4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_4_pytest.log
tail -4 gpurun_out/r02_4_pytest.log
ab() {
  local n=$1 lib=$2; shift 2
  ( if [ -n "$lib" ]; then export CHS_LIB_PATH=$PWD/cvids_b200/_ab_$lib.so; fi
    env "$@" timeout 600 python bench.py --no-cpu --quick --no-side-lines --parity-steps 1 2>gpurun_out/r02_4_$n.err | tail -1 > gpurun_out/r02_4_$n.json
    python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_4_%s.json'%n).read()); r=d['roofline']
    print(n, 'step %.1f us' % (1000*d['ms_per_step']), 'prep %.1f cand %.1f wait %.1f bricks %.1f span %.1f' % (1000*r['prepare_ms_per_launch'], 1000*r['candidates_ms_per_launch'], 1000*r['new_chunks_ms_per_launch'], 1000*r['bricks_ms_per_launch'], 1000*r.get('bricks_span_ms_per_launch',0)), 'GVox/s %.1f' % d['value'], 'frac %.3f' % r['frac'], 'parity', (d.get('parity_check') or {}).get('state_bit_exact'))
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/r02_4_%s.err'%n).read()[-2000:])
PY
  )
}
ab base "" X=1
ab base2 "" X=1
