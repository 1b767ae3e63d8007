#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_3_pytest.log
tail -4 gpurun_out/r02_3_pytest.log
ab() {
  local n=$1 lib=$2; shift 2
  ( if [ -n "$lib" ]; then export CHS_LIB_PATH=$PWD/cvids_b200/_ab_$lib.so; fi
    env "$@" timeout 600 python bench.py --no-cpu --quick --no-side-lines --parity-steps 1 2>gpurun_out/r02_3_$n.err | tail -1 > gpurun_out/r02_3_$n.json
    python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02_3_%s.json'%n).read()); r=d['roofline']
    print(n, 'step %.1f us' % (1000*d['ms_per_step']), 'prep %.1f cand %.1f wait %.1f bricks %.1f span %.1f' % (1000*r['prepare_ms_per_launch'], 1000*r['candidates_ms_per_launch'], 1000*r['new_chunks_ms_per_launch'], 1000*r['bricks_ms_per_launch'], 1000*r.get('bricks_span_ms_per_launch',0)), 'GVox/s %.1f' % d['value'], 'frac %.3f' % r['frac'], 'parity', (d.get('parity_check') or {}).get('state_bit_exact'))
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/r02_3_%s.err'%n).read()[-2000:])
PY
  )
}
ab base "" X=1
ab base2 "" X=1
ab notma "" CHS_NO_TMA_HIZ=1
ab nofast "" CHS_NO_FAST_BRICKS=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"batch_hiz_tma|batch_candidates" --launch-skip 10 --launch-count 2 -o gpurun_out/r02_3_hc python bench.py --steps 4 --warmup 3 --no-cpu --quick --no-side-lines --parity-steps 0 > gpurun_out/r02_3_ncu.log 2>&1
tail -2 gpurun_out/r02_3_ncu.log
