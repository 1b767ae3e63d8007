#!/bin/bash
# usage: tools/ab_run.sh name1 name2 ...   (variants built by tools/ab_build.py; "base" = the product build)
for n in "$@"; do
  if [ "$n" = base ]; then unset CHS_LIB_PATH; else export CHS_LIB_PATH=$PWD/cvids_b200/_ab_$n.so; fi
  python bench.py --no-cpu --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$n', 'step %.1f us' % (1000*d['ms_per_step']), 'prep %.1f cand %.1f bricks %.1f' % (1000*r['prepare_ms_per_launch'], 1000*r['candidates_ms_per_launch'], 1000*r['bricks_ms_per_launch']), 'GVox/s %.1f' % d['value'])"
done
