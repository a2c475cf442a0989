#!/bin/bash
# GPU tuning sweep (run under gpurun): SpMM loads-in-flight x occupancy variants on config 5
for v in 0 1 2 3 4 5 6; do
  echo "variant $v"
  CBRS_SPMM_VARIANT=$v timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --catalog-users 2048 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  ms/step %.1f  spmm %.1f ms frac %.3f  value %.3e  pairs %.3e (%.1f ms)'%(d['ms_per_step'], r['launch_ms'], r['frac'], d['value'], d['pairs']['value'], d['pairs']['ms']))
"
done
