#!/bin/bash
# quick tuning sweep of the persistent kernel geometry on the bench workload
for t in ${TUNINGS:-0 1 22}; do
  HMCX_TUNING=$t python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tuning', $t, 'ms/step', round(l['ms_per_step'],3), 'value %.3e' % l['value'], 'kernel_ms', round(l['roofline']['kernel_ms'],3), 'stream GB/s', round(l['roofline_streaming']['achieved']), 'e2e %.3e' % l['e2e']['value'], 'acc', round(l['accept_rate'],4))"
done
