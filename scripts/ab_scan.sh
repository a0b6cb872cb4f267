#!/bin/bash
# A/B of the stage-1 kernel generations on one box: scripts/ab_scan.sh [rounds]
for round in $(seq 1 ${1:-2}); do
  for v in 6 7; do
    B200ADSB_SCAN=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('scan v$v round $round: scan %.4f ms  step %.4f ms  frac %.4f value %.0f' % (r['kernel_ms_per_step'], d['ms_per_step'], r['frac'], d['value']))"
  done
done
