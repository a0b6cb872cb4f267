#!/bin/bash
# A/B of the stage-1 kernel generations on one box: scripts/ab_scan.sh ROUNDS VER...   (VER = 6 or 7)
rounds=${1:-2}; shift
vers=${@:-6 7}
for round in $(seq 1 $rounds); do
  for v in $vers; do
    ver=${v%%:*}; chunk=0; [[ $v == *:* ]] && chunk=${v##*:}
    B200ADSB_SCAN=$ver B200ADSB_CHUNK=$chunk python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('scan v$v round $round: scan %.4f ms  step %.4f ms  frac %.4f value %.0f' % (r['kernel_ms_per_step'], d['ms_per_step'], r['frac'], d['value']))"
  done
done
