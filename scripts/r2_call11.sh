#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for round in 1 2 3; do for n in q0 q1; do
  B200ADSB_LIB=$PWD/variants/lib_$n.so python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-sub 2>/dev/null | python scripts/benchline.py "$n r$round"
done; done
