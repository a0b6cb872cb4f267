#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
for round in 1 2; do for n in s0 s1k s3k s6k; do
  B200ADSB_LIB=$PWD/variants/lib_$n.so python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-sub 2>/dev/null | python scripts/benchline.py "$n r$round"
  B200ADSB_LIB=$PWD/variants/lib_$n.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub --buffers 8192 2>/dev/null | python scripts/benchline.py "$n r$round 8192"
done; done
