#!/bin/bash
# round 2, call 1: GPU suite on the refactored library, A/B of occupancy / load-path variants, ncu of HEAD
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
scripts/ab.sh run base ring2 mb8 ring0 ring0mb8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_kernel -s 3 -c 1 \
  -o gpurun_out/prof_scan_r2a python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_r2a.log 2>&1
tail -2 gpurun_out/ncu_r2a.log | cut -c1-200
