#!/bin/bash
# A/B timing of kernel build variants on ONE box (box-to-box noise is +-3%).
#  here (no GPU):  scripts/ab.sh build name1 "-DX=1" name2 "-DX=2" ...   -> variants/lib_<name>.so
#  under gpurun:   scripts/ab.sh run name1 name2 ...                      (3 interleaved rounds)
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p variants
  while [ $# -gt 0 ]; do
    n=$1; f=$2; shift 2
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -shared $f \
      -Xptxas -v -o variants/lib_$n.so dump1090_rs_b200/csrc/b200adsb.cu 2>&1 | grep -A2 "scan7_kernelILb0ELi7384ELb1" | grep -E "registers|spill" | sed "s/^/$n: /"
  done
else
  for round in 1 2 3; do
    for n in "$@"; do
      B200ADSB_LIB=$PWD/variants/lib_$n.so python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$n round $round: scan %.4f ms  step %.4f ms  frac %.4f' % (r['kernel_ms_per_step'], d['ms_per_step'], r['frac']))"
    done
  done
fi
