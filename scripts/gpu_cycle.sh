#!/bin/bash
# One build->measure cycle on the GPU box (run under gpurun):
#   scripts/gpu_cycle.sh TAG [pytest|nopytest] [ncu|noncu]
# Writes gpurun_out/bench_TAG.json, gpurun_out/prof_scan_TAG.ncu-rep
TAG=${1:-x}
if [ "${2:-pytest}" = "pytest" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
r = d["roofline"]
print("value %.0f Msps  step %.3f ms  scan %.3f ms decode %.3f ms resolve %.3f ms  frac %.4f  e2e %.0f  cpu %.1f parity %s launches %d" % (
    d["value"], d["ms_per_step"], r["kernel_ms_per_step"], r["decode_kernel_ms_per_step"], r["resolve_kernels_ms_per_step"], r["frac"], (d["e2e"] or {}).get("value", 0),
    (d["cpu_baseline"] or {}).get("value", 0), (d["cpu_baseline"] or {}).get("parity_on_sample"), d["gpu_launches"]))
PY
tail -3 gpurun_out/bench_$TAG.err
if [ "${3:-ncu}" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_kernel -s 3 -c 1 \
    -o gpurun_out/prof_scan_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
fi
