#!/bin/bash
# Round-end measurement set on one B200 (run under gpurun): scripts/final_cycle.sh TAG
TAG=${1:-x}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
for m in 1 10 100; do
  timeout 300 python bench.py --steps 10 --msgs $m --msgs-all --no-cpu > gpurun_out/bench_${TAG}_msgs$m.json 2>/dev/null
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
# launch list of the same command (serialised, cold cache: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_kernel -s 3 -c 1 \
  -o gpurun_out/prof_scan_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import json
for n in ("bench_$TAG", "bench_${TAG}_msgs1", "bench_${TAG}_msgs10", "bench_${TAG}_msgs100", "bench_ref_$TAG"):
    try:
        d = json.load(open("gpurun_out/%s.json" % n))
        r = d.get("roofline") or {}
        print(n, "value %.0f step %.3f ms scan %.3f frac %s e2e %s" % (d["value"], d["ms_per_step"], r.get("kernel_ms_per_step", 0), r.get("frac"), (d.get("e2e") or {}).get("value")))
    except Exception as e:
        print(n, "failed", e)
PY
