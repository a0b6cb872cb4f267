#!/bin/bash
# round-2 evidence set on one B200 (run under gpurun): scripts/r2_profile.sh TAG
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-sub > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_kernel -s 3 -c 1 \
  -o gpurun_out/prof_scan_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-sub > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/ncu_$TAG.log | cut -c1-150
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/sanitize_$TAG.txt
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 >> gpurun_out/sanitize_$TAG.txt
cat gpurun_out/sanitize_$TAG.txt
python - <<PY
import json
for n in ("bench_$TAG", "bench_ref_$TAG"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(n, "value %.0f step %.4f ms scan %s frac %s e2e %s parity %s" % (d["value"], d["ms_per_step"], r.get("kernel_ms_per_step"), r.get("frac"), (d.get("e2e") or {}).get("value"), d.get("parity_on_sample")))
    except Exception as e:
        print(n, "failed", e)
PY
