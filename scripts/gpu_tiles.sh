#!/bin/bash
# bench value / scan time for several tile sizes (run under gpurun)
for t in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --tile $t > gpurun_out/bench_tile_$t.json 2>gpurun_out/bench_tile_$t.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_tile_$t.json")); r = d["roofline"]
print("tile $t: value %.0f step %.3f ms scan %.3f ms frac %.4f" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"]))
PY
done
