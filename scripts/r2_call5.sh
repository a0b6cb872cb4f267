#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --steps 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err
python - <<'PY'
import json
for n in ("bench_n1","bench_n2","bench_ref"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%n).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(n,"value %.0f step %.4f frac %s e2e %s parity %s" % (d["value"], d["ms_per_step"], r.get("frac"), (d.get("e2e") or {}).get("value"), d.get("parity")))
        for k in ("strong_8192","h2d_ceiling","iq_scatter","configs0","configs3"):
            if k in d: print("   ",k,json.dumps(d[k])[:400])
        if n=="bench_ref": print("   ", json.dumps(d["cpu_baseline"])[:300])
    except Exception as e:
        print(n,"failed",e)
PY
