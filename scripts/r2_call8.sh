#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py -q -m gpu -s 2>&1 | grep -E "SHARDED_OK|passed|failed|Error|error" | tail -12
for ex in symm nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --no-sub --no-e2e --exchange $ex > gpurun_out/bench_n2_$ex.json 2> gpurun_out/bench_n2_$ex.err; grep -v "^\[W\|^W1\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2_$ex.err | tail -5
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n2_$ex.json").read().strip().splitlines()[-1])
    print("$ex", "value %.0f step %.4f parity %s" % (d["value"], d["ms_per_step"], d["parity"]), d["config"]["step_call"][:120])
except Exception as e: print("$ex failed", e)
PY
done
