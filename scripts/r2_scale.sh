#!/bin/bash
# N = 2, 4, 8 on one 8-GPU box (run under gpurun --gpus 8): the driver's scaling bench + the sharded GPU test
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29559 \
   bench.py --gpus 8 --steps 20 --no-sub --no-e2e --no-cpu --exchange nccl > gpurun_out/scale_n8_nccl.json 2> gpurun_out/scale_n8_nccl.err
python -c "
import json
d=json.loads(open('gpurun_out/scale_n8_nccl.json').read().strip().splitlines()[-1]); print('N=8 nccl exchange: value %.0f step %.4f' % (d['value'], d['ms_per_step']))"
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n \
     bench.py --gpus $n --steps 20 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  grep -v "^\[W\|^W1\|OMP_NUM\|\*\*\*\*" gpurun_out/scale_n$n.err | tail -4
done
python - <<'PY'
import json
for n in (2,4,8):
    try:
        d=json.loads(open("gpurun_out/scale_n%d.json"%n).read().strip().splitlines()[-1])
        print("N=%d value %.0f step %.4f frac %.4f e2e %.0f parity %s | %s" % (n, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"], d["config"]["step_call"][:70]))
        for k in ("strong_8192","h2d_ceiling","iq_scatter"):
            if k in d: print("   ",k,json.dumps(d[k])[:330])
    except Exception as e:
        print(n,"failed",e)
PY
