#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 \
   bench.py --gpus 8 --steps 20 --no-sub --no-e2e > gpurun_out/n8_min.json 2> gpurun_out/n8_min.err
python -c "
import json
d=json.loads(open('gpurun_out/n8_min.json').read().strip().splitlines()[-1]); print('N=8 value %.0f step %.4f parity %s' % (d['value'], d['ms_per_step'], d['parity'])); print(d['config']['step_call'][:140])" || grep -v "^\[W\|^W1\|OMP_NUM\|\*\*\*\*" gpurun_out/n8_min.err | tail -5
