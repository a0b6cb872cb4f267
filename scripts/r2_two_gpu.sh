#!/bin/bash
# two-GPU validation (run under gpurun --gpus 2): sharded test (both exchanges) + the bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 > gpurun_out/final_n2.json 2> gpurun_out/final_n2.err
grep -v "^\[W\|^W1\|OMP_NUM\|\*\*\*\*" gpurun_out/final_n2.err | tail -4
python -c "
import json
d=json.loads(open('gpurun_out/final_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.0f step %.4f frac %.4f e2e %.0f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity']))
print(d['config']['step_call']); print(d['strong_8192'])"
