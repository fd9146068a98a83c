#!/bin/bash
SKM_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu > gpurun_out/bt2.json 2> gpurun_out/bt2.err
grep "skm trace" gpurun_out/bt2.err | tail -24
python3 -c "
import json
for l in open('gpurun_out/bt2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['em'])"
