#!/bin/bash
# round 2, step g3 (2 GPUs): bench at N = 2 with the timing split of the peer-access step
cd "$GRAFT_REPO_ROOT"
for mb in 6 4; do
MMP_DEALT_MINB=$mb timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_g3_n2_mb$mb.json 2> gpurun_out/r02_g3_n2_mb$mb.err
python - <<PY
import json
for line in open('gpurun_out/r02_g3_n2_mb$mb.json'):
    if line.startswith('{'):
        d = json.loads(line); p = d['instance_sharded']['peer_access']
        print('minb $mb: n2 value %.3f G/s; collective %.3f; peer %.3f G/s %.4f ms; last step %s' % (d['value'] / 1e9, d['instance_sharded']['value'] / 1e9, p['value'] / 1e9, p['ms_per_step'], p['rank0_last_step_ms']))
PY
done
