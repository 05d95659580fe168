#!/bin/bash
# round 2, step p (8 GPUs): both instance-shard paths on 8 shards (parity), bench at N = 8
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_instance_shards_gpu.py -q -x -k "sharded_matches_unsharded and 8" > gpurun_out/r02_p8_shards.log 2>&1; tail -5 gpurun_out/r02_p8_shards.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_p8_n8.json 2> gpurun_out/r02_p8_n8.err
python - <<'PY'
import json
for line in open('gpurun_out/r02_p8_n8.json'):
    if line.startswith('{'):
        d = json.loads(line); i = d['instance_sharded']; p = i.get('peer_access', {})
        print('n8 value %.3f G/s %.4f ms; collective %.3f G/s; peer %s' % (d['value'] / 1e9, d['ms_per_step'], i['value'] / 1e9, json.dumps(p)))
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r02_p8_n8.err | tail -5
