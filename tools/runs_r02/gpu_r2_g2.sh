#!/bin/bash
# round 2, step g2 (2 GPUs): the peer-access path of the instance shards, with worker errors reported
cd "$GRAFT_REPO_ROOT"
timeout 400 python -m pytest tests/test_instance_shards_gpu.py -q -x -k "sharded_matches_unsharded and 2" > gpurun_out/r02_g2_shards.log 2>&1; tail -40 gpurun_out/r02_g2_shards.log
nvidia-smi --query-gpu=index,memory.used --format=csv
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_g2_n2.json 2> gpurun_out/r02_g2_n2.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02_g2_n2.json'))
    print('n2 value %.3f G/s' % (d['value'] / 1e9)); print(json.dumps(d.get('instance_sharded'), indent=1))
except Exception as ex:
    print('no json', ex)
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r02_g2_n2.err | tail -8
