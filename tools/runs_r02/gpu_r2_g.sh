#!/bin/bash
# round 2, step g (2 GPUs): instance shards over peer memory (k_place_dealt) -- parity on 2 shards, bench at N = 2
cd "$GRAFT_REPO_ROOT"
nvidia-smi topo -m > gpurun_out/r02_g_topo.txt 2>&1
timeout 900 python -m pytest tests/test_instance_shards_gpu.py -q -k "sharded_matches_unsharded and 2" > gpurun_out/r02_g_shards.log 2>&1; tail -15 gpurun_out/r02_g_shards.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_g_n2.json 2> gpurun_out/r02_g_n2.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02_g_n2.json'))
    print('n2 value %.3f G/s' % (d['value'] / 1e9)); print(json.dumps(d.get('instance_sharded'), indent=1))
except Exception as ex:
    print('no json', ex)
PY
tail -5 gpurun_out/r02_g_n2.err
