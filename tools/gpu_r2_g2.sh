#!/bin/bash
# round 2, step g2 (2 GPUs): the peer-access path of the instance shards, with worker errors reported
cd "$GRAFT_REPO_ROOT"
timeout 400 python -m pytest tests/test_instance_shards_gpu.py -q -x -k "sharded_matches_unsharded and 2" > gpurun_out/r02_g2_shards.log 2>&1; tail -40 gpurun_out/r02_g2_shards.log
nvidia-smi --query-gpu=index,memory.used --format=csv
