mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_instance_shards_gpu.py -x -q -m gpu > gpurun_out/pytest_shards.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_shards.log
tail -25 gpurun_out/pytest_shards.log
