mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_instance_shards_gpu.py -x -q -m gpu > gpurun_out/pytest_shards.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_shards.log
tail -4 gpurun_out/pytest_shards.log
bash tools/gpu_bench_n.sh 2
