mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_instance_shards_gpu.py tests/test_parity_gpu.py::test_batch_properties -x -q -m gpu > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_quick.log
tail -12 gpurun_out/pytest_quick.log
