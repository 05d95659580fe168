#!/bin/bash
# round-2 baseline on one B200: GPU tests, then bench lines for C3 / C5 / C2 (kept under gpurun_out/)
set -x
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_a.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_a.json 2> gpurun_out/r02_bench_c3_a.err
BENCH_CONFIG=C5 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c5_a.json 2> gpurun_out/r02_bench_c5_a.err
BENCH_CONFIG=C2 BENCH_MODELS=100000 BENCH_INSTANCES=1000 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2_a.json 2> gpurun_out/r02_bench_c2_a.err
tail -c 600 gpurun_out/r02_pytest_gpu_a.log
