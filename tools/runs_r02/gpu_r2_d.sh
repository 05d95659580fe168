#!/bin/bash
# round 2, step d: fixed checked-load path, device-commit cross-check, threading stress, ncu source captures of k_place_lanes
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_churn_gpu.py -x -q -k "checked_load or device_commit" > gpurun_out/r02_d_churn1.log 2>&1; tail -12 gpurun_out/r02_d_churn1.log
timeout 900 python -m pytest tests/test_threads_gpu.py -x -q > gpurun_out/r02_d_threads.log 2>&1; tail -12 gpurun_out/r02_d_threads.log
BENCH_MODELS=200000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_place_lanes -s 2 -c 1 -o gpurun_out/r02_d_ncu_c3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_d_ncu_c3.log 2>&1; tail -2 gpurun_out/r02_d_ncu_c3.log
BENCH_CONFIG=C5 BENCH_MODELS=200000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_place_lanes -s 2 -c 1 -o gpurun_out/r02_d_ncu_c5 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_d_ncu_c5.log 2>&1; tail -2 gpurun_out/r02_d_ncu_c5.log
ls -la gpurun_out/*.ncu-rep
