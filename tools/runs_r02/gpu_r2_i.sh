#!/bin/bash
# round 2, step i: ncu source-level captures of the lane kernel (v5) on C3 and C5; remaining GPU tests
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_churn_gpu.py::test_closed_loop_c4_full_size > gpurun_out/r02_i_pytest.log 2>&1; tail -8 gpurun_out/r02_i_pytest.log
for cfg in C3 C5; do
  BENCH_CONFIG=$cfg timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_place_lanes -s 3 -c 1 -f -o gpurun_out/r02_i_ncu_$cfg python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_i_ncu_$cfg.log 2>&1
  tail -2 gpurun_out/r02_i_ncu_$cfg.log
done
ls -la gpurun_out/*.ncu-rep
