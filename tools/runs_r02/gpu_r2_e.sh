#!/bin/bash
# round 2, step e: chunk-synchronous lane routine (decide_stream v3) + remaining churn tests
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_churn_gpu.py -q -k "checked_load or device_commit" > gpurun_out/r02_e_churn1.log 2>&1; tail -12 gpurun_out/r02_e_churn1.log
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_churn_gpu.py::test_closed_loop_c4_full_size > gpurun_out/r02_e_pytest.log 2>&1; tail -5 gpurun_out/r02_e_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_e_$name.json 2> gpurun_out/r02_e_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_e_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step']))"; grep phases gpurun_out/r02_e_$name.err; }
run c3
run c3_t MMP_LANE_MODE=2
run c5 BENCH_CONFIG=C5
run c5_t BENCH_CONFIG=C5 MMP_LANE_MODE=2
run c5_b96 BENCH_CONFIG=C5 MMP_LANE_BUDGET=96
run c5_b320 BENCH_CONFIG=C5 MMP_LANE_BUDGET=320
run c2 BENCH_CONFIG=C2
run c2_1m BENCH_CONFIG=C2 BENCH_MODELS=1000000
BENCH_MODELS=200000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_place_lanes -s 2 -c 1 -o gpurun_out/r02_e_ncu_c3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_e_ncu_c3.log 2>&1; tail -1 gpurun_out/r02_e_ncu_c3.log
