#!/bin/bash
# round 2, step c: churn tests (first run), whole GPU suite, kernel A/B after the synchronised chunk refills
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_churn_gpu.py -x -q -k "checked_load or device_commit" > gpurun_out/r02_c_churn1.log 2>&1; tail -15 gpurun_out/r02_c_churn1.log
timeout 900 python -m pytest tests/test_churn_gpu.py -x -q -k "small" > gpurun_out/r02_c_churn2.log 2>&1; tail -25 gpurun_out/r02_c_churn2.log
timeout 900 python -m pytest tests/test_churn_gpu.py -x -q -k "full_size" > gpurun_out/r02_c_churn3.log 2>&1; tail -15 gpurun_out/r02_c_churn3.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_churn_gpu.py > gpurun_out/r02_c_pytest.log 2>&1; tail -5 gpurun_out/r02_c_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_c_$name.json 2> gpurun_out/r02_c_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_c_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step']))"; grep phases gpurun_out/r02_c_$name.err; }
run c3
run c3_t MMP_LANE_MODE=2
run c5 BENCH_CONFIG=C5
run c5_t BENCH_CONFIG=C5 MMP_LANE_MODE=2
run c5_b96 BENCH_CONFIG=C5 MMP_LANE_BUDGET=96
run c2 BENCH_CONFIG=C2 BENCH_MODELS=100000 BENCH_INSTANCES=1000
run c2_w16 BENCH_CONFIG=C2 BENCH_MODELS=100000 BENCH_INSTANCES=1000 MMP_LANE_WARPS=16
run c2_1m BENCH_CONFIG=C2 BENCH_MODELS=1000000 BENCH_INSTANCES=1000
run c2_1m_w16 BENCH_CONFIG=C2 BENCH_MODELS=1000000 BENCH_INSTANCES=1000 MMP_LANE_WARPS=16
