#!/bin/bash
# round 2, step j: lane routine v6 (chunks carry filtered / preferred words and count classes); small launches on k_place_small
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_threads_gpu.py -q -x > gpurun_out/r02_j_pytest.log 2>&1; tail -4 gpurun_out/r02_j_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_j_$name.json 2> gpurun_out/r02_j_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_j_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step']))"; grep phases gpurun_out/r02_j_$name.err; }
run c3
run c5 BENCH_CONFIG=C5
run c5_t BENCH_CONFIG=C5 MMP_LANE_MODE=2
run c5_b96 BENCH_CONFIG=C5 MMP_LANE_BUDGET=96
run c2 BENCH_CONFIG=C2
run c2_small BENCH_CONFIG=C2 MMP_SMALL_MAX=200000
run c2_1m BENCH_CONFIG=C2 BENCH_MODELS=1000000
run c2_1m_small BENCH_CONFIG=C2 BENCH_MODELS=1000000 MMP_SMALL_MAX=2000000
run c3_small MMP_SMALL_MAX=2000000
