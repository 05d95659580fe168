#!/bin/bash
# round 2, step h: lane routine v5 (dense window on shared-memory tables + compressed lists beyond), whole GPU suite, benches
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_churn_gpu.py::test_closed_loop_c4_full_size > gpurun_out/r02_h_pytest.log 2>&1; tail -6 gpurun_out/r02_h_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_h_$name.json 2> gpurun_out/r02_h_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_h_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f lat %s' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step'], {k: (round(v['p50_us'],1), round(v['p99_us'],1)) for k, v in d['latency_b1'].items() if isinstance(v, dict)}))"; grep phases gpurun_out/r02_h_$name.err; }
run c3
run c3_t MMP_LANE_MODE=2
run c3_w16 MMP_LANE_WARPS=16
run c3_w10 MMP_LANE_WARPS=10
run c5 BENCH_CONFIG=C5
run c5_t BENCH_CONFIG=C5 MMP_LANE_MODE=2
run c2 BENCH_CONFIG=C2
