#!/bin/bash
# round 2, step f: lane routine v4 (window loop + chunk loop), latency kernel, registry scans, remaining churn tests
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_churn_gpu.py -q -k "checked_load or device_commit" > gpurun_out/r02_f_churn1.log 2>&1; tail -8 gpurun_out/r02_f_churn1.log
timeout 900 python -m pytest tests/test_registry_scans_gpu.py -q > gpurun_out/r02_f_scans.log 2>&1; tail -12 gpurun_out/r02_f_scans.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_churn_gpu.py::test_closed_loop_c4_full_size --deselect tests/test_registry_scans_gpu.py > gpurun_out/r02_f_pytest.log 2>&1; tail -6 gpurun_out/r02_f_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_f_$name.json 2> gpurun_out/r02_f_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_f_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f lat %s' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step'], {k: (round(v['p50_us'],1), round(v['p99_us'],1)) for k, v in d['latency_b1'].items() if isinstance(v, dict)}))"; grep phases gpurun_out/r02_f_$name.err; }
run c3
run c3_t MMP_LANE_MODE=2
run c5 BENCH_CONFIG=C5
run c5_t BENCH_CONFIG=C5 MMP_LANE_MODE=2
run c2 BENCH_CONFIG=C2
