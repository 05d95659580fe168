#!/bin/bash
# round 2, step n: chunk storage in shared memory; whole GPU suite; benches with the MINB variants
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_churn_gpu.py::test_closed_loop_c4_full_size > gpurun_out/r02_n_pytest.log 2>&1; tail -4 gpurun_out/r02_n_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_n_$name.json 2> gpurun_out/r02_n_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_n_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f lat %s' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step'], {k: (round(v['p50_us'],1), round(v['p99_us'],1)) for k, v in d['latency_b1'].items() if isinstance(v, dict)}))"; tail -2 gpurun_out/r02_n_$name.err | cut -c1-300; }
run c3
run c3_d8 MMP_DIRECT_MINB=8
run c3_d4 MMP_DIRECT_MINB=4
run c5 BENCH_CONFIG=C5
run c5_d8 BENCH_CONFIG=C5 MMP_DIRECT_MINB=8
run c5_d4 BENCH_CONFIG=C5 MMP_DIRECT_MINB=4
run c2 BENCH_CONFIG=C2
