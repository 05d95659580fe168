#!/bin/bash
# round 2, step l: k_place_direct as the default kernel, slot-sorted batches, the single-line server protocol; whole GPU suite
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_churn_gpu.py::test_closed_loop_c4_full_size > gpurun_out/r02_l_pytest.log 2>&1; tail -6 gpurun_out/r02_l_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_l_$name.json 2> gpurun_out/r02_l_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_l_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f lat %s' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step'], {k: (round(v['p50_us'],1), round(v['p99_us'],1)) for k, v in d['latency_b1'].items() if isinstance(v, dict)}))"; tail -2 gpurun_out/r02_l_$name.err | cut -c1-300; }
run c3
run c3_sort MMP_SORT_SLOTS=1
run c5
run c5_nosort BENCH_CONFIG=C5 MMP_SORT_SLOTS=0
run c5_sort BENCH_CONFIG=C5 MMP_SORT_SLOTS=1
run c2 BENCH_CONFIG=C2
BENCH_CONFIG=C4 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_l_c4.json 2> gpurun_out/r02_l_c4.err; python -c "
import json; d=json.load(open('gpurun_out/r02_l_c4.json')); print('c4', d['value'], d['unit'], d.get('ms_per_step'))"; tail -3 gpurun_out/r02_l_c4.err | cut -c1-300
