#!/bin/bash
# round 2, step k: k_place_direct (no landing stages) in three occupancy variants; the resident B = 1 server
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_threads_gpu.py -q -x -k "direct or latency or concurrent" > gpurun_out/r02_k_pytest.log 2>&1; tail -6 gpurun_out/r02_k_pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_k_$name.json 2> gpurun_out/r02_k_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_k_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f lat %s' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step'], {k: (round(v['p50_us'],1), round(v['p99_us'],1)) for k, v in d['latency_b1'].items() if isinstance(v, dict)}))"; tail -2 gpurun_out/r02_k_$name.err | cut -c1-300; }
run c3_d4 MMP_KERNEL=direct
run c3_d6 MMP_KERNEL=direct MMP_DIRECT_MINB=6
run c3_d8 MMP_KERNEL=direct MMP_DIRECT_MINB=8
run c5_d4 BENCH_CONFIG=C5 MMP_KERNEL=direct
run c5_d8 BENCH_CONFIG=C5 MMP_KERNEL=direct MMP_DIRECT_MINB=8
run c2_d4 BENCH_CONFIG=C2 MMP_KERNEL=direct
run c2_d8 BENCH_CONFIG=C2 MMP_KERNEL=direct MMP_DIRECT_MINB=8
run c2_1m_d8 BENCH_CONFIG=C2 BENCH_MODELS=1000000 MMP_KERNEL=direct MMP_DIRECT_MINB=8
