#!/bin/bash
# round 2, run w: non-simple (b) in the lane routine: GPU suite, C5 / C3 bench lines
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_w_pytest_gpu.log 2>&1; tail -3 $O/r02_w_pytest_gpu.log
BENCH_CONFIG=C5 timeout 900 python bench.py > $O/r02_w_bench_c5_n1.json 2> $O/r02_w_bench_c5_n1.err
timeout 900 python bench.py --no-cpu > $O/r02_w_bench_c3.json 2> $O/r02_w_bench_c3.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_w_bench_c5_n1.json', 'gpurun_out/r02_w_bench_c3.json'):
    d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, d['value'], d['ms_per_step'], d['clocks'], (d.get('latency_b1') or {}).get('p50_us'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('parity_mismatches_vs_gpu'))
PY
