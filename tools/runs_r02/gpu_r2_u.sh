#!/bin/bash
# round 2, run u: the bench lines with the in-process clock sampler
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_u_pytest_gpu.log 2>&1; tail -3 $O/r02_u_pytest_gpu.log
timeout 900 python bench.py > $O/r02_u_bench_c3_n1.json 2> $O/r02_u_bench_c3_n1.err
BENCH_CONFIG=C4 timeout 900 python bench.py --no-cpu > $O/r02_u_bench_c4.json 2> $O/r02_u_bench_c4.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_u_bench_c3_n1.json', 'gpurun_out/r02_u_bench_c4.json'):
    d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, d['value'], d['clocks'], (d.get('latency_b1') or {}).get('p50_us'), (d.get('e2e') or {}).get('value'), d.get('commit'))
PY
tail -3 $O/r02_u_bench_c3_n1.err | cut -c1-300
