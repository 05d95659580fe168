#!/bin/bash
# round 2, last single-GPU run at HEAD: GPU suite, smoke, the bench lines profiles/ keeps
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_s_pytest_gpu.log 2>&1; tail -3 $O/r02_s_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > $O/r02_s_bench_c3_n1.json 2> $O/r02_s_bench_c3_n1.err
BENCH_CONFIG=C5 timeout 900 python bench.py > $O/r02_s_bench_c5_n1.json 2> $O/r02_s_bench_c5_n1.err
BENCH_CONFIG=C2 timeout 900 python bench.py > $O/r02_s_bench_c2_n1.json 2> $O/r02_s_bench_c2_n1.err
BENCH_CONFIG=C4 timeout 900 python bench.py > $O/r02_s_bench_c4_churn.json 2> $O/r02_s_bench_c4_churn.err
BENCH_MODELS=125000 timeout 600 python bench.py --no-cpu --no-e2e > $O/r02_s_bench_c3_125k.json 2> $O/r02_s_bench_c3_125k.err
BENCH_MODELS=125000 MMP_DIRECT_MINB=4 timeout 600 python bench.py --no-cpu --no-e2e > $O/r02_s_bench_c3_125k_d4.json 2> $O/r02_s_bench_c3_125k_d4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_s_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], d.get('value'), d.get('unit'), 'ms', d.get('ms_per_step'), 'frac', (d.get('roofline') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'mism', (d.get('cpu_baseline') or {}).get('parity_mismatches_vs_gpu'), (d.get('latency_b1') or {}).get('p50_us'), d.get('phases_ms'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
