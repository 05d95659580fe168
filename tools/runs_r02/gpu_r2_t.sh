#!/bin/bash
# round 2, run t at HEAD (2-D rank count, commit leg in the bench): GPU suite, default bench, C5 / C2 / C4 benches
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_t_pytest_gpu.log 2>&1; tail -3 $O/r02_t_pytest_gpu.log
timeout 900 python bench.py > $O/r02_t_bench_c3_n1.json 2> $O/r02_t_bench_c3_n1.err
BENCH_CONFIG=C5 timeout 900 python bench.py > $O/r02_t_bench_c5_n1.json 2> $O/r02_t_bench_c5_n1.err
BENCH_CONFIG=C2 timeout 900 python bench.py > $O/r02_t_bench_c2_n1.json 2> $O/r02_t_bench_c2_n1.err
BENCH_CONFIG=C4 timeout 900 python bench.py > $O/r02_t_bench_c4_churn.json 2> $O/r02_t_bench_c4_churn.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_t_bench_reference_arm.json 2> $O/r02_t_bench_reference_arm.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_t_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], d.get('value'), d.get('unit'), 'ms', d.get('ms_per_step'), 'frac', (d.get('roofline') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'mism', (d.get('cpu_baseline') or {}).get('parity_mismatches_vs_gpu'), (d.get('latency_b1') or {}).get('p50_us'), d.get('commit'), d.get('phases_ms'), d.get('mmp_fleet_commit_ms'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
tail -3 $O/r02_t_bench_c3_n1.err | cut -c1-300
