#!/bin/bash
# round 2, final single-GPU measurement run: the artifacts profiles/ holds for this round
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_q_pytest_gpu.log 2>&1; tail -3 $O/r02_q_pytest_gpu.log
# ncu: full captures of the scoring kernel (C3, C5), then the launch list of the default bench command
for cfg in C3 C5; do
  BENCH_CONFIG=$cfg timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_place_direct -s 3 -c 1 -f -o $O/r02_q_ncu_direct_$cfg python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/r02_q_ncu_$cfg.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_q_ncu_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_q_launch_list_bench.log 2>&1
BENCH_CONFIG=C4 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_q_ncu_launch_list_c4.csv python bench.py --steps 2 --warmup 3 > $O/r02_q_launch_list_c4.log 2>&1
# benches (never under a profiler)
timeout 900 python bench.py > $O/r02_q_bench_c3_n1.json 2> $O/r02_q_bench_c3_n1.err; tail -c 400 $O/r02_q_bench_c3_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_q_bench_reference_arm.json 2> $O/r02_q_bench_reference_arm.err
BENCH_CONFIG=C5 timeout 900 python bench.py > $O/r02_q_bench_c5_n1.json 2> $O/r02_q_bench_c5_n1.err
BENCH_CONFIG=C5 MMP_LANE_BUDGET=512 timeout 900 python bench.py --no-cpu --no-e2e > $O/r02_q_bench_c5_budget512.json 2> $O/r02_q_bench_c5_budget512.err
BENCH_CONFIG=C2 timeout 900 python bench.py > $O/r02_q_bench_c2_n1.json 2> $O/r02_q_bench_c2_n1.err
BENCH_CONFIG=C4 timeout 900 python bench.py > $O/r02_q_bench_c4_churn.json 2> $O/r02_q_bench_c4_churn.err
MMP_KERNEL=lanes MMP_LANE_MODE=2 timeout 900 python bench.py --no-cpu --no-e2e > $O/r02_q_bench_c3_lanes.json 2> $O/r02_q_bench_c3_lanes.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_q_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], d.get('metric', '')[:30], d.get('value'), d.get('unit'), 'frac', (d.get('roofline') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'mism', (d.get('cpu_baseline') or {}).get('parity_mismatches_vs_gpu'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
