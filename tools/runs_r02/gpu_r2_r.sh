#!/bin/bash
# round 2, step r: LRU event kernel with shared-memory staging of hot instances: whole GPU suite (incl. full-size C4), C4 bench
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_r_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_r_pytest_gpu.log
BENCH_CONFIG=C4 timeout 900 python bench.py > gpurun_out/r02_r_bench_c4_churn.json 2> gpurun_out/r02_r_bench_c4_churn.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_r_bench_c4_churn.json') if l.startswith('{')][-1])
print('c4', d['value'], d['unit'], d['ms_per_step'], d['phases_ms'], d['cpu_baseline'].get('parity_mismatching_windows'))
PY
tail -3 gpurun_out/r02_r_bench_c4_churn.err | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_r_$name.json 2> gpurun_out/r02_r_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_r_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step']))"; tail -2 gpurun_out/r02_r_$name.err | cut -c1-300; }
run c5 BENCH_CONFIG=C5
run c3
run c2 BENCH_CONFIG=C2
