# round-end style verification on one B200: full GPU test suite, default bench, one full ncu capture of the scoring kernel
# and the launch list of the same command (most important first: the call may be cut short by the GPU-minute budget)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -3 gpurun_out/pytest_final.log
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print('FINAL value %.3f G/s frac %.3f e2e %s e2e_records %s cpu %s lat %s' % (d['value']/1e9, d['roofline']['frac'], d['e2e'], d.get('e2e_records'), (d['cpu_baseline'] or {}).get('value'), d['latency_b1']))"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_place_lanes -c 1 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_final.log 2>&1
tail -1 gpurun_out/ncu_final.log | cut -c1-120
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_final.log 2>&1
echo done
