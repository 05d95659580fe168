# round-end style verification on one B200: full GPU test suite, smoke, default bench (+ reference arm), small-launch A/B,
# ncu launch list and one full capture of the scoring kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -3 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print('FINAL value %.3f G/s frac %.3f e2e %.3f G/s cpu %s lat %s' % (d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['cpu_baseline'], d['latency_b1']))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-250 gpurun_out/bench_reference.json
BENCH_MODELS=125000 MMP_LANE_MODE=2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_small_phases.json 2> gpurun_out/bench_small_phases.err
grep phases gpurun_out/bench_small_phases.err
MMP_LANE_MODE=2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_1m_phases.json 2> gpurun_out/bench_1m_phases.err
grep phases gpurun_out/bench_1m_phases.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_final.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_place_lanes -c 1 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_final.log 2>&1
tail -1 gpurun_out/ncu_final.log | cut -c1-120
