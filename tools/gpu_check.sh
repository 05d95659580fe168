mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v2.log
tail -5 gpurun_out/pytest_v2.log
for w in 16 8 12 20; do
  MMP_LANE_WARPS=$w timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_v2_w$w.json 2> gpurun_out/bench_v2_w$w.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v2_w$w.json')); print('warps $w value %.3f G/s frac %.3f e2e %.3f lat %s' % (d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['latency_b1']))"
done
BENCH_MODELS=300000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_place_lanes -c 1 -o gpurun_out/prof_v2 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_v2.log 2>&1
tail -3 gpurun_out/ncu_v2.log | cut -c1-200
