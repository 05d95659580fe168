#!/bin/bash
# round 2, step b: compressed word lists + chunked second-chance reads in k_place_lanes
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_b.log
tail -3 gpurun_out/r02_pytest_gpu_b.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_b_$name.json 2> gpurun_out/r02_b_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_b_$name.json')); print('$name value %.3f G/s frac %.3f ms %.4f' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step']))"; grep phases gpurun_out/r02_b_$name.err; }
run c3
run c3_t MMP_LANE_MODE=2
for b in 64 96 160 256; do run c5_b$b BENCH_CONFIG=C5 MMP_LANE_BUDGET=$b; done
run c5_t BENCH_CONFIG=C5 MMP_LANE_MODE=2
run c5_t256 BENCH_CONFIG=C5 MMP_LANE_MODE=2 MMP_LANE_BUDGET=256
run c2 BENCH_CONFIG=C2 BENCH_MODELS=100000 BENCH_INSTANCES=1000
run c2_1m BENCH_CONFIG=C2 BENCH_MODELS=1000000 BENCH_INSTANCES=1000
