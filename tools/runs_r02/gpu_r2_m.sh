#!/bin/bash
# round 2, step m: ncu captures of k_place_direct (C3, C5 slot-sorted), C4 bench
cd "$GRAFT_REPO_ROOT"
for cfg in C3 C5; do
  BENCH_CONFIG=$cfg timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_place_direct -s 3 -c 1 -f -o gpurun_out/r02_m_ncu_direct_$cfg python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_m_ncu_$cfg.log 2>&1
  tail -2 gpurun_out/r02_m_ncu_$cfg.log
done
BENCH_CONFIG=C4 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_m_c4.json 2> gpurun_out/r02_m_c4.err; python -c "
import json; d=json.load(open('gpurun_out/r02_m_c4.json')); print('c4', d['value'], d['unit'], d.get('ms_per_step'))"; tail -3 gpurun_out/r02_m_c4.err | cut -c1-300
