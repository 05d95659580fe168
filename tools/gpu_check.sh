mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_stream.log
tail -5 gpurun_out/pytest_stream.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
cat gpurun_out/bench_stream.json | cut -c1-300
BENCH_MODELS=300000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_place_lanes -c 1 -o gpurun_out/prof_stream -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log | cut -c1-300
