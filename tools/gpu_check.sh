mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_x_$name.json 2> gpurun_out/bench_x_$name.err; python -c "
import json; d=json.load(open('gpurun_out/bench_x_$name.json')); print('$name value %.3f G/s frac %.3f' % (d['value']/1e9, d['roofline']['frac']))"; grep phases gpurun_out/bench_x_$name.err; }
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
run v2ctl MMP_LIB=$PWD/modelmesh_b200/csrc/libmmplace_v2.so
run nofront_w12_s4 MMP_LANE_WARPS=12 MMP_LANE_FRONT=0 MMP_LANE_STAGES=4
run front_w12_s4 MMP_LANE_WARPS=12 MMP_LANE_FRONT=1
run front_w16_s4 MMP_LANE_WARPS=16 MMP_LANE_FRONT=1
run front_w12_s3 MMP_LANE_WARPS=12 MMP_LANE_FRONT=1 MMP_LANE_STAGES=3
run nofront_w12_s3 MMP_LANE_WARPS=12 MMP_LANE_FRONT=0 MMP_LANE_STAGES=3
run nofront_w16_s3 MMP_LANE_WARPS=16 MMP_LANE_FRONT=0 MMP_LANE_STAGES=3
run t_nofront_w12_s4 MMP_LANE_WARPS=12 MMP_LANE_FRONT=0 MMP_LANE_STAGES=4 MMP_LANE_MODE=2
run t_front_w12_s4 MMP_LANE_WARPS=12 MMP_LANE_FRONT=1 MMP_LANE_MODE=2
run t_stream_w12_s4 MMP_LANE_WARPS=12 MMP_LANE_FRONT=0 MMP_LANE_MODE=3 MMP_LANE_STAGES=4
