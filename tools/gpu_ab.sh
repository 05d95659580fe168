mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_ab_$name.json 2> gpurun_out/bench_ab_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_ab_$name.json').read().strip().splitlines()[-1]); print('$name value %.3f G/s frac %.3f ms %.4f' % (d['value']/1e9, d['roofline']['frac'], d['ms_per_step']))"; }
run ctl_1m MMP_LIB=$PWD/modelmesh_b200/csrc/libmmplace_ctl.so
run new_1m X=1
run ctl_125k MMP_LIB=$PWD/modelmesh_b200/csrc/libmmplace_ctl.so BENCH_MODELS=125000
run new_125k BENCH_MODELS=125000
run new_1m_b X=1
run ctl_1m_b MMP_LIB=$PWD/modelmesh_b200/csrc/libmmplace_ctl.so
