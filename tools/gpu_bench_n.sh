# usage: gpu_bench_n.sh N  — the driver's launch line for N ranks, plus (N >= 4) the multi-GPU parity test
N=$1
mkdir -p gpurun_out
if [ "$N" -ge 4 ]; then
  timeout 600 python -m pytest tests/test_instance_shards_gpu.py -x -q -m gpu > gpurun_out/pytest_shards_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_shards_n$N.log
  tail -4 gpurun_out/pytest_shards_n$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.3f G/s ms %.4f e2e %.3f" % (d["value"]/1e9, d["ms_per_step"], (d["e2e"] or {}).get("value", 0)/1e9))
    print("instance_sharded", d.get("instance_sharded"))
except Exception as e:
    print("parse failed", e)
PY
